/*
 * aad_oracle.c -- plain-C CPU restatement of the AAD ADPCM hot path.
 * TEST INFRASTRUCTURE ONLY (see aad_oracle.h).  Parity: PINNED (tests/test_oracle.py).
 *
 * Written from the behaviour of the reference, not from its text: one chain structure
 * shared by encoder and decoder, int16 PCM at the boundary, any channel count up to 8.
 * Each function names the reference lines whose behaviour it restates.
 */
#include "aad_oracle.h"
#include "aad_oracle_tables.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define TAPS 4
#define FILE_HEADER_BYTES 31           /* src/aad.h:22 */
#define CHANNEL_HEADER_BYTES 18        /* src/aad_internal.h:37 : 2 + 4*(2+2) */
#define FORMAT_VERSION 4u              /* src/aad.h:10 */
#define CODEC_VERSION 18u              /* src/aad.h:7 */
#define INDEX_MAX (255 << 4)           /* src/aad_tables.h:38-39 */

static const uint16_t k_step[256] = AAD_ORACLE_STEP_TABLE_INIT;
static const int16_t k_delta2[2] = AAD_ORACLE_DELTA2_INIT;
static const int16_t k_delta3[4] = AAD_ORACLE_DELTA3_INIT;
static const int16_t k_delta4[8] = AAD_ORACLE_DELTA4_INIT;

/* ---- wrapping int32 helpers ---------------------------------- */
static inline int32_t wmul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
static inline int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
static inline int32_t clamp16(int32_t v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }

/* One adaptive chain: 4-tap sign-data LMS predictor + adaptive step index. */
struct chain {
  int32_t hist[TAPS];   /* hist[0] = newest reconstructed sample (int16 range) */
  int32_t weight[TAPS]; /* Q15 */
  int32_t index;        /* Q4 step index */
  int32_t qdiff;        /* last dequantised difference (quantize_error in the reference) */
};

static inline const int16_t *delta_table(uint32_t bits)
{
  return bits == 4 ? k_delta4 : (bits == 3 ? k_delta3 : k_delta2);
}

/* src/aad_tables.h:28 */
static inline int32_t chain_stepsize(const struct chain *c) { return k_step[(c->index + 8) >> 4]; }

/* src/aad_encoder.c:358-363 == src/aad_decoder.c:290-295 */
static inline int32_t chain_predict(const struct chain *c)
{
  int32_t acc = 1 << 14;
  for (int k = 0; k < TAPS; k++) acc = wadd(acc, wmul(c->hist[k], c->weight[k]));
  return acc >> 15;
}

/* Everything after the code is known: dequantise, reconstruct, adapt.
 * src/aad_encoder.c:378-406 and src/aad_decoder.c:283-315 perform the same updates
 * (the order of the independent ones differs, the results do not). */
static inline int32_t chain_absorb(struct chain *c, uint32_t code, uint32_t bits, int32_t predict, int32_t step)
{
  const uint32_t signbit = 1u << (bits - 1);
  const uint32_t mag = code & (signbit - 1);
  int32_t qdiff = (step * (int32_t)(2 * mag + 1)) >> (bits - 1);
  if (code & signbit) qdiff = -qdiff;

  /* step index: src/aad_tables.h:31-43 (int16 narrowing then clip) */
  int32_t idx = (int16_t)(c->index + delta_table(bits)[mag]);
  c->index = idx < 0 ? 0 : (idx > INDEX_MAX ? INDEX_MAX : idx);
  c->qdiff = qdiff;

  const int32_t recon = clamp16(wadd(qdiff, predict));
  for (int k = 0; k < TAPS; k++)
    c->weight[k] = wadd(c->weight[k], wadd(wmul(qdiff, c->hist[k]), 1 << 14) >> 18);
  c->hist[3] = c->hist[2];
  c->hist[2] = c->hist[1];
  c->hist[1] = c->hist[0];
  c->hist[0] = (int16_t)recon;
  return recon;
}

/* src/aad_encoder.c:343-410 */
static inline uint32_t chain_encode_sample(struct chain *c, int32_t sample, uint32_t bits)
{
  const uint32_t signbit = 1u << (bits - 1);
  const int32_t maxmag = (int32_t)signbit - 1;
  const int32_t step = chain_stepsize(c);
  const int32_t predict = chain_predict(c);
  const int32_t diff = wsub(sample, predict);
  const int neg = diff < 0;
  const int32_t mag_in = neg ? wsub(0, diff) : diff;
  int32_t q = (int32_t)((uint32_t)mag_in << (bits - 2)) / step; /* operands non-negative */
  if (q > maxmag) q = maxmag;
  const uint32_t code = (uint32_t)q | (neg ? signbit : 0u);
  chain_absorb(c, code, bits, predict, step);
  return code;
}

/* src/aad_decoder.c:269-318 */
static inline int32_t chain_decode_sample(struct chain *c, uint32_t code, uint32_t bits)
{
  const int32_t step = chain_stepsize(c);
  const int32_t predict = chain_predict(c);
  return chain_absorb(c, code, bits, predict, step);
}

/* ---- geometry / headers ------------------------------------------------------------ */

static uint32_t gcd_u32(uint32_t a, uint32_t b) { while (b) { uint32_t t = a % b; a = b; b = t; } return a; }

/* samples per interleave group and bytes per group per channel: 4-bit 2/1, 3-bit 8/3, 2-bit 4/1 */
static inline uint32_t group_bytes(uint32_t bits) { return (8 * bits / gcd_u32(8, bits)) / 8; }
static inline uint32_t group_samples(uint32_t bits) { return group_bytes(bits) * 8 / bits; }

int aad_oracle_geometry(uint32_t max_block_size, uint32_t channels, uint32_t bits,
                        uint32_t *block_size, uint32_t *samples_per_block)
{
  if (!block_size) return AAD_ORACLE_INVALID_ARGUMENT;
  if (channels == 0 || channels > AAD_ORACLE_MAX_CHANNELS || bits == 0 || bits > 4)
    return AAD_ORACLE_INVALID_FORMAT;
  const uint32_t hdr = CHANNEL_HEADER_BYTES * channels;
  if (max_block_size < hdr) return AAD_ORACLE_INVALID_FORMAT;
  /* bits == 1 passes the reference's range check too (src/aad_encoder.c:100-103) */
  const uint32_t unit = channels * group_bytes(bits);
  const uint32_t units = (max_block_size - hdr) / unit;
  *block_size = (uint16_t)(hdr + units * unit);
  if (samples_per_block) *samples_per_block = units * group_samples(bits) + TAPS;
  return AAD_ORACLE_OK;
}

static inline void put_be16(uint8_t **p, uint32_t v) { (*p)[0] = (uint8_t)(v >> 8); (*p)[1] = (uint8_t)v; *p += 2; }
static inline void put_be32(uint8_t **p, uint32_t v) { put_be16(p, v >> 16); put_be16(p, v & 0xFFFF); }
static inline uint32_t get_be16(const uint8_t *p) { return ((uint32_t)p[0] << 8) | p[1]; }
static inline uint32_t get_be32(const uint8_t *p) { return (get_be16(p) << 16) | get_be16(p + 2); }

static int info_is_valid(const struct aad_oracle_info *h, int check_versions)
{
  if (check_versions && (h->format_version != FORMAT_VERSION || h->codec_version != CODEC_VERSION)) return 0;
  if (h->channels == 0 || h->channels > AAD_ORACLE_MAX_CHANNELS) return 0;
  if (h->num_samples == 0 || h->sampling_rate == 0) return 0;
  if (h->bits < 2 || h->bits > 4) return 0;
  if (h->block_size <= CHANNEL_HEADER_BYTES * h->channels) return 0;
  if (h->samples_per_block == 0) return 0;
  if (h->ms >= 2) return 0;
  if (h->ms == 1 && h->channels == 1) return 0;
  return 1;
}

int aad_oracle_write_header(const struct aad_oracle_info *info, uint8_t *out, size_t cap)
{
  if (!info || !out) return AAD_ORACLE_INVALID_ARGUMENT;
  if (cap < FILE_HEADER_BYTES) return AAD_ORACLE_INSUFFICIENT_DATA;
  if (!info_is_valid(info, 0)) return AAD_ORACLE_INVALID_FORMAT;
  uint8_t *p = out;
  *p++ = 'A'; *p++ = 'A'; *p++ = 'D'; *p++ = 0;
  put_be32(&p, FORMAT_VERSION);           /* struct fields ignored: src/aad_encoder.c:195-200 */
  put_be32(&p, CODEC_VERSION);
  put_be16(&p, info->channels);
  put_be32(&p, info->num_samples);
  put_be32(&p, info->sampling_rate);
  put_be16(&p, info->bits);
  put_be16(&p, info->block_size);
  put_be32(&p, info->samples_per_block);
  *p++ = (uint8_t)info->ms;
  return AAD_ORACLE_OK;
}

int aad_oracle_read_header(const uint8_t *data, size_t size, struct aad_oracle_info *info, int validate)
{
  if (!data || !info) return AAD_ORACLE_INVALID_ARGUMENT;
  if (size < FILE_HEADER_BYTES) return AAD_ORACLE_INSUFFICIENT_DATA;
  if (data[0] != 'A' || data[1] != 'A' || data[2] != 'D' || data[3] != 0) return AAD_ORACLE_INVALID_FORMAT;
  struct aad_oracle_info h;
  h.format_version = get_be32(data + 4);
  h.codec_version = get_be32(data + 8);
  h.channels = get_be16(data + 12);
  h.num_samples = get_be32(data + 14);
  h.sampling_rate = get_be32(data + 18);
  h.bits = get_be16(data + 22);
  h.block_size = get_be16(data + 24);
  h.samples_per_block = get_be32(data + 26);
  h.ms = data[30];
  if (validate && !info_is_valid(&h, 1)) return AAD_ORACLE_INVALID_FORMAT;
  *info = h;
  return AAD_ORACLE_OK;
}

/* ---- encoder ----------------------------------------------------------------------- */

/* src/aad_encoder.c:413-428 */
static void lr_to_ms(int32_t *l, int32_t *r, uint32_t n)
{
  for (uint32_t i = 0; i < n; i++) {
    const int32_t mid = wadd(l[i], r[i]) >> 1, side = wsub(l[i], r[i]) >> 1;
    l[i] = clamp16(mid);
    r[i] = clamp16(side);
  }
}

/* src/aad_encoder.c:431-467: dry-run one block from `c`, return sqrt(mean wrapped q^2) */
static double chain_trial_rmse(struct chain *c, const int32_t *x, uint32_t n, uint32_t bits)
{
  if (n < TAPS) return 0.0;                    /* state untouched */
  for (int k = 0; k < TAPS; k++) c->hist[TAPS - 1 - k] = (int16_t)x[k];
  double sum = 0.0;
  for (uint32_t i = TAPS; i < n; i++) {
    chain_encode_sample(c, x[i], bits);
    sum += (double)wmul(c->qdiff, c->qdiff);   /* 32-bit wrapping product, then widened */
  }
  return sqrt(sum / n);
}

struct enc_ctx {
  struct aad_oracle_info info;
  uint32_t trials;
  struct chain chain[AAD_ORACLE_MAX_CHANNELS];
  int32_t *cur[AAD_ORACLE_MAX_CHANNELS];   /* samples_per_block each */
  int32_t *prev[AAD_ORACLE_MAX_CHANNELS];
};

static void load_block(const struct enc_ctx *e, int32_t *const *dst, const int16_t *pcm, size_t ch_stride,
                       uint32_t first, uint32_t n, int zero_fill)
{
  const uint32_t C = e->info.channels;
  for (uint32_t c = 0; c < C; c++) {
    if (zero_fill) memset(dst[c], 0, sizeof(int32_t) * e->info.samples_per_block);
    for (uint32_t i = 0; i < n; i++) dst[c][i] = pcm[c * ch_stride + first + i];
  }
  if (C >= 2 && e->info.ms == 1) lr_to_ms(dst[0], dst[1], n);
}

/* src/aad_encoder.c:470-562 */
static void search_start_state(struct enc_ctx *e, const int16_t *pcm, size_t ch_stride, uint32_t progress, uint32_t n)
{
  const uint32_t C = e->info.channels, bits = e->info.bits, spb = e->info.samples_per_block;
  const int have_prev = progress >= spb;
  load_block(e, e->cur, pcm, ch_stride, progress, n, 0);
  if (have_prev) load_block(e, e->prev, pcm, ch_stride, progress - spb, spb, 0);
  for (uint32_t c = 0; c < C; c++) {
    struct chain probe = e->chain[c];
    struct chain best = e->chain[c];
    double best_rmse = chain_trial_rmse(&probe, e->cur[c], n, bits);   /* baseline */
    struct chain run = e->chain[c];
    for (uint32_t t = 0; t < e->trials; t++) {
      if (have_prev) (void)chain_trial_rmse(&run, e->prev[c], spb, bits);
      const struct chain candidate = run;
      const double rmse = chain_trial_rmse(&run, e->cur[c], n, bits);
      if (best_rmse > rmse) { best_rmse = rmse; best = candidate; }     /* NaN compares false */
    }
    e->chain[c] = best;
  }
}

/* src/aad_encoder.c:565-727 */
static uint32_t encode_block(struct enc_ctx *e, const int16_t *pcm, size_t ch_stride, uint32_t progress, uint32_t n,
                             uint8_t *out)
{
  const uint32_t C = e->info.channels, bits = e->info.bits;
  uint8_t *p = out;
  load_block(e, e->cur, pcm, ch_stride, progress, n, 1);
  for (uint32_t c = 0; c < C; c++) {
    struct chain *ch = &e->chain[c];
    for (uint32_t i = 0; i < TAPS; i++) ch->hist[TAPS - 1 - i] = (i < n) ? (int16_t)e->cur[c][i] : 0;
  }
  for (uint32_t c = 0; c < C; c++) {
    struct chain *ch = &e->chain[c];
    int32_t maxabs = 0;
    for (int k = 0; k < TAPS; k++) {
      const int32_t a = ch->weight[k] >= 0 ? ch->weight[k] : wsub(0, ch->weight[k]);
      if (maxabs < a) maxabs = a;
    }
    uint32_t shift = 0;
    while (maxabs > 32767) { maxabs >>= 1; shift++; }
    const int32_t keep = (int32_t)~((1u << shift) - 1u);
    for (int k = 0; k < TAPS; k++) ch->weight[k] &= keep;          /* live state is rounded too */
    put_be16(&p, (((uint32_t)ch->index << 4) | (shift & 0xF)) & 0xFFFF);
    for (int k = 0; k < TAPS; k++) {
      put_be16(&p, (uint32_t)(ch->weight[k] >> shift) & 0xFFFF);
      put_be16(&p, (uint32_t)ch->hist[k] & 0xFFFF);
    }
  }
  const uint32_t gs = group_samples(bits), gb = group_bytes(bits);
  for (uint32_t i = TAPS; i < n; i += gs) {
    for (uint32_t c = 0; c < C; c++) {
      uint32_t packed = 0;
      for (uint32_t j = 0; j < gs; j++)
        packed = (packed << bits) | chain_encode_sample(&e->chain[c], e->cur[c][i + j], bits);
      for (uint32_t b = gb; b-- > 0;) *p++ = (uint8_t)(packed >> (8 * b));
    }
  }
  return (uint32_t)(p - out);
}

/* exact bytes the block loop emits: full blocks + a ragged tail of whole groups
 * (src/aad_encoder.c:663,678,704 step by whole groups past num_samples) */
static uint64_t encoded_size(const struct aad_oracle_info *h)
{
  const uint32_t spb = h->samples_per_block;
  const uint64_t full = h->num_samples / spb;
  const uint32_t tail = h->num_samples % spb;
  uint64_t bytes = FILE_HEADER_BYTES + full * h->block_size;
  if (tail) {
    const uint32_t gs = group_samples(h->bits), gb = group_bytes(h->bits);
    const uint32_t groups = tail > TAPS ? (tail - TAPS + gs - 1) / gs : 0;
    bytes += (uint64_t)CHANNEL_HEADER_BYTES * h->channels + (uint64_t)groups * gb * h->channels;
  }
  return bytes;
}

int64_t aad_oracle_encode(const int16_t *pcm, size_t ch_stride, uint32_t channels, uint32_t num_samples,
                          uint32_t sampling_rate, uint32_t bits, uint32_t max_block_size, uint32_t ms,
                          uint32_t trials, struct aad_oracle_chain *state, uint8_t *out, size_t cap)
{
  if (!pcm || !out) return -AAD_ORACLE_INVALID_ARGUMENT;
  struct enc_ctx e;
  memset(&e, 0, sizeof(e));
  /* parameter checks: src/aad_encoder.c:741-770 */
  if (bits == 0 || bits > 4 || ms >= 2) return -AAD_ORACLE_INVALID_FORMAT;
  uint32_t bs = 0, spb = 0;
  if (aad_oracle_geometry(max_block_size, channels, bits, &bs, &spb) != AAD_ORACLE_OK) return -AAD_ORACLE_INVALID_FORMAT;
  e.info.format_version = FORMAT_VERSION;
  e.info.codec_version = CODEC_VERSION;
  e.info.channels = channels;
  e.info.num_samples = num_samples;
  e.info.sampling_rate = sampling_rate;
  e.info.bits = bits;
  e.info.block_size = bs;
  e.info.samples_per_block = spb;
  e.info.ms = ms;
  e.trials = trials;
  const int rc = aad_oracle_write_header(&e.info, out, cap);
  if (rc != AAD_ORACLE_OK) return -rc;
  if (encoded_size(&e.info) > cap) return -AAD_ORACLE_INSUFFICIENT_BUFFER; /* the reference only asserts */
  for (uint32_t c = 0; c < channels; c++) {
    e.cur[c] = (int32_t *)malloc(sizeof(int32_t) * spb);
    e.prev[c] = (int32_t *)malloc(sizeof(int32_t) * spb);
    if (state) {
      memcpy(e.chain[c].weight, state[c].weight, sizeof(state[c].weight));
      e.chain[c].index = state[c].stepsize_index;
    }
  }
  uint32_t progress = 0;
  size_t written = FILE_HEADER_BYTES;
  while (progress < num_samples) {
    const uint32_t n = (num_samples - progress < spb) ? num_samples - progress : spb;
    if (trials > 0) search_start_state(&e, pcm, ch_stride, progress, n);
    written += encode_block(&e, pcm, ch_stride, progress, n, out + written);
    progress += n;
  }
  for (uint32_t c = 0; c < channels; c++) {
    free(e.cur[c]);
    free(e.prev[c]);
    if (state) {
      memcpy(state[c].weight, e.chain[c].weight, sizeof(state[c].weight));
      state[c].stepsize_index = e.chain[c].index;
    }
  }
  return (int64_t)written;
}

/* ---- decoder ----------------------------------------------------------------------- */

static inline uint32_t byte_or_zero(const uint8_t *data, size_t size, size_t pos) { return pos < size ? data[pos] : 0u; }

/* src/aad_decoder.c:321-475; `want` = min(samples_per_block, room left in the buffer) */
static void decode_block(const struct aad_oracle_info *h, const uint8_t *data, size_t size, size_t pos,
                         int16_t *pcm, size_t ch_stride, uint32_t first, uint32_t want)
{
  const uint32_t C = h->channels, bits = h->bits;
  struct chain ch[AAD_ORACLE_MAX_CHANNELS];
  for (uint32_t c = 0; c < C; c++) {
    const uint32_t head = (byte_or_zero(data, size, pos) << 8) | byte_or_zero(data, size, pos + 1);
    pos += 2;
    ch[c].index = (int16_t)(head >> 4);
    const uint32_t shift = head & 0xF;
    for (int k = 0; k < TAPS; k++) {
      const uint32_t w = (byte_or_zero(data, size, pos) << 8) | byte_or_zero(data, size, pos + 1);
      const uint32_t s = (byte_or_zero(data, size, pos + 2) << 8) | byte_or_zero(data, size, pos + 3);
      pos += 4;
      ch[c].weight[k] = (int32_t)((uint32_t)(int32_t)(int16_t)w << shift);
      ch[c].hist[k] = (int16_t)s;
    }
    for (uint32_t i = 0; i < TAPS && i < want; i++) pcm[c * ch_stride + first + i] = (int16_t)ch[c].hist[TAPS - 1 - i];
  }
  const uint32_t gs = group_samples(bits), gb = group_bytes(bits);
  for (uint32_t i = TAPS; i < want; i += gs) {
    for (uint32_t c = 0; c < C; c++) {
      uint32_t packed = 0;
      for (uint32_t b = 0; b < gb; b++) packed = (packed << 8) | byte_or_zero(data, size, pos++);
      for (uint32_t j = 0; j < gs; j++) {
        const uint32_t code = (packed >> (bits * (gs - 1 - j))) & ((1u << bits) - 1u);
        const int32_t s = chain_decode_sample(&ch[c], code, bits);
        if (i + j < want) pcm[c * ch_stride + first + i + j] = (int16_t)s;
      }
    }
  }
  if (h->ms == 1) {                               /* src/aad_decoder.c:458-470 */
    for (uint32_t i = 0; i < want; i++) {
      const int32_t m = pcm[first + i], s = pcm[ch_stride + first + i];
      pcm[first + i] = (int16_t)clamp16(m + s);
      pcm[ch_stride + first + i] = (int16_t)clamp16(m - s);
    }
  }
}

int aad_oracle_decode(const uint8_t *data, size_t size, int16_t *pcm, size_t ch_stride, uint32_t buf_channels,
                      uint32_t buf_samples, struct aad_oracle_info *info_out)
{
  if (!data || !pcm) return AAD_ORACLE_INVALID_ARGUMENT;
  struct aad_oracle_info h;
  const int rc = aad_oracle_read_header(data, size, &h, 1);
  if (rc != AAD_ORACLE_OK) return rc;
  if (info_out) *info_out = h;
  if (buf_channels < h.channels || buf_samples < h.num_samples) return AAD_ORACLE_INSUFFICIENT_BUFFER;
  uint32_t progress = 0;
  size_t pos = FILE_HEADER_BYTES;
  while (progress < h.num_samples && pos < size) {
    const size_t avail = (size - pos < h.block_size) ? size - pos : h.block_size;
    if (avail < CHANNEL_HEADER_BYTES * h.channels) return AAD_ORACLE_INSUFFICIENT_DATA;
    const uint32_t room = buf_samples - progress;
    const uint32_t want = room < h.samples_per_block ? room : h.samples_per_block;
    decode_block(&h, data, size, pos, pcm, ch_stride, progress, want);
    pos += avail;
    progress += want;
  }
  return AAD_ORACLE_OK;
}

/* ---- batch loops for the CPU baseline ---------------------------------------------- */

int aad_oracle_encode_batch(const int16_t *pcm, size_t clip_stride, size_t ch_stride, uint32_t num_clips,
                            uint32_t channels, uint32_t num_samples, uint32_t sampling_rate, uint32_t bits,
                            uint32_t max_block_size, uint32_t ms, uint32_t trials, uint8_t *aad, size_t aad_stride,
                            uint32_t *out_sizes)
{
  for (uint32_t i = 0; i < num_clips; i++) {
    const int64_t n = aad_oracle_encode(pcm + (size_t)i * clip_stride, ch_stride, channels, num_samples, sampling_rate,
                                        bits, max_block_size, ms, trials, NULL, aad + (size_t)i * aad_stride, aad_stride);
    if (n < 0) return (int)-n;
    if (out_sizes) out_sizes[i] = (uint32_t)n;
  }
  return AAD_ORACLE_OK;
}

int aad_oracle_decode_batch(const uint8_t *aad, size_t aad_stride, const uint32_t *sizes, uint32_t num_clips,
                            int16_t *pcm, size_t clip_stride, size_t ch_stride, uint32_t buf_channels,
                            uint32_t buf_samples)
{
  for (uint32_t i = 0; i < num_clips; i++) {
    const int rc = aad_oracle_decode(aad + (size_t)i * aad_stride, sizes ? sizes[i] : aad_stride,
                                     pcm + (size_t)i * clip_stride, ch_stride, buf_channels, buf_samples, NULL);
    if (rc != AAD_ORACLE_OK) return rc;
  }
  return AAD_ORACLE_OK;
}
