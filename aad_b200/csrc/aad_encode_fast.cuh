/*
 * aad_encode_fast.cuh -- the production encoder kernel (included by aad_kernels.cu).
 *
 * One thread = one (stream, channel) chain, walking its blocks in order with the predictor
 * weights and step index in registers (the reference carries them in processor[],
 * src/aad_encoder.c:21,853-886).  Per block: start-state search (src/aad_encoder.c:470-562) as
 * 1 + 2*trials dry-run passes, then the emitting pass (src/aad_encoder.c:565-727).
 *
 * What makes it fast:
 *  - no integer divide: floor((|diff| << (b-2)) / step) is umulhi(|diff| << (b-1), M[step]) >> L[step]
 *    with a per-step magic pair, exact for every reachable operand (tools/gen_tables.py);
 *  - one 8-byte shared-memory load per sample fetches {step, L, M} from a table indexed directly
 *    by the Q4 step index (kept pre-multiplied by 8 in a register), the index update is one
 *    VIADDMNMX.RELU, the index-delta table is replicated per lane (no bank conflicts);
 *  - PCM is read with 8-byte vector loads (4 samples) one unit ahead of use, never per sample;
 *  - one warp per CTA when there are few chains, so every SM sub-partition gets a warp.
 */
#pragma once

namespace {

constexpr int kEncLutEntries = AADF_INDEX_MAX + 1;

struct EncShared {
  uint2 lut[kEncLutEntries];   /* x = (step << 16) | shift, y = magic;  indexed by stepsize_index */
};

/* 8 * (Q4 index delta) per magnitude code as a register-resident byte LUT: three PRMTs instead
 * of a shared-memory load on the step-index dependency chain (src/aad_tables.c:8-45). */
template <int BITS>
struct EncDelta {
  static __host__ __device__ constexpr int v(int k)
  {
    constexpr int d4[8] = AADK_DELTA4_INIT;
    constexpr int d3[4] = AADK_DELTA3_INIT;
    constexpr int d2[2] = AADK_DELTA2_INIT;
    return 8 * (BITS == 4 ? d4[k & 7] : (BITS == 3 ? d3[k & 3] : d2[k & 1]));
  }
  static __host__ __device__ constexpr uint32_t pack(int first, int shift)
  {
    return ((uint32_t)((v(first) >> shift) & 0xFF)) | ((uint32_t)((v(first + 1) >> shift) & 0xFF) << 8) |
           ((uint32_t)((v(first + 2) >> shift) & 0xFF) << 16) | ((uint32_t)((v(first + 3) >> shift) & 0xFF) << 24);
  }
  static __device__ __forceinline__ int32_t lookup(uint32_t mag)
  {
    constexpr uint32_t lo0 = pack(0, 0), lo1 = pack(4, 0), hi0 = pack(0, 8), hi1 = pack(4, 8);
    uint32_t lo, hi, r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(lo) : "r"(lo0), "r"(lo1), "r"(mag));
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(hi0), "r"(hi1), "r"(mag));
    asm("prmt.b32 %0, %1, %2, 0xCC40;" : "=r"(r) : "r"(lo), "r"(hi));   /* lo | hi << 8, sign extended */
    return (int32_t)r;
  }
};

__device__ const uint32_t g_step_magic[256] = AADK_STEP_MAGIC_INIT;
__device__ const uint8_t g_step_shift[256] = AADK_STEP_SHIFT_INIT;

template <int BITS>
__device__ __forceinline__ void enc_load_shared(EncShared &s)
{
  for (int i = threadIdx.x; i < kEncLutEntries; i += blockDim.x) {
    const int e = (i + 8) >> 4;
    s.lut[i] = make_uint2(((uint32_t)g_step_table[e] << 16) | g_step_shift[e], g_step_magic[e]);
  }
  __syncthreads();
}

struct EncChain {
  int32_t h0, h1, h2, h3;   /* h0 newest */
  int32_t w0, w1, w2, w3;
  int32_t idx8;             /* 8 * stepsize_index */
};

/* src/aad_encoder.c:343-410, one sample.  Returns the magnitude|sign code; q = signed dequantised diff. */
template <int BITS>
__device__ __forceinline__ uint32_t enc_sample(EncChain &c, int32_t x, const EncShared &s, int32_t &q)
{
  constexpr uint32_t kMaxMag = (1u << (BITS - 1)) - 1u;
  const uint2 e = *reinterpret_cast<const uint2 *>(reinterpret_cast<const char *>(s.lut) + c.idx8);
  const int32_t step = (int32_t)(e.x >> 16);
  const uint32_t acc = (1u << 14) + (uint32_t)c.h0 * (uint32_t)c.w0 + (uint32_t)c.h1 * (uint32_t)c.w1 +
                       (uint32_t)c.h2 * (uint32_t)c.w2 + (uint32_t)c.h3 * (uint32_t)c.w3;
  const int32_t p = (int32_t)acc >> 15;
  const int32_t d = (int32_t)((uint32_t)x - (uint32_t)p);
  const uint32_t a = (uint32_t)(d < 0 ? -d : d);
  uint32_t mag = __funnelshift_r(__umulhi(a << (BITS - 1), e.y), 0u, e.x);   /* >> (e.x & 31) */
  mag = min(mag, kMaxMag);
  const int32_t qa = (step * (int32_t)(2u * mag + 1u)) >> (BITS - 1);
  q = d < 0 ? -qa : qa;
  const int32_t r = max(__viaddmin_s32(q, p, 32767), -32768);
  c.w0 += (int32_t)((uint32_t)q * (uint32_t)c.h0 + (1u << 14)) >> 18;
  c.w1 += (int32_t)((uint32_t)q * (uint32_t)c.h1 + (1u << 14)) >> 18;
  c.w2 += (int32_t)((uint32_t)q * (uint32_t)c.h2 + (1u << 14)) >> 18;
  c.w3 += (int32_t)((uint32_t)q * (uint32_t)c.h3 + (1u << 14)) >> 18;
  c.idx8 = __viaddmin_s32_relu(c.idx8, EncDelta<BITS>::lookup(mag), 8 * AADF_INDEX_MAX);
  c.h3 = c.h2;
  c.h2 = c.h1;
  c.h1 = c.h0;
  c.h0 = r;
  return mag | (d < 0 ? (1u << (BITS - 1)) : 0u);
}

/* One channel's samples: plain row, or mid / side computed on the fly from the L and R rows
 * (src/aad_encoder.c:413-428; with int16 inputs the clip there can never trigger). */
template <int MS>
struct EncSource {
  const int16_t *a, *b;
  int mode;   /* 0 plain, 1 mid, 2 side */

  __device__ __forceinline__ int32_t combine(int32_t x, int32_t y) const
  {
    if (!MS || mode == 0) return x;
    return (x + (mode == 1 ? y : -y)) >> 1;
  }
  __device__ __forceinline__ int32_t at(uint32_t i) const { return combine(a[i], MS ? (int32_t)b[i] : 0); }
};

/* Four consecutive samples as loaded (two packed int16 per register); unpacked only when used,
 * so a whole unit can sit in registers while the previous one is being encoded. */
template <int MS>
struct EncQuad {
  uint2 va, vb;
  __device__ __forceinline__ void load(const EncSource<MS> &src, uint32_t i)   /* i % 4 == 0 */
  {
    va = __ldg(reinterpret_cast<const uint2 *>(src.a + i));
    if (MS) vb = (src.mode != 0) ? __ldg(reinterpret_cast<const uint2 *>(src.b + i)) : make_uint2(0u, 0u);
  }
  __device__ __forceinline__ int32_t get(const EncSource<MS> &src, int j) const
  {
    const uint32_t wa = (j < 2) ? va.x : va.y;
    const int32_t x = (j & 1) ? ((int32_t)wa >> 16) : (int32_t)(int16_t)(wa & 0xFFFFu);
    if (!MS) return x;
    const uint32_t wb = (j < 2) ? vb.x : vb.y;
    const int32_t y = (j & 1) ? ((int32_t)wb >> 16) : (int32_t)(int16_t)(wb & 0xFFFFu);
    return src.combine(x, y);
  }
};

constexpr int kEncUnit = 16;            /* samples per prefetched unit */
constexpr int kEncUnitQuads = kEncUnit / 4;

template <int MS>
__device__ __forceinline__ void enc_load_history(EncChain &c, const EncSource<MS> &src, uint32_t first, uint32_t n)
{
  if (n >= 4) {
    EncQuad<MS> hq;
    hq.load(src, first);
    c.h3 = hq.get(src, 0); c.h2 = hq.get(src, 1); c.h1 = hq.get(src, 2); c.h0 = hq.get(src, 3);
  } else {   /* a stream shorter than the filter: missing taps are zero (src/aad_encoder.c:608-615) */
    c.h3 = (n > 0) ? src.at(first) : 0;
    c.h2 = (n > 1) ? src.at(first + 1) : 0;
    c.h1 = (n > 2) ? src.at(first + 2) : 0;
    c.h0 = 0;
  }
}

/* src/aad_encoder.c:431-467: dry run over [first, first+n), returns sqrt(mean of wrapped q^2).
 * Not inlined more than once: the caller loops over passes. */
template <int BITS, int MS>
__device__ __forceinline__ double enc_trial_pass(EncChain &c, const EncSource<MS> &src, uint32_t first, uint32_t n,
                                                 const EncShared &s)
{
  if (n < AADF_TAPS) return 0.0;   /* state untouched */
  enc_load_history<MS>(c, src, first, n);
  long long sum = 0;
  uint32_t i = first + AADF_TAPS;
  const uint32_t end = first + n;
  const uint32_t units = (n - AADF_TAPS) / kEncUnit;
  EncQuad<MS> cur[kEncUnitQuads], nxt[kEncUnitQuads];
  if (units) {
#pragma unroll
    for (int k = 0; k < kEncUnitQuads; k++) cur[k].load(src, i + 4 * k);
  }
  for (uint32_t u = 0; u < units; u++) {
    if (u + 1 < units) {
#pragma unroll
      for (int k = 0; k < kEncUnitQuads; k++) nxt[k].load(src, i + kEncUnit + 4 * k);
    }
#pragma unroll
    for (int j = 0; j < kEncUnit; j++) {
      int32_t q;
      enc_sample<BITS>(c, cur[j >> 2].get(src, j & 3), s, q);
      sum += (long long)(int32_t)((uint32_t)q * (uint32_t)q);   /* 32-bit wrapping square, as compiled in the reference */
    }
#pragma unroll
    for (int k = 0; k < kEncUnitQuads; k++) cur[k] = nxt[k];
    i += kEncUnit;
  }
  for (; i + 4 <= end; i += 4) {
    EncQuad<MS> qd;
    qd.load(src, i);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int32_t q;
      enc_sample<BITS>(c, qd.get(src, j), s, q);
      sum += (long long)(int32_t)((uint32_t)q * (uint32_t)q);
    }
  }
  for (; i < end; i++) {
    int32_t q;
    enc_sample<BITS>(c, src.at(i), s, q);
    sum += (long long)(int32_t)((uint32_t)q * (uint32_t)q);
  }
  return sqrt((double)sum / (double)n);
}

/* Per-lane output stream for the mono case, where a chain's code bytes are contiguous: bytes are
 * collected in a 64-bit register and leave as aligned 4-byte stores (single bytes only up to the
 * first 4-byte boundary and for the last partial word): one store per 8 samples instead of four
 * lane-strided byte stores. */
struct EncByteStream {
  uint8_t *ptr;          /* next byte to write */
  unsigned long long acc;
  uint32_t cnt;          /* bytes held in acc */
  __device__ __forceinline__ void begin(uint8_t *p) { ptr = p; acc = 0ull; cnt = 0u; }
  __device__ __forceinline__ void put(uint32_t value, uint32_t nbytes)   /* value: first byte in bits 0-7 */
  {
    acc |= (unsigned long long)value << (8u * cnt);
    cnt += nbytes;
    while (cnt >= 4u) {
      if (((uintptr_t)ptr & 3u) == 0u) {
        *reinterpret_cast<uint32_t *>(ptr) = (uint32_t)acc;
        ptr += 4; acc >>= 32; cnt -= 4u;
      } else {
        *ptr++ = (uint8_t)acc; acc >>= 8; cnt -= 1u;
      }
    }
  }
  __device__ __forceinline__ void end()
  {
    while (cnt) { *ptr++ = (uint8_t)acc; acc >>= 8; cnt--; }
  }
};

template <int BITS>
__device__ __forceinline__ void enc_store_group(uint8_t *dp, uint32_t packed)
{
  if (BITS == 3) {
    dp[0] = (uint8_t)(packed >> 16); dp[1] = (uint8_t)(packed >> 8); dp[2] = (uint8_t)packed;
  } else {
    dp[0] = (uint8_t)packed;
  }
}

template <int BITS, int MS>
__global__ void __launch_bounds__(128) aad_encode_fast(const aadk_encode_params p)
{
  __shared__ EncShared sh;
  enc_load_shared<BITS>(sh);

  constexpr uint32_t GS = (BITS == 4) ? 2 : (BITS == 3 ? 8 : 4);
  constexpr uint32_t GB = (BITS == 3) ? 3 : 1;
  const uint32_t C = p.geo.channels;
  const uint32_t spb = p.geo.samples_per_block;
  const uint32_t bs = p.geo.block_size;

  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stream = t / C;
  if (stream >= p.num_streams) return;
  const uint32_t ch = (uint32_t)(t % C);
  const uint32_t ns = p.num_samples ? p.num_samples[stream] : p.uniform_samples;

  uint8_t *out = p.aad + stream * p.aad_stride;
  if (ch == 0 && p.block_begin == 0) {
    if (ns > 0) aadf_write_file_header(out, C, ns, p.sampling_rate, BITS, bs, spb, p.geo.ms);
    if (p.out_sizes) p.out_sizes[stream] = ns ? (uint32_t)aadf_stream_bytes(ns, C, BITS, bs, spb) : 0u;
  }

  EncSource<MS> src;
  {
    const int16_t *base = (const int16_t *)p.pcm + stream * p.pcm_clip_stride;
    const bool pair = MS && ch < 2;
    src.a = base + (uint64_t)(pair ? 0 : ch) * p.pcm_ch_stride;
    src.b = base + p.pcm_ch_stride;
    src.mode = pair ? (ch == 0 ? 1 : 2) : 0;
  }

  EncChain c;
  const uint64_t st = (stream * C + ch) * AADK_STATE_WORDS;
  c.h0 = c.h1 = c.h2 = c.h3 = 0;
  c.w0 = p.state_in ? p.state_in[st + 0] : 0;
  c.w1 = p.state_in ? p.state_in[st + 1] : 0;
  c.w2 = p.state_in ? p.state_in[st + 2] : 0;
  c.w3 = p.state_in ? p.state_in[st + 3] : 0;
  c.idx8 = 8 * (p.state_in ? p.state_in[st + 4] : 0);

  const uint32_t nblk = min(aadf_num_blocks(ns, spb), p.block_end);
  for (uint32_t b = p.block_begin; b < nblk; b++) {
    const uint32_t first = b * spb;
    const uint32_t n = min(spb, ns - first);
    const uint32_t limit = first + n;

    if (p.trials > 0) {
      /* src/aad_encoder.c:470-562 as one loop over dry-run passes (a single copy of the sample
       * loop in the instruction stream):
       *   pass 0        baseline: current block from the carried state
       *   pass 2t+1     trial t: previous block (skipped in the first block), state keeps running
       *   pass 2t+2     trial t: snapshot = candidate, current block, keep candidate if strictly better */
      EncChain run = c;
      int32_t bw0 = c.w0, bw1 = c.w1, bw2 = c.w2, bw3 = c.w3, bidx = c.idx8;   /* best so far = carried state */
      int32_t cw0 = c.w0, cw1 = c.w1, cw2 = c.w2, cw3 = c.w3, cidx = c.idx8;   /* candidate snapshot */
      double best = 0.0;
      const uint32_t passes = 1 + 2 * p.trials;
      for (uint32_t k = 0; k < passes; k++) {
        const bool on_prev = (k & 1u) != 0u;
        if (on_prev && b == 0) continue;
        if (!on_prev && k > 0) { cw0 = run.w0; cw1 = run.w1; cw2 = run.w2; cw3 = run.w3; cidx = run.idx8; }
        const double rmse = enc_trial_pass<BITS, MS>(run, src, on_prev ? first - spb : first, on_prev ? spb : n, sh);
        if (k == 0) {
          best = rmse;
          run = c;                 /* the trial chain restarts from the carried state */
        } else if (!on_prev && best > rmse) {   /* NaN compares false, like the reference */
          best = rmse;
          bw0 = cw0; bw1 = cw1; bw2 = cw2; bw3 = cw3; bidx = cidx;
        }
      }
      c.w0 = bw0; c.w1 = bw1; c.w2 = bw2; c.w3 = bw3; c.idx8 = bidx;
    }

    /* block header, src/aad_encoder.c:606-655 */
    enc_load_history<MS>(c, src, first, n);
    int32_t maxabs = 0;
    maxabs = max(maxabs, c.w0 >= 0 ? c.w0 : (int32_t)(0u - (uint32_t)c.w0));
    maxabs = max(maxabs, c.w1 >= 0 ? c.w1 : (int32_t)(0u - (uint32_t)c.w1));
    maxabs = max(maxabs, c.w2 >= 0 ? c.w2 : (int32_t)(0u - (uint32_t)c.w2));
    maxabs = max(maxabs, c.w3 >= 0 ? c.w3 : (int32_t)(0u - (uint32_t)c.w3));
    uint32_t shift = 0;
    while (maxabs > 32767) { maxabs >>= 1; shift++; }
    const int32_t keep = (int32_t)~((1u << shift) - 1u);
    c.w0 &= keep; c.w1 &= keep; c.w2 &= keep; c.w3 &= keep;
    uint8_t *blk = out + AADF_FILE_HEADER_BYTES + (uint64_t)b * bs;
    uint8_t *hp = blk + ch * AADF_CHANNEL_HEADER_BYTES;
    aadf_put_be16(hp, ((((uint32_t)c.idx8 >> 3) << 4) | (shift & 0xFu)) & 0xFFFFu);
    aadf_put_be16(hp + 2, (uint32_t)(c.w0 >> shift) & 0xFFFFu);  aadf_put_be16(hp + 4, (uint32_t)c.h0 & 0xFFFFu);
    aadf_put_be16(hp + 6, (uint32_t)(c.w1 >> shift) & 0xFFFFu);  aadf_put_be16(hp + 8, (uint32_t)c.h1 & 0xFFFFu);
    aadf_put_be16(hp + 10, (uint32_t)(c.w2 >> shift) & 0xFFFFu); aadf_put_be16(hp + 12, (uint32_t)c.h2 & 0xFFFFu);
    aadf_put_be16(hp + 14, (uint32_t)(c.w3 >> shift) & 0xFFFFu); aadf_put_be16(hp + 16, (uint32_t)c.h3 & 0xFFFFu);

    /* code groups, src/aad_encoder.c:661-722: full 16-sample units from prefetched vector loads ... */
    uint8_t *dp = blk + C * AADF_CHANNEL_HEADER_BYTES + ch * GB;
    const uint32_t gstride = C * GB;
    const bool mono = (C == 1);                      /* uniform: contiguous code bytes -> word stores */
    EncByteStream bytes;
    bytes.begin(dp);
    uint32_t i = first + AADF_TAPS;
    const uint32_t units = (n > AADF_TAPS) ? (n - AADF_TAPS) / kEncUnit : 0;
    EncQuad<MS> cur[kEncUnitQuads], nxt[kEncUnitQuads];
    if (units) {
#pragma unroll
      for (int k = 0; k < kEncUnitQuads; k++) cur[k].load(src, i + 4 * k);
    }
    for (uint32_t u = 0; u < units; u++) {
      if (u + 1 < units) {
#pragma unroll
        for (int k = 0; k < kEncUnitQuads; k++) nxt[k].load(src, i + kEncUnit + 4 * k);
      }
      /* the unit's codes, first sample in the most significant bits (16*BITS <= 64 bits) */
      unsigned long long codes = 0ull;
#pragma unroll
      for (int j = 0; j < kEncUnit; j++) {
        int32_t q;
        codes = (codes << BITS) | enc_sample<BITS>(c, cur[j >> 2].get(src, j & 3), sh, q);
      }
      constexpr int kUnitBytes = kEncUnit * BITS / 8;          /* 8 / 6 / 4 */
      if (mono) {
        /* stream order = most significant byte first: byte-reverse, then feed little-endian */
        const uint32_t hi = (uint32_t)(codes >> 32), lo = (uint32_t)codes;
        if (BITS == 4) {
          bytes.put(__byte_perm(hi, 0u, 0x0123), 4u);
          bytes.put(__byte_perm(lo, 0u, 0x0123), 4u);
        } else if (BITS == 3) {      /* 48 bits: bytes 5..0 of codes */
          bytes.put(__byte_perm(hi, lo, 0x6701), 4u);   /* hi.b1, hi.b0, lo.b3, lo.b2 */
          bytes.put(__byte_perm(lo, 0u, 0x4401), 2u);   /* lo.b1, lo.b0 */
        } else {
          bytes.put(__byte_perm(lo, 0u, 0x0123), 4u);
        }
      } else {
#pragma unroll
        for (int g = 0; g < kUnitBytes / (int)GB; g++) {
          const uint32_t packed = (uint32_t)(codes >> (8 * GB * (kUnitBytes / GB - 1 - g))) & ((1u << (8 * GB)) - 1u);
          enc_store_group<BITS>(dp, packed);
          dp += gstride;
        }
      }
#pragma unroll
      for (int k = 0; k < kEncUnitQuads; k++) cur[k] = nxt[k];
      i += kEncUnit;
    }
    /* ... then the tail in whole groups, zero padded past the end (src/aad_encoder.c:592-593) */
    for (; i < limit; i += GS) {
      uint32_t packed = 0;
#pragma unroll
      for (uint32_t j = 0; j < GS; j++) {
        int32_t q;
        const int32_t xs = (i + j < limit) ? src.at(i + j) : 0;
        packed = (packed << BITS) | enc_sample<BITS>(c, xs, sh, q);
      }
      if (mono) {
        if (BITS == 3) bytes.put(__byte_perm(packed, 0u, 0x4012), 3u);
        else bytes.put(packed, 1u);
      } else {
        enc_store_group<BITS>(dp, packed);
        dp += gstride;
      }
    }
    if (mono) bytes.end();
  }

  if (p.state_out) {
    p.state_out[st + 0] = c.w0;
    p.state_out[st + 1] = c.w1;
    p.state_out[st + 2] = c.w2;
    p.state_out[st + 3] = c.w3;
    p.state_out[st + 4] = c.idx8 >> 3;
  }
}

/* rows must be 8-byte aligned wherever a quad is loaded */
inline bool enc_fast_eligible(const aadk_encode_params &p)
{
  if (p.in32) return false;
  if (p.geo.samples_per_block % 4u) return false;
  if (((uintptr_t)p.pcm & 7u) || (p.pcm_clip_stride % 4u) || (p.pcm_ch_stride % 4u)) return false;
  return true;
}

template <int BITS>
int enc_fast_launch(const aadk_encode_params &p, cudaStream_t s)
{
  const uint64_t chains = (uint64_t)p.num_streams * p.geo.channels;
  /* few chains: one warp per CTA so the warps spread over all SMs / sub-partitions */
  const unsigned block = (chains <= 148ull * 6 * 32) ? 32 : 128;
  const unsigned grid = (unsigned)((chains + block - 1) / block);
  if (p.geo.ms && p.geo.channels >= 2) aad_encode_fast<BITS, 1><<<grid, block, 0, s>>>(p);
  else aad_encode_fast<BITS, 0><<<grid, block, 0, s>>>(p);
  return (int)cudaGetLastError();
}

}  // namespace
