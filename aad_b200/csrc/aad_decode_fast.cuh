/*
 * aad_decode_fast.cuh -- the production decoder kernels (included by aad_kernels.cu).
 *
 * One thread = one (block, channel) chain (every block header reloads the whole chain state,
 * src/aad_decoder.c:364-380).  A warp task = 32/C consecutive blocks of the batch's (stream, block) sequence
 * (aad_decode_fast: a task may end in the next stream, so short streams leave no lane idle; aad_decode_wide: blocks of
 * one stream), walked in windows of 128 samples per chain:
 *
 *   global .aad  --16-byte coalesced loads (aligned superset, prefetched one window ahead)-->
 *   shared input rows --per-lane word reads + funnel shift (any byte alignment)--> sample chain
 *   --8-byte stores--> shared output rows --coalesced 8-byte row stores--> global PCM
 *
 * so no global access is ever lane-strided (the generic kernel's byte loads / 2-byte stores touch
 * 32 lines per instruction).  Windows are cut on input bytes (TB = 16*bits*C per block), chosen so
 * that a window never splits a code group and the block header (18*C bytes) plus, for mono, one
 * half step bring the read pointer back to word alignment.
 *
 * The grid is persistent: one wave of 16-warp CTAs (one per SM), every warp looping over its share of
 * the warp tasks, so the shared-memory tables are built once per resident CTA.
 *
 * aad_decode_fast<BITS, C>  mono / stereo, sliding 32-bit words;
 * aad_decode_wide<BITS>     any channel count (run time), byte reads from the shared input rows.
 */
#pragma once

namespace {

#ifndef AAD_DEC_WARPS
#define AAD_DEC_WARPS 16
#endif
constexpr int kDecWarps = AAD_DEC_WARPS;
constexpr int kDecWindow = 128;           /* samples per chain per window */
constexpr int kDecOutPitch = 264;         /* 256 + 8: conflict-free 8-byte shared accesses */
constexpr int kDecBulkPitch = 272;        /* BULK: 16-byte aligned rows (cp.async.bulk), conflict-free 16-byte accesses */


/* Dequantised difference per (step row, code) and index delta per magnitude, both in shared memory.
 * The row of src/aad_tables.h:41 ((index + 8) >> 4) starts at byte 4 * (index + 8) & ~63 of q, and a code
 * shifted to bits 2..5 selects the entry: one multiply-add, one LOP3 and one LDS return
 * +-((step * (2 mag + 1)) >> (bits - 1)) (src/aad_decoder.c:284-296) -- no step load, multiply, shift,
 * sign test or negate per sample.  Rows hold 16 entries for every bit depth (entry e = code e mod 2^bits),
 * so the bits above a narrower code need no masking.  The 8 deltas sit in 8 different banks: lanes
 * either read the same word (broadcast) or different banks, never a conflict.
 * (Measured alternatives -- step table + arithmetic, 511 half-step rows, CTA shapes: profiles/r01_v7_decode.md.) */
constexpr int kDecRows = 256;
struct DecTables {
  int32_t q[kDecRows][16];
  int32_t delta[8];                       /* index delta per magnitude code */
};

template <int BITS, int C>
struct DecGeom {
  static constexpr int GB = (BITS == 3) ? 3 : 1;
  static constexpr int TB = 16 * BITS * C;                 /* input bytes per block per window */
  static constexpr int IN_CHUNKS = TB / 16 + 1;            /* aligned superset */
  /* row pitch: an ODD number of 16-byte chunks, so the rows of a warp start in 8 different banks (4 words apart);
   * with an even count -- 4 chunks for mono 3-bit -- 32 rows would share 2 banks and every per-lane word read of
   * the sample loop would be a 16-way conflict */
  static constexpr int IN_PITCH = 16 * (IN_CHUNKS | 1);
  static constexpr int IN_ROWS = 32 / C;
  static constexpr int IN_LOADS = (IN_ROWS * IN_CHUNKS + 31) / 32;
  static constexpr int IN_BYTES = IN_ROWS * IN_PITCH + 16; /* + slack for the funnel look-ahead word */
  static constexpr int STEP_BYTES = (BITS == 3) ? 12 : 4;
  static constexpr int SPS = STEP_BYTES * 8 / (BITS * C);  /* samples per chain per step */
  static constexpr int HALF_BYTES = STEP_BYTES / 2;        /* mono only: 18-byte header -> word alignment */
  static constexpr int WARP_BYTES = ((IN_BYTES + 15) & ~15) + 32 * kDecOutPitch;
  static constexpr int WARP_BYTES_BULK = ((IN_BYTES + 15) & ~15) + 32 * kDecBulkPitch;
};

template <int BITS>
__device__ __forceinline__ void dec_load_tables(DecTables &t)
{
  for (int i = threadIdx.x; i < kDecRows * 16; i += blockDim.x) {
    const int32_t step = g_step_table[i >> 4];
    const int code = i & ((1 << BITS) - 1);
    const int mag = code & ((1 << (BITS - 1)) - 1);
    const int32_t qa = (step * (2 * mag + 1)) >> (BITS - 1);
    t.q[i >> 4][i & 15] = (code >> (BITS - 1)) ? -qa : qa;
  }
  /* dec_sample merges code bits into table addresses: the tables' own address bits there must be zero */
  if (threadIdx.x == 0 && (((uint32_t)__cvta_generic_to_shared(t.q) & 63u) || ((uint32_t)__cvta_generic_to_shared(t.delta) & 31u))) __trap();
  if (threadIdx.x < 8) {
    const int k = threadIdx.x;
    int d = 0;
    if (BITS == 4) d = g_delta4[k];
    if (BITS == 3) d = g_delta3[k & 3];
    if (BITS == 2) d = g_delta2[k & 1];
    t.delta[k] = d;
  }
  __syncthreads();
}

__constant__ uint32_t g_dec_four = 4u;
#ifndef AAD_DEC_BIASCLIP
#define AAD_DEC_BIASCLIP 0
#endif
__constant__ int32_t g_dec_one = 1;   /* AAD_DEC_BIASCLIP: an opaque 1, so that x * 1 + c stays a multiply-add */

/* ADDR = 1 (the stream-spanning kernels): the chain carries complete shared-space table addresses, see dec_sample */
template <int ADDR>
struct DecChainT {
  int32_t h0, h1, h2, h3;
  int32_t w0, w1, w2, w3;
  int32_t idx;    /* stepsize_index */
  uint32_t four;  /* 4, opaque to the compiler (g_dec_four) */
  uint32_t qrow;  /* ADDR: shared-space address of DecTables::q (a multiple of 64) + 32: the addend of the row multiply-add */
  uint32_t dtab;  /* ADDR: shared-space address of DecTables::delta (a multiple of 32) */
  __device__ __forceinline__ void tables(const DecTables &t)
  {
    if (ADDR) {
      /* through volatile moves: ptxas must keep the three in registers instead of re-deriving them (constant load,
       * SR_CgaCtaId read: variable-latency instructions) inside the sample loop */
      asm volatile("mov.u32 %0, %1;" : "=r"(four) : "r"(g_dec_four));
      asm volatile("mov.u32 %0, %1;" : "=r"(qrow) : "r"((uint32_t)__cvta_generic_to_shared(t.q) + 32u));
      asm volatile("mov.u32 %0, %1;" : "=r"(dtab) : "r"((uint32_t)__cvta_generic_to_shared(t.delta)));
    } else {
      four = g_dec_four;
      qrow = dtab = 0u;
    }
  }
};
typedef DecChainT<0> DecChain;

/* src/aad_decoder.c:269-318 for the code whose least significant bit sits at bit POS of v. */
template <int BITS, int POS, int ADDR>
__device__ __forceinline__ int32_t dec_sample(DecChainT<ADDR> &c, uint32_t v, const DecTables &t)
{
  constexpr uint32_t kMagField = ((1u << (BITS - 1)) - 1u) << 2;
  const uint32_t x = (POS >= 2) ? (v >> (POS >= 2 ? POS - 2 : 0)) : (v << (POS >= 2 ? 0 : 2 - POS));   /* code at bits 2.. */
  /* 4 * (index + 8) on the multiply pipe (the ALU pipe is the one this kernel saturates): the factor comes from
   * constant memory, so ptxas cannot turn the multiply-add into an ALU-pipe LEA */
  int32_t q, d;
  if (ADDR) {
    /* Both lookups go through complete shared-space addresses formed by instructions the sample needs anyway (the table
     * bases ride in the multiply-add's addend and in the mask operation): LDS [R].  With generic pointers ptxas re-added
     * the table base per lookup in these kernels (2 more instructions per sample). */
    uint32_t row, qoff, doff;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(row) : "r"(c.idx), "r"(c.four), "r"(c.qrow));
    asm("lop3.b32 %0, %1, %2, 0x3C, 0xD8;" : "=r"(qoff) : "r"(row), "r"(x));             /* (row & ~0x3C) | (x & 0x3C) */
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(doff) : "r"(x), "n"(kMagField), "r"(c.dtab));   /* (x & kMagField) | dtab */
    asm("ld.shared.s32 %0, [%1];" : "=r"(q) : "r"(qoff));
    asm("ld.shared.s32 %0, [%1];" : "=r"(d) : "r"(doff));
  } else {
    uint32_t row;
    asm("mad.lo.u32 %0, %1, %2, 32;" : "=r"(row) : "r"(c.idx), "r"(c.four));
    uint32_t qoff;   /* (row & ~0x3C) | (x & 0x3C); bits 0-1 of row are zero */
    asm("lop3.b32 %0, %1, %2, 0x3C, 0xD8;" : "=r"(qoff) : "r"(row), "r"(x));
    q = *reinterpret_cast<const int32_t *>(reinterpret_cast<const char *>(t.q) + qoff);
    d = *reinterpret_cast<const int32_t *>(reinterpret_cast<const char *>(t.delta) + (x & kMagField));
  }
  const uint32_t acc = (1u << 14) + (uint32_t)c.h0 * (uint32_t)c.w0 + (uint32_t)c.h1 * (uint32_t)c.w1 +
                       (uint32_t)c.h2 * (uint32_t)c.w2 + (uint32_t)c.h3 * (uint32_t)c.w3;
  const int32_t p = (int32_t)acc >> 15;
#if AAD_DEC_BIASCLIP
  /* experiment (profiles/r02_decoder_experiments.md 6): the two-sided 16-bit clip as ONE ALU-pipe instruction on biased
   * values -- relu(min(p + q + 32768, 65535)) - 32768 -- with the two bias adds on the multiply pipe (opaque factor 1) */
  int32_t qb, r;
  asm("mad.lo.s32 %0, %1, %2, 32768;" : "=r"(qb) : "r"(q), "r"(g_dec_one));
  const int32_t rb = __viaddmin_s32_relu(p, qb, 65535);
  asm("mad.lo.s32 %0, %1, %2, -32768;" : "=r"(r) : "r"(rb), "r"(g_dec_one));
#else
  const int32_t r = max(__viaddmin_s32(q, p, 32767), -32768);
#endif
  c.idx = __viaddmin_s32_relu(c.idx, d, AADF_INDEX_MAX);
  c.w0 += (int32_t)((uint32_t)q * (uint32_t)c.h0 + (1u << 14)) >> 18;
  c.w1 += (int32_t)((uint32_t)q * (uint32_t)c.h1 + (1u << 14)) >> 18;
  c.w2 += (int32_t)((uint32_t)q * (uint32_t)c.h2 + (1u << 14)) >> 18;
  c.w3 += (int32_t)((uint32_t)q * (uint32_t)c.h3 + (1u << 14)) >> 18;
  c.h3 = c.h2;
  c.h2 = c.h1;
  c.h1 = c.h0;
  c.h0 = r;
  return r;
}

__device__ __forceinline__ uint32_t dec_pack2(int32_t a, int32_t b) { return __byte_perm((uint32_t)a, (uint32_t)b, 0x5410); }

/* All codes of one byte (4-bit: 2, 2-bit: 4), byte at bits [8*K, 8*K+8) of v; samples appended to o[]. */
template <int BITS, int K, class Chain>
__device__ __forceinline__ void dec_byte(Chain &c, uint32_t v, const DecTables &t, int32_t *o)
{
  if (BITS == 4) {
    o[0] = dec_sample<4, 8 * K + 4>(c, v, t);
    o[1] = dec_sample<4, 8 * K>(c, v, t);
  } else {
    o[0] = dec_sample<2, 8 * K + 6>(c, v, t);
    o[1] = dec_sample<2, 8 * K + 4>(c, v, t);
    o[2] = dec_sample<2, 8 * K + 2>(c, v, t);
    o[3] = dec_sample<2, 8 * K>(c, v, t);
  }
}

/* the 8 codes of one 3-bit group held big-endian in the low 24 bits of g */
template <class Chain>
__device__ __forceinline__ void dec_group3(Chain &c, uint32_t g, const DecTables &t, int32_t *o)
{
  o[0] = dec_sample<3, 21>(c, g, t);
  o[1] = dec_sample<3, 18>(c, g, t);
  o[2] = dec_sample<3, 15>(c, g, t);
  o[3] = dec_sample<3, 12>(c, g, t);
  o[4] = dec_sample<3, 9>(c, g, t);
  o[5] = dec_sample<3, 6>(c, g, t);
  o[6] = dec_sample<3, 3>(c, g, t);
  o[7] = dec_sample<3, 0>(c, g, t);
}

/* store N (multiple of 4) samples to the lane's shared output row at sample offset `at` */
template <int N>
__device__ __forceinline__ void dec_emit(unsigned char *orow, uint32_t at, const int32_t *o)
{
#pragma unroll
  for (int k = 0; k < N; k += 4)
    *reinterpret_cast<uint2 *>(orow + 2 * (at + k)) = make_uint2(dec_pack2(o[k], o[k + 1]), dec_pack2(o[k + 2], o[k + 3]));
}

/* BULK: 8 samples as one 16-byte store (at % 8 == 0, rows 16-byte aligned) */
__device__ __forceinline__ void dec_emit8(unsigned char *orow, uint32_t at, const int32_t *o)
{
  *reinterpret_cast<uint4 *>(orow + 2 * at) =
      make_uint4(dec_pack2(o[0], o[1]), dec_pack2(o[2], o[3]), dec_pack2(o[4], o[5]), dec_pack2(o[6], o[7]));
}

/* BULK flush: `bytes` (a multiple of 16) of this lane's shared output row to its global row through the TMA unit
 * (cp.async.bulk, shared -> global).  The lane wrote the row itself: a proxy fence orders its stores before the
 * asynchronous read, no warp synchronisation is involved. */
__device__ __forceinline__ void dec_bulk_store(void *gdst, const unsigned char *srow, uint32_t bytes)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               :: "l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(srow)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dec_bulk_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void dec_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
/* the bulk reads of this thread's earlier groups have left shared memory: the row may be written again */
__device__ __forceinline__ void dec_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

/* 16 bytes at q (16-byte aligned) for the shared input rows; bytes at or past `end` -- the end of the stream's
 * valid data -- read as zero, a null q (a loader lane without a chunk) as well */
__device__ __forceinline__ uint4 dec_fetch16(const uint8_t *q, const uint8_t *end)
{
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (q == nullptr || q >= end) return v;
  if (q + 16 <= end) return __ldg(reinterpret_cast<const uint4 *>(q));
  unsigned char tmp[16];                      /* the chunk that straddles the end of the data */
#pragma unroll
  for (int k = 0; k < 16; k++) tmp[k] = (q + k < end) ? q[k] : (unsigned char)0;
  v.x = tmp[0] | (tmp[1] << 8) | (tmp[2] << 16) | ((uint32_t)tmp[3] << 24);
  v.y = tmp[4] | (tmp[5] << 8) | (tmp[6] << 16) | ((uint32_t)tmp[7] << 24);
  v.z = tmp[8] | (tmp[9] << 8) | (tmp[10] << 16) | ((uint32_t)tmp[11] << 24);
  v.w = tmp[12] | (tmp[13] << 8) | (tmp[14] << 16) | ((uint32_t)tmp[15] << 24);
  return v;
}

/* Flush of a window when the chains of the warp do not all deliver a whole block (a stream's last blocks,
 * short output buffers, missing data): row rr of the shared output rows holds `produced` samples of the chain
 * in lane rr, of which those below that chain's n_row go to its global row, as one coalesced run of 8-byte
 * pieces plus a scalar tail. */
/* mid / side -> left (M + S) or right (M - S), clipped to 16 bits (src/aad_decoder.c:458-470), on two packed samples */
__device__ __forceinline__ uint32_t dec_ms2(uint32_t m, uint32_t s, bool right) { return right ? __vsubss2(m, s) : __vaddss2(m, s); }

/* ms_pairs: rows 2k / 2k+1 hold the mid / side chains of one block and go out as its left / right channel */
/* lanes [a, b) as a mask (a <= b <= 32) */
__device__ __forceinline__ uint32_t dec_lanes(uint32_t a, uint32_t b)
{
  return (b < 32u ? (1u << b) - 1u : 0xFFFFFFFFu) & ~(a < 32u ? (1u << a) - 1u : 0xFFFFFFFFu);
}

__device__ __forceinline__ void dec_flush_ragged(const unsigned char *out_rows, uint32_t todo, uint32_t n_row,
                                                 uint32_t produced, int16_t *grow, uint32_t out_base, uint32_t lane,
                                                 bool ms_pairs = false)
{
  /* the rows of `todo` that still deliver samples in this window */
  uint32_t live = __ballot_sync(0xFFFFFFFFu, n_row > out_base) & todo;
  for (; live != 0u; live &= live - 1u) {
    const uint32_t rr = (uint32_t)__ffs((int)live) - 1u;
    const uint32_t n_rr = __shfl_sync(0xFFFFFFFFu, n_row, rr);
    const uint32_t made = __shfl_sync(0xFFFFFFFFu, produced, rr);
    const uint64_t gp = __shfl_sync(0xFFFFFFFFu, (unsigned long long)(uintptr_t)grow, rr);
    const uint32_t count = min(made, n_rr - out_base);
    int16_t *dst = reinterpret_cast<int16_t *>((uintptr_t)gp) + out_base;
    const unsigned char *srow = out_rows + rr * kDecOutPitch;
    const uint32_t s0 = lane * 4u;
    if (ms_pairs) {
      const unsigned char *mrow = out_rows + (rr & ~1u) * kDecOutPitch, *trow = out_rows + (rr | 1u) * kDecOutPitch;
      const bool right = (rr & 1u) != 0u;
      if (s0 + 4u <= count) {
        const uint2 m = *reinterpret_cast<const uint2 *>(mrow + 2u * s0), t = *reinterpret_cast<const uint2 *>(trow + 2u * s0);
        *reinterpret_cast<uint2 *>(dst + s0) = make_uint2(dec_ms2(m.x, t.x, right), dec_ms2(m.y, t.y, right));
      } else {
        for (uint32_t k = s0; k < count; k++) {
          const int32_t m = *reinterpret_cast<const int16_t *>(mrow + 2u * k), t = *reinterpret_cast<const int16_t *>(trow + 2u * k);
          dst[k] = (int16_t)max(-32768, min(32767, right ? m - t : m + t));
        }
      }
    } else if (s0 + 4u <= count) {
      *reinterpret_cast<uint2 *>(dst + s0) = *reinterpret_cast<const uint2 *>(srow + 2u * s0);
    } else {
      for (uint32_t k = s0; k < count; k++) dst[k] = *reinterpret_cast<const int16_t *>(srow + 2u * k);
    }
  }
}

/* the whole-block rows [0, nrows) of a stereo mid/side warp task: rows 2k / 2k+1 -> left / right plane of block k */
__device__ __forceinline__ void dec_flush_ms_rows(const unsigned char *srow, int16_t *dst, uint32_t nrows, uint64_t ch_stride,
                                                  uint32_t spb)
{
  for (uint32_t rr = 0; rr < nrows; rr += 2, dst += spb) {
    const uint2 m = *reinterpret_cast<const uint2 *>(srow + rr * kDecOutPitch);
    const uint2 t = *reinterpret_cast<const uint2 *>(srow + (rr + 1u) * kDecOutPitch);
    *reinterpret_cast<uint2 *>(dst) = make_uint2(dec_ms2(m.x, t.x, false), dec_ms2(m.y, t.y, false));
    *reinterpret_cast<uint2 *>(dst + ch_stride) = make_uint2(dec_ms2(m.x, t.x, true), dec_ms2(m.y, t.y, true));
  }
}

/* ---- WAV-order output (aadk_decode_params::interleaved): the flush forms whole frames ---------------------- */

/* Four consecutive frames of one block: rows srow, srow + pitch, ... hold the CH channels of the block (8 bytes =
 * this lane's 4 samples of each), dst = the first of the lane's 4 frames (4 * CH * 2 bytes, 16-byte aligned for
 * CH >= 2).  ms: rows 0 / 1 are mid / side (src/aad_decoder.c:458-470). */
template <int CH>
__device__ __forceinline__ void dec_flush_frames4(const unsigned char *srow, int16_t *dst, bool ms)
{
  uint2 m[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) m[c] = *reinterpret_cast<const uint2 *>(srow + c * kDecOutPitch);
  if (ms) {
    const uint2 mid = m[0], side = m[1];
    m[0] = make_uint2(dec_ms2(mid.x, side.x, false), dec_ms2(mid.y, side.y, false));
    m[1] = make_uint2(dec_ms2(mid.x, side.x, true), dec_ms2(mid.y, side.y, true));
  }
  /* 16 bytes at a time: 8 / CH frames of CH / 2 channel pairs each (few live registers: 8 channels would otherwise
   * hold 32) */
  constexpr int FPG = 8 / CH;
#pragma unroll
  for (int g = 0; g < 4 / FPG; g++) {
    uint32_t w[4];
#pragma unroll
    for (int ff = 0; ff < FPG; ff++) {
      const int f = g * FPG + ff;
#pragma unroll
      for (int j = 0; j < CH / 2; j++) {
        const uint32_t a = (f < 2) ? m[2 * j].x : m[2 * j].y, b = (f < 2) ? m[2 * j + 1].x : m[2 * j + 1].y;
        w[ff * (CH / 2) + j] = __byte_perm(a, b, (f & 1) ? 0x7632 : 0x5410);
      }
    }
    *reinterpret_cast<uint4 *>(dst + 8 * g) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

/* the same for blocks whose chains do not all deliver a whole block (see dec_flush_ragged): the rows of `todo`,
 * CH rows per block, every row of a block with the same count */
template <int CH>
__device__ __noinline__ void dec_flush_ragged_frames(const unsigned char *out_rows, uint32_t todo,
                                                        uint32_t n_row, uint32_t produced, int16_t *gframe,
                                                        uint32_t out_base, uint32_t lane, bool ms)
{
  uint32_t heads = 0u;   /* lanes that hold channel 0 of a block */
#pragma unroll
  for (int k = 0; k < 32; k += CH) heads |= 1u << k;
  uint32_t live = __ballot_sync(0xFFFFFFFFu, n_row > out_base) & heads & todo;
  for (; live != 0u; live &= live - 1u) {
    const uint32_t rr = (uint32_t)__ffs((int)live) - 1u;
    const uint32_t n_rr = __shfl_sync(0xFFFFFFFFu, n_row, rr);
    const uint32_t made = __shfl_sync(0xFFFFFFFFu, produced, rr);
    const uint64_t gp = __shfl_sync(0xFFFFFFFFu, (unsigned long long)(uintptr_t)gframe, rr);
    const uint32_t count = min(made, n_rr - out_base);
    int16_t *dst = reinterpret_cast<int16_t *>((uintptr_t)gp) + (uint64_t)out_base * CH;
    const unsigned char *srow = out_rows + rr * kDecOutPitch;
    const uint32_t s0 = lane * 4u;
    if (s0 + 4u <= count) {
      dec_flush_frames4<CH>(srow + 2u * s0, dst + (uint64_t)s0 * CH, ms);
    } else {
      for (uint32_t k = s0; k < count; k++) {
        if (ms) {
          const int32_t m = *reinterpret_cast<const int16_t *>(srow + 2u * k);
          const int32_t t = *reinterpret_cast<const int16_t *>(srow + kDecOutPitch + 2u * k);
          dst[(uint64_t)k * CH] = (int16_t)max(-32768, min(32767, m + t));
          dst[(uint64_t)k * CH + 1] = (int16_t)max(-32768, min(32767, m - t));
        } else {
#pragma unroll
          for (int c = 0; c < CH; c++) dst[(uint64_t)k * CH + c] = *reinterpret_cast<const int16_t *>(srow + c * kDecOutPitch + 2u * k);
        }
      }
    }
  }
}

/* IL = 1: stereo output in WAV order (frames), formed in the flush.
 * BULK = 1 (mono 4-bit, where every window of a block is a whole number of 16-byte pieces at a 16-byte aligned place
 * of the PCM row): every lane hands its own shared output row to the TMA unit (cp.async.bulk shared -> global) instead
 * of the warp copying 32 rows through registers. */
/* SPAN = 1: warp tasks are runs of consecutive blocks of the whole batch and may cross from one stream into the next
 * (streams whose block count is far from a multiple of 32 / C would otherwise idle lanes in their last task: 14 % of
 * them for the 110 / 165 blocks of a 10-second mono clip at 2 / 3 bits, a third for 1-second clips).  SPAN = 0: tasks
 * are blocks of one stream; everything about the stream is warp-uniform (the launcher's choice when little is wasted). */
template <int BITS, int C, int IL, int BULK = 0, int SPAN = 0>
__global__ void __launch_bounds__(kDecWarps * 32) aad_decode_fast(const aadk_decode_params p)
{
  static_assert(!BULK || (BITS == 4 && C == 1 && IL == 0 && SPAN == 0), "bulk-store flush: mono 4-bit planar");
  using G = DecGeom<BITS, C>;
  extern __shared__ __align__(128) unsigned char dec_smem[];
  DecTables &tab = *reinterpret_cast<DecTables *>(dec_smem);
  dec_load_tables<BITS>(tab);

  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = threadIdx.x >> 5;
  unsigned char *in_rows = dec_smem + ((sizeof(DecTables) + 15) & ~(size_t)15) +
                           (size_t)warp * (BULK ? G::WARP_BYTES_BULK : G::WARP_BYTES);
  unsigned char *out_rows = in_rows + ((G::IN_BYTES + 15) & ~15);

  const uint32_t spb = p.geo.samples_per_block;
  const uint32_t bs = p.geo.block_size;
  const uint32_t nblocks = p.block_end - p.block_begin;
  /* SPAN: a warp task = IN_ROWS consecutive blocks of the batch's (stream, block) sequence: it may begin in one stream
   * and end in the next (or span several short ones).  `single` tasks (all rows in one stream: most of them) keep
   * arithmetic row addresses throughout. */
  const uint32_t warps_per_stream = (nblocks + G::IN_ROWS - 1) / G::IN_ROWS;
  const uint64_t total_warps = SPAN ? ((uint64_t)p.num_streams * nblocks + G::IN_ROWS - 1) / G::IN_ROWS
                                    : (uint64_t)p.num_streams * warps_per_stream;
  const bool ms2 = (C == 2) && p.geo.ms != 0u;   /* mid/side stereo: left / right formed in the flush */
  /* persistent: the warps of the grid share out the warp tasks round robin (the tables are built once per CTA) */
  for (uint64_t gw = (uint64_t)blockIdx.x * kDecWarps + warp; gw < total_warps; gw += (uint64_t)gridDim.x * kDecWarps) {
    uint64_t stream0;
    uint32_t bo0;                                    /* row 0's block, counted from block_begin */
    if (SPAN) {
      const uint64_t row0 = gw * G::IN_ROWS;
      stream0 = row0 / nblocks;
      bo0 = (uint32_t)(row0 - stream0 * nblocks);
    } else {
      stream0 = gw / warps_per_stream;
      bo0 = (uint32_t)(gw % warps_per_stream) * G::IN_ROWS;
    }
    const bool single = !SPAN || bo0 + G::IN_ROWS <= nblocks;

    /* this lane's chain */
    const uint32_t row = lane / C, ch = lane % C;
    uint64_t stream = stream0;
    uint32_t bo = bo0 + row;
    if (!single) {
      const uint32_t over = bo / nblocks;
      stream += over;
      bo -= over * nblocks;
    }
    const bool in_batch = !SPAN || stream < p.num_streams;    /* SPAN: the batch's last task may run past its last stream */
    if (!in_batch) stream = p.num_streams - 1u;
    const uint32_t b = p.block_begin + bo;

    const uint8_t *slot = p.aad + stream * p.aad_stride;
    const uint32_t size = p.sizes ? p.sizes[stream] : p.uniform_size;
    uint32_t ns = p.uniform_samples;
    if (p.read_headers) ns = dec_header_samples(slot, size, p.uniform_samples);
    const uint32_t buf = p.buf_samples ? p.buf_samples : ns;

    const uint64_t blk_off = AADF_FILE_HEADER_BYTES + (uint64_t)b * bs;
    const bool have = in_batch && b < p.block_end && (uint64_t)b * spb < ns &&
                      blk_off + (uint64_t)AADF_CHANNEL_HEADER_BYTES * C <= size;
    /* samples this chain delivers: what the output buffer still holds at the block's first sample */
    const uint32_t n_row = (have && (uint64_t)b * spb < buf) ? min(spb, buf - b * spb) : 0u;
    /* rows of p.pcm start at sample p.sample_base, p.aad at byte p.byte_base of the stream (shards of one stream) */
    const uint64_t srel = (uint64_t)b * spb - p.sample_base;
    /* this chain's row (IL: first frame of the block) */
    int16_t *grow = IL ? (int16_t *)p.pcm + stream * p.pcm_clip_stride + srel * C
                       : (int16_t *)p.pcm + stream * p.pcm_clip_stride + (uint64_t)ch * p.pcm_ch_stride + srel;
    /* Arithmetic-address segments of the flush: lanes [0, nfull) -- rows of stream0 from row 0 of the task that deliver
     * whole blocks, base grow0 -- and, in a task that goes on into the next stream, lanes [seg1, seg1 + nfull1) with base
     * grow1.  Everything else (a stream's last, partial block; rows of a third stream; absent blocks) is flushed row by
     * row (dec_flush_ragged).  All of these are multiples of C: the channels of a block share its sample count. */
    const uint32_t full_mask = __ballot_sync(0xFFFFFFFFu, n_row == spb);
    const uint32_t seg1 = single ? 32u : (nblocks - bo0) * C;                     /* first lane of the next stream */
    const uint32_t nfull = min((uint32_t)__ffs((int)~full_mask) - 1u, seg1);       /* __ffs(0) - 1 wraps to all ones */
    const bool all_full = single && nfull >= 32u;
    uint32_t nfull1 = 0u;
    if (!single) nfull1 = min(min((uint32_t)__ffs((int)~(full_mask >> seg1)) - 1u, 32u - seg1), nblocks * C);
    int16_t *grow0, *grow1;
    if (SPAN) {
      grow0 = (int16_t *)(uintptr_t)__shfl_sync(0xFFFFFFFFu, (unsigned long long)(uintptr_t)grow, 0);
      grow1 = (int16_t *)(uintptr_t)__shfl_sync(0xFFFFFFFFu, (unsigned long long)(uintptr_t)grow, seg1 & 31u);
    } else {
      const uint64_t srel0 = (uint64_t)(p.block_begin + bo0) * spb - p.sample_base;
      grow0 = grow1 = (int16_t *)p.pcm + stream0 * p.pcm_clip_stride + (IL ? srel0 * C : srel0);
    }
    const uint32_t ragged = ~(dec_lanes(0u, nfull) | dec_lanes(seg1, seg1 + nfull1));   /* lanes flushed row by row */

    /* loader role: IN_LOADS 16-byte chunks per lane per window, the aligned superset of every row's window.
     * A task whose rows all keep their windows inside their streams' data needs no bounds check (`careful` == false: no
     * row is one of a stream's last blocks or absent: 6 tasks in 7 of the bench batch; the checks -- 64-bit compares per
     * chunk and window -- were 7 % of the kernel's stall samples); otherwise each chunk is checked against the end of its
     * own row's stream. */
    const uint8_t *rbase = slot + (blk_off - p.byte_base);
    const uint8_t *rend = slot + ((uint64_t)size > p.byte_base ? (uint64_t)size - p.byte_base : 0u);
    /* every 16-byte chunk this row's windows will ever touch -- the last window may reach past the block's end, the
     * aligned superset 15 bytes past that -- lies inside its stream's data */
    const uint64_t row_span = (uint64_t)((bs + G::TB - 1) / G::TB) * G::TB + 16u;
    const bool careful = __any_sync(0xFFFFFFFFu, !in_batch || b >= p.block_end || blk_off + row_span > size);
    const uint8_t *ld_ptr[G::IN_LOADS];
    uint32_t ld_smem[G::IN_LOADS], ld_lane[G::IN_LOADS];
#pragma unroll
    for (int m = 0; m < G::IN_LOADS; m++) {
      const uint32_t f = lane + 32u * m;
      const uint32_t rr = f / G::IN_CHUNKS, cc = f % G::IN_CHUNKS;
      const bool on = f < (uint32_t)(G::IN_ROWS * G::IN_CHUNKS);
      ld_lane[m] = on ? rr * C : 0u;
      uintptr_t grr;
      bool row_in = true;
      if (SPAN) {
        grr = (uintptr_t)__shfl_sync(0xFFFFFFFFu, (unsigned long long)(uintptr_t)rbase, ld_lane[m]);
        row_in = __shfl_sync(0xFFFFFFFFu, (int)in_batch, ld_lane[m]) != 0;
      } else {   /* row 0 of the task is lane 0's; rows of one stream are bs bytes apart */
        grr = (uintptr_t)(slot + (AADF_FILE_HEADER_BYTES + (uint64_t)(p.block_begin + bo0) * bs - p.byte_base) + (uint64_t)rr * bs);
      }
      ld_ptr[m] = (on && row_in) ? (const uint8_t *)((grr & ~(uintptr_t)15) + 16u * cc) : nullptr;
      ld_smem[m] = rr * G::IN_PITCH + 16u * cc;
    }
    auto fetch_all = [&](uint4 (&pre)[G::IN_LOADS]) {
      if (!careful) {
#pragma unroll
        for (int m = 0; m < G::IN_LOADS; m++)
          pre[m] = (ld_ptr[m] != nullptr) ? __ldg(reinterpret_cast<const uint4 *>(ld_ptr[m])) : make_uint4(0u, 0u, 0u, 0u);
      } else {
#pragma unroll
        for (int m = 0; m < G::IN_LOADS; m++) {
          const uint8_t *end = SPAN ? (const uint8_t *)(uintptr_t)__shfl_sync(0xFFFFFFFFu, (unsigned long long)(uintptr_t)rend, ld_lane[m])
                                    : rend;
          pre[m] = dec_fetch16(ld_ptr[m], end);
        }
      }
    };

    /* reader role: this lane's block row in shared memory */
    const uint32_t a_r = (uint32_t)((uintptr_t)rbase & 15u);
    const unsigned char *irow = in_rows + row * G::IN_PITCH;
    unsigned char *orow = out_rows + lane * (BULK ? kDecBulkPitch : kDecOutPitch);
    auto in_u8 = [&](uint32_t pos) -> uint32_t { return irow[a_r + pos]; };

    DecChainT<SPAN> c;
    c.h0 = c.h1 = c.h2 = c.h3 = c.w0 = c.w1 = c.w2 = c.w3 = c.idx = 0;
    c.tables(tab);

    const uint32_t windows = (bs + G::TB - 1) / G::TB;
    uint4 pre[G::IN_LOADS];
    fetch_all(pre);

    uint32_t out_base = 0;                                   /* first output sample of this window */
    for (uint32_t w = 0; w < windows; w++) {
      /* stage this window, start the next one */
#pragma unroll
      for (int m = 0; m < G::IN_LOADS; m++)
        if (ld_ptr[m] != nullptr) *reinterpret_cast<uint4 *>(in_rows + ld_smem[m]) = pre[m];
      __syncwarp();
      if (w + 1 < windows) {
#pragma unroll
        for (int m = 0; m < G::IN_LOADS; m++)
          if (ld_ptr[m] != nullptr) ld_ptr[m] += G::TB;
        fetch_all(pre);
      }

      uint32_t produced = 0;      /* samples this chain wrote into its output row in this window */
      uint32_t pos = 0;           /* byte position inside the window */
      bool row_busy = BULK != 0;                  /* BULK: the previous window's row may still be on its way out */
      if (w == 0) {
        /* block header, src/aad_decoder.c:364-380: u16 (index << 4 | shift), 4 x (u16 weight, u16 history) */
        const uint32_t hp = AADF_CHANNEL_HEADER_BYTES * ch;
        const uint32_t head = (in_u8(hp) << 8) | in_u8(hp + 1);
        c.idx = (int32_t)(int16_t)(head >> 4);
        c.idx = max(0, min(c.idx, AADF_INDEX_MAX));    /* a corrupt header must not index outside the table */
        const uint32_t shift = head & 0xFu;
        int32_t wv[4], hv[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          wv[k] = (int32_t)((uint32_t)(int32_t)(int16_t)((in_u8(hp + 2 + 4 * k) << 8) | in_u8(hp + 3 + 4 * k)) << shift);
          hv[k] = (int32_t)(int16_t)((in_u8(hp + 4 + 4 * k) << 8) | in_u8(hp + 5 + 4 * k));
        }
        c.w0 = wv[0]; c.w1 = wv[1]; c.w2 = wv[2]; c.w3 = wv[3];
        c.h0 = hv[0]; c.h1 = hv[1]; c.h2 = hv[2]; c.h3 = hv[3];
        const int32_t first4[4] = { c.h3, c.h2, c.h1, c.h0 };   /* src/aad_decoder.c:386-391 */
        if (!BULK) dec_emit<4>(orow, 0, first4);
        produced = 4;
        pos = AADF_CHANNEL_HEADER_BYTES * C;
        if (BULK) {
          /* the 4 header samples and the half step's 4 as one 16-byte piece */
          const uint32_t v = in_u8(pos) | (in_u8(pos + 1) << 8);
          int32_t o[8] = { first4[0], first4[1], first4[2], first4[3], 0, 0, 0, 0 };
          dec_byte<BITS, 0>(c, v, tab, o + 4);
          dec_byte<BITS, 1>(c, v, tab, o + 6);
          dec_bulk_wait_read();          /* as late as possible: the row is written for the first time here */
          row_busy = false;
          dec_emit8(orow, 0, o);
          produced = 8;
          pos += G::HALF_BYTES;
        } else if (C == 1) {
          /* half a step: the 18-byte header leaves the row 2 bytes off word alignment */
          if (BITS == 3) {
            int32_t o[16];
            dec_group3(c, (in_u8(pos) << 16) | (in_u8(pos + 1) << 8) | in_u8(pos + 2), tab, o);
            dec_group3(c, (in_u8(pos + 3) << 16) | (in_u8(pos + 4) << 8) | in_u8(pos + 5), tab, o + 8);
            dec_emit<16>(orow, produced, o);
            produced += 16;
          } else {
            const uint32_t v = in_u8(pos) | (in_u8(pos + 1) << 8);
            int32_t o[2 * (BITS == 4 ? 2 : 4)];
            dec_byte<BITS, 0>(c, v, tab, o);
            dec_byte<BITS, 1>(c, v, tab, o + (BITS == 4 ? 2 : 4));
            dec_emit<2 * (BITS == 4 ? 2 : 4)>(orow, produced, o);
            produced += 2 * (BITS == 4 ? 2 : 4);
          }
          pos += G::HALF_BYTES;
        }
      }

      /* whole steps: sliding 32-bit words + funnel shift, any byte alignment */
      {
        const uint32_t at = a_r + pos;
        const uint32_t *wp = reinterpret_cast<const uint32_t *>(irow + (at & ~3u));
        const uint32_t sh = (at & 3u) * 8u;
        uint32_t lo = *wp++;
        /* the block's last window stops where the block's samples end (the same in every lane) */
        const uint32_t left = (spb > out_base + produced) ? spb - out_base - produced : 0u;
        const uint32_t steps = min((uint32_t)(G::TB - pos) / G::STEP_BYTES, (left + G::SPS - 1u) / G::SPS);
        /* two words per turn where a step is one word: halves the loop's bookkeeping (a 3-bit step is 3 words
         * = 32 samples already) */
#pragma unroll(BITS == 3 ? 1 : 2)
        for (uint32_t s = 0; s < steps; s++) {
          if (BITS == 3) {
            uint32_t x[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
              const uint32_t hi = *wp++;
              x[k] = __funnelshift_r(lo, hi, sh);
              lo = hi;
            }
            int32_t o[G::SPS];
            if (C == 1) {
              dec_group3(c, __byte_perm(x[0], 0u, 0x4012), tab, o);
              dec_group3(c, __byte_perm(x[0], x[1], 0x4345), tab, o + 8);
              dec_group3(c, __byte_perm(x[1], x[2], 0x4234), tab, o + 16);
              dec_group3(c, __byte_perm(x[2], 0u, 0x4123), tab, o + 24);
            } else {   /* two channels: groups alternate, 3 bytes each */
              const uint32_t sel_a = ch ? 0x0345u : 0x0012u, sel_b = ch ? 0x0567u : 0x0234u;
              dec_group3(c, __byte_perm(x[0], x[1], sel_a), tab, o);
              dec_group3(c, __byte_perm(x[1], x[2], sel_b), tab, o + 8);
            }
            dec_emit<G::SPS>(orow, produced, o);
          } else {
            const uint32_t hi = *wp++;
            uint32_t v = __funnelshift_r(lo, hi, sh);
            lo = hi;
            constexpr int PER_BYTE = (BITS == 4) ? 2 : 4;
            int32_t o[G::SPS];
            if (C == 1) {
              dec_byte<BITS, 0>(c, v, tab, o);
              dec_byte<BITS, 1>(c, v, tab, o + PER_BYTE);
              dec_byte<BITS, 2>(c, v, tab, o + 2 * PER_BYTE);
              dec_byte<BITS, 3>(c, v, tab, o + 3 * PER_BYTE);
            } else {   /* two channels: bytes alternate */
              v >>= 8u * ch;
              dec_byte<BITS, 0>(c, v, tab, o);
              dec_byte<BITS, 2>(c, v, tab, o + PER_BYTE);
            }
            if (BULK) {
              if (row_busy) {
                dec_bulk_wait_read();
                row_busy = false;
              }
              dec_emit8(orow, produced, o);
            } else {
              dec_emit<G::SPS>(orow, produced, o);
            }
          }
          produced += G::SPS;
        }
      }
      if (BULK) {
        /* every lane hands its own row to the TMA unit: whole 16-byte pieces in bulk, a ragged tail (a stream's last
         * block, a short buffer) by itself */
        const uint32_t room = n_row > out_base ? n_row - out_base : 0u;
        const uint32_t count = min(min(produced, spb - out_base), room);
        const uint32_t whole = count & ~7u;
        dec_bulk_fence();
        if (whole) dec_bulk_store(grow + out_base, orow, whole * 2u);
        dec_bulk_commit();
        for (uint32_t k = whole; k < count; k++) grow[out_base + k] = *reinterpret_cast<const int16_t *>(orow + 2u * k);
        out_base += produced;
        __syncwarp();                 /* every lane is done with the input rows before the next window is staged */
        continue;
      }
      __syncwarp();

      /* flush: row rr of the warp goes out as one coalesced run of 8-byte pieces */
      const bool mine = lane * 4u + 4u <= min(produced, spb - out_base);   /* this lane's 4 samples of every whole-block row */
      if (IL) {
        /* WAV order: the two rows of a block leave as frames, 16 bytes (4 frames) per lane and block */
        if (mine) {
          const unsigned char *srow = out_rows + 8u * lane;
          int16_t *dst = grow0 + (uint64_t)(out_base + 4u * lane) * C;
          for (uint32_t rr = 0; rr < nfull; rr += C, dst += (uint64_t)spb * C, srow += C * kDecOutPitch)
            dec_flush_frames4<2>(srow, dst, ms2);
          if (nfull1 != 0u) {
            srow = out_rows + 8u * lane + seg1 * kDecOutPitch;
            dst = grow1 + (uint64_t)(out_base + 4u * lane) * C;
            for (uint32_t rr = 0; rr < nfull1; rr += C, dst += (uint64_t)spb * C, srow += C * kDecOutPitch)
              dec_flush_frames4<2>(srow, dst, ms2);
          }
        }
        if (!all_full) dec_flush_ragged_frames<2>(out_rows, ragged, n_row, produced, grow, out_base, lane, ms2);
      } else if (ms2) {
        /* stereo mid/side: the left / right planes are formed here, from the two rows of a block (src/aad_decoder.c:458-470) */
        if (mine) {
          dec_flush_ms_rows(out_rows + 8u * lane, grow0 + out_base + 4u * lane, nfull, p.pcm_ch_stride, spb);
          if (nfull1 != 0u)
            dec_flush_ms_rows(out_rows + 8u * lane + seg1 * kDecOutPitch, grow1 + out_base + 4u * lane, nfull1, p.pcm_ch_stride, spb);
        }
        if (!all_full) dec_flush_ragged(out_rows, ragged, n_row, produced, grow, out_base, lane, true);
      } else if (all_full) {
        /* every chain of the warp delivers a whole block of one stream: row addresses are plain arithmetic and the
         * count is the same for every row -- `produced` (identical in every lane), clipped where the
         * last window runs past the block; both are multiples of 4 */
        if (mine) {
          const unsigned char *srow = out_rows + 8u * lane;
          int16_t *dst = grow0 + out_base + 4u * lane;
#pragma unroll
          for (uint32_t rr = 0; rr < 32; rr++)
            *reinterpret_cast<uint2 *>(dst + (uint64_t)(rr % C) * p.pcm_ch_stride + (uint64_t)(rr / C) * spb) =
                *reinterpret_cast<const uint2 *>(srow + rr * kDecOutPitch);
        }
      } else {
        /* the whole-block rows of the (up to) two arithmetic segments, the rest one by one */
        if (mine) {
          const unsigned char *srow = out_rows + 8u * lane;
          int16_t *dst = grow0 + out_base + 4u * lane;
          for (uint32_t rr = 0; rr < nfull; rr++)
            *reinterpret_cast<uint2 *>(dst + (uint64_t)(rr % C) * p.pcm_ch_stride + (uint64_t)(rr / C) * spb) =
                *reinterpret_cast<const uint2 *>(srow + rr * kDecOutPitch);
          srow += seg1 * kDecOutPitch;
          dst = grow1 + out_base + 4u * lane;
          for (uint32_t rr = 0; rr < nfull1; rr++)
            *reinterpret_cast<uint2 *>(dst + (uint64_t)(rr % C) * p.pcm_ch_stride + (uint64_t)(rr / C) * spb) =
                *reinterpret_cast<const uint2 *>(srow + rr * kDecOutPitch);
        }
        dec_flush_ragged(out_rows, ragged, n_row, produced, grow, out_base, lane);
      }
      out_base += produced;       /* identical in every lane */
      __syncwarp();
    }
  }
  if (BULK) dec_bulk_wait_read();   /* shared memory stays valid until the last rows have been read */
}

/*
 * aad_decode_wide<BITS> -- the same staged data path for any channel count up to 32 (run time C):
 * the multichannel streams of BASELINE config 4 (8 channels, 3-bit).  A warp owns floor(32 / C)
 * consecutive blocks, lane = (block row, channel); the lanes past rows * C idle.  Windows are the
 * same 128 samples per chain (16 * BITS * C input bytes per block).  A block interleaves the
 * channels group by group (src/aad_decoder.c:394-455), so a chain's codes are GB bytes every
 * GB * C bytes: each lane picks its own bytes out of the shared input row (the lanes of a row read
 * neighbouring bytes of the same words, the rows sit 4 banks apart or more) -- byte reads instead
 * of the mono / stereo kernel's sliding words, everything else as there: 16-byte coalesced global
 * loads one window ahead, 8-byte shared output stores, coalesced 8-byte row stores to the PCM planes.
 */
template <int BITS, int IL>   /* IL = 1: output in WAV order (4 or 8 channels), frames formed in the flush */
__global__ void __launch_bounds__(kDecWarps * 32) aad_decode_wide(const aadk_decode_params p)
{
  constexpr uint32_t GB = (BITS == 3) ? 3 : 1;
  constexpr int kLoads = BITS + 1;          /* ceil(rows * chunks / 32) <= BITS + 1 for every C */
  extern __shared__ __align__(128) unsigned char dec_smem[];
  DecTables &tab = *reinterpret_cast<DecTables *>(dec_smem);
  dec_load_tables<BITS>(tab);

  const uint32_t C = p.geo.channels;
  const uint32_t rows = 32u / C, active = rows * C;
  const uint32_t TB = 16u * BITS * C;                  /* input bytes per block per window */
  const uint32_t chunks = BITS * C + 1u;               /* 16-byte chunks per row: aligned superset */
  const uint32_t pitch = 16u * (chunks | 1u);              /* odd number of chunks: rows spread over the banks (DecGeom) */
  const uint32_t in_bytes = (rows * pitch + 16u + 15u) & ~15u;

  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = threadIdx.x >> 5;
  unsigned char *in_rows = dec_smem + ((sizeof(DecTables) + 15) & ~(size_t)15) + (size_t)warp * (in_bytes + 32u * kDecOutPitch);
  unsigned char *out_rows = in_rows + in_bytes;

  const uint32_t spb = p.geo.samples_per_block;
  const uint32_t bs = p.geo.block_size;
  const uint32_t nblocks = p.block_end - p.block_begin;
  const uint32_t warps_per_stream = (nblocks + rows - 1) / rows;
  const uint64_t total_warps = (uint64_t)p.num_streams * warps_per_stream;
  /* persistent: the warps of the grid share out the warp tasks round robin (the tables are built once per CTA) */
  for (uint64_t gw = (uint64_t)blockIdx.x * kDecWarps + warp; gw < total_warps; gw += (uint64_t)gridDim.x * kDecWarps) {
    const uint64_t stream = gw / warps_per_stream;
    const uint32_t b0 = p.block_begin + (uint32_t)(gw % warps_per_stream) * rows;

    const uint8_t *slot = p.aad + stream * p.aad_stride;
    const uint32_t size = p.sizes ? p.sizes[stream] : p.uniform_size;
    uint32_t ns = p.uniform_samples;
    if (p.read_headers) ns = dec_header_samples(slot, size, p.uniform_samples);
    const uint32_t buf = p.buf_samples ? p.buf_samples : ns;

    /* this lane's chain (idle lanes shadow row 0 and deliver nothing) */
    const bool lane_on = lane < active;
    const uint32_t row = lane_on ? lane / C : 0u, ch = lane_on ? lane % C : 0u;
    const uint32_t b = b0 + row;
    const uint64_t blk_off = AADF_FILE_HEADER_BYTES + (uint64_t)b * bs;
    const bool have = lane_on && b < p.block_end && (uint64_t)b * spb < ns &&
                      blk_off + (uint64_t)AADF_CHANNEL_HEADER_BYTES * C <= size;
    const uint32_t n_row = (have && (uint64_t)b * spb < buf) ? min(spb, buf - b * spb) : 0u;
    constexpr bool il = IL != 0;   /* WAV order (C = 4 or 8 here: aadk_decode_interleaved_ok) */
    const uint64_t srel = (uint64_t)b * spb - p.sample_base, srel0 = (uint64_t)b0 * spb - p.sample_base;
    int16_t *grow = (int16_t *)p.pcm + stream * p.pcm_clip_stride + (il ? srel * C : (uint64_t)ch * p.pcm_ch_stride + srel);
    int16_t *grow0 = (int16_t *)p.pcm + stream * p.pcm_clip_stride + (il ? srel0 * C : srel0);
    /* rows [0, nfull) deliver whole blocks; a multiple of C (the channels of a block share its sample count) */
    const uint32_t nfull = min(active, (uint32_t)__ffs((int)~__ballot_sync(0xFFFFFFFFu, n_row == spb)) - 1u);
    const bool all_full = nfull >= active;

    /* loader role */
    const uint8_t *g0 = slot + (AADF_FILE_HEADER_BYTES + (uint64_t)b0 * bs - p.byte_base);
    const uint8_t *ld_ptr[kLoads];
    uint32_t ld_smem[kLoads];
#pragma unroll
    for (int m = 0; m < kLoads; m++) {
      const uint32_t f = lane + 32u * m;
      const uint32_t rr = f / chunks, cc = f % chunks;
      const uintptr_t grr = (uintptr_t)(g0 + (uint64_t)rr * bs);
      ld_ptr[m] = (f < rows * chunks) ? (const uint8_t *)((grr & ~(uintptr_t)15) + 16u * cc) : nullptr;
      ld_smem[m] = rr * pitch + 16u * cc;
    }
    const uint8_t *slot_end = slot + ((uint64_t)size > p.byte_base ? (uint64_t)size - p.byte_base : 0u);
    auto fetch = [&](int m) -> uint4 { return dec_fetch16(ld_ptr[m], slot_end); };

    /* reader role */
    const uint32_t a_r = (uint32_t)((uintptr_t)(g0 + (uint64_t)row * bs) & 15u);
    const unsigned char *irow = in_rows + row * pitch + a_r;
    unsigned char *orow = out_rows + lane * kDecOutPitch;

    DecChain c;
    c.h0 = c.h1 = c.h2 = c.h3 = c.w0 = c.w1 = c.w2 = c.w3 = c.idx = 0;
    c.tables(tab);

    const uint32_t windows = (bs + TB - 1) / TB;
    const uint32_t gstride = GB * C;
    uint4 pre[kLoads];
#pragma unroll
    for (int m = 0; m < kLoads; m++) pre[m] = fetch(m);

    uint32_t out_base = 0;
    for (uint32_t w = 0; w < windows; w++) {
#pragma unroll
      for (int m = 0; m < kLoads; m++)
        if (ld_ptr[m] != nullptr) *reinterpret_cast<uint4 *>(in_rows + ld_smem[m]) = pre[m];
      __syncwarp();
      if (w + 1 < windows) {
#pragma unroll
        for (int m = 0; m < kLoads; m++) {
          if (ld_ptr[m] != nullptr) ld_ptr[m] += TB;
          pre[m] = fetch(m);
        }
      }

      uint32_t produced = 0;
      uint32_t pos = 0;
      if (w == 0) {   /* block header, src/aad_decoder.c:364-391 */
        const unsigned char *hp = irow + AADF_CHANNEL_HEADER_BYTES * ch;
        const uint32_t head = ((uint32_t)hp[0] << 8) | hp[1];
        c.idx = (int32_t)(int16_t)(head >> 4);
        c.idx = max(0, min(c.idx, AADF_INDEX_MAX));    /* a corrupt header must not index outside the table */
        const uint32_t shift = head & 0xFu;
        int32_t wv[4], hv[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          wv[k] = (int32_t)((uint32_t)(int32_t)(int16_t)(((uint32_t)hp[2 + 4 * k] << 8) | hp[3 + 4 * k]) << shift);
          hv[k] = (int32_t)(int16_t)(((uint32_t)hp[4 + 4 * k] << 8) | hp[5 + 4 * k]);
        }
        c.w0 = wv[0]; c.w1 = wv[1]; c.w2 = wv[2]; c.w3 = wv[3];
        c.h0 = hv[0]; c.h1 = hv[1]; c.h2 = hv[2]; c.h3 = hv[3];
        const int32_t first4[4] = { c.h3, c.h2, c.h1, c.h0 };
        dec_emit<4>(orow, 0, first4);
        produced = 4;
        pos = AADF_CHANNEL_HEADER_BYTES * C;
      }

      /* this chain's groups of the window: GB bytes every GB * C bytes; 4 (2-bit, 4-bit) or 8 (3-bit)
       * samples per turn so the shared output row takes whole 8-byte pieces */
      {
        const unsigned char *bp = irow + pos + GB * ch;
        /* whole window, or what is left of the block's samples (the same in every lane); the 4-bit loop takes
         * groups in pairs and may run one group past that, still inside the window */
        constexpr uint32_t GS = (BITS == 4) ? 2 : (BITS == 3 ? 8 : 4);
        const uint32_t left = (spb > out_base + produced) ? spb - out_base - produced : 0u;
        const uint32_t groups = min((TB - pos) / gstride, (left + GS - 1u) / GS);
        if (BITS == 4) {
#pragma unroll 2
          for (uint32_t g = 0; g < groups; g += 2) {
            const uint32_t v = (uint32_t)bp[0] | ((uint32_t)bp[gstride] << 8);
            bp += 2u * gstride;
            int32_t o[4];
            dec_byte<4, 0>(c, v, tab, o);
            dec_byte<4, 1>(c, v, tab, o + 2);
            dec_emit<4>(orow, produced, o);
            produced += 4;
          }
        } else if (BITS == 3) {
#pragma unroll 2
          for (uint32_t g = 0; g < groups; g++) {
            const uint32_t v = ((uint32_t)bp[0] << 16) | ((uint32_t)bp[1] << 8) | bp[2];
            bp += gstride;
            int32_t o[8];
            dec_group3(c, v, tab, o);
            dec_emit<8>(orow, produced, o);
            produced += 8;
          }
        } else {
#pragma unroll 2
          for (uint32_t g = 0; g < groups; g++) {
            const uint32_t v = bp[0];
            bp += gstride;
            int32_t o[4];
            dec_byte<2, 0>(c, v, tab, o);
            dec_emit<4>(orow, produced, o);
            produced += 4;
          }
        }
      }
      __syncwarp();

      /* flush: row rr of the warp goes out as one coalesced run of 8-byte pieces.  The leading whole-block rows
       * (all of them but in a stream's last warp task) have arithmetic addresses and one common count --
       * `produced` (identical in every lane), clipped where the last window runs past the block; both are
       * multiples of 4 -- the rest go one by one */
      if (il) {
        /* WAV order: the C rows of a block leave as frames, 4 frames (8 C bytes) per lane and block */
        if (lane * 4u + 4u <= min(produced, spb - out_base)) {
          const unsigned char *srow = out_rows + 8u * lane;
          int16_t *dst = grow0 + (uint64_t)(out_base + 4u * lane) * C;
          for (uint32_t rb = 0; rb < nfull / C; rb++, dst += (uint64_t)spb * C, srow += C * kDecOutPitch) {
            if (C == 8) dec_flush_frames4<8>(srow, dst, false);
            else dec_flush_frames4<4>(srow, dst, false);
          }
        }
        if (!all_full) {
          if (C == 8) dec_flush_ragged_frames<8>(out_rows, dec_lanes(nfull, active), n_row, produced, grow, out_base, lane, false);
          else dec_flush_ragged_frames<4>(out_rows, dec_lanes(nfull, active), n_row, produced, grow, out_base, lane, false);
        }
      } else {
      if (lane * 4u + 4u <= min(produced, spb - out_base)) {
        /* row = (block, channel): channels step by the plane pitch, blocks by spb samples */
        const unsigned char *srow = out_rows + 8u * lane;
        int16_t *dst_blk = grow0 + out_base + 4u * lane;
        for (uint32_t rb = 0; rb < nfull / C; rb++, dst_blk += spb) {
          int16_t *dst = dst_blk;
#pragma unroll 4
          for (uint32_t rc = 0; rc < C; rc++, dst += p.pcm_ch_stride, srow += kDecOutPitch)
            *reinterpret_cast<uint2 *>(dst) = *reinterpret_cast<const uint2 *>(srow);
        }
      }
      if (!all_full) dec_flush_ragged(out_rows, dec_lanes(nfull, active), n_row, produced, grow, out_base, lane);
      }
      out_base += produced;       /* identical in every lane */
      __syncwarp();
    }
  }
}

/*
 * aad_decode_tma<BITS> -- A/B variant of aad_decode_fast<BITS, 1> (mono 4-bit / 2-bit, planar output, tasks inside one
 * stream) whose input staging is done by the TMA unit: the batch's blocks are described to it as a 3-D tensor
 * [stream][block][byte of the block] (CUtensorMap built per launch, dec_tma_launch), and one elected lane per warp
 * asks for the window of all 32 block rows of its task with ONE cp.async.bulk.tensor (UTMALDG in the SASS) that
 * completes on an mbarrier; two windows are in flight per warp (two shared buffers).  It replaces the register
 * prefetch (IN_LOADS 16-byte loads per lane and window held in registers), the shared stores and the per-chunk
 * pointer bookkeeping of aad_decode_fast.  The box is [32 rows][TB bytes] with the TMA's 64-byte (32-byte for 2-bit)
 * swizzle, which spreads the 32 rows' words over 8 banks exactly like aad_decode_fast's odd row pitch.
 *
 * TMA needs 16-byte aligned global addresses and strides: block 0 of stream 0 (p.aad + 31) 16-byte aligned, block
 * size and stream stride multiples of 16 (dec_tma_eligible; anything else takes the default kernel).  Rows past the
 * tensor's block extent are zero-filled by the unit; the extent counts the whole blocks inside uniform_size, and the
 * launcher only takes this path when every block the batch can deliver is one of those.
 * Selected by AADGpu_SetKernelPath(7) only: measured against the default kernel in profiles/r02_decoder_experiments.md.
 */
__device__ __forceinline__ void dec_mbar_init(uint32_t bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void dec_mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dec_mbar_wait(uint32_t bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "DEC_TMA_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DEC_TMA_DONE;\n"
      "bra DEC_TMA_WAIT;\n"
      "DEC_TMA_DONE:\n"
      "}\n" :: "r"(bar), "r"(parity) : "memory");
}
/* box at (byte c0 of the block, block c1, stream c2) -> shared memory at dst, completion counted on bar */
__device__ __forceinline__ void dec_tma_load_3d(uint32_t dst, const void *tmap, uint32_t bar, int32_t c0, int32_t c1, int32_t c2)
{
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               :: "r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

/* HALF = 1: the output rows hold 64 samples and are flushed twice per window (two rows per store instruction), which
 * brings a warp's shared memory from 12.5 KB to 8.3 KB: 24 warps per SM (64 registers allow it) instead of 16. */
template <int BITS, int HALF>
struct DecTmaGeom {
  static constexpr int TB = 16 * BITS;                       /* input bytes per block per window (mono) */
  static constexpr int STAGE_BYTES = 32 * TB;                 /* one box: 32 block rows */
  static constexpr int SWZ_SHIFT = (TB == 64) ? 1 : 2;        /* 64-byte swizzle: address bits 7-8 onto bits 4-5; 32-byte: bit 7 onto bit 4 */
  static constexpr uint32_t SWZ_MASK = (TB == 64) ? 3u : 1u;
  static constexpr int WARPS = HALF ? 24 : kDecWarps;
  static constexpr int OUT_SAMPLES = HALF ? 64 : kDecWindow;  /* samples an output row holds between flushes */
  static constexpr int OUT_PITCH = 2 * OUT_SAMPLES + 8;       /* conflict-free 8-byte shared accesses, as kDecOutPitch */
  /* tables | (pad to 1024) | 2 stages x warps | output rows x warps | 2 mbarriers x warps */
  static constexpr size_t SMEM = ((sizeof(DecTables) + 1023) & ~(size_t)1023) + 1024 +
                                 (size_t)WARPS * (2 * STAGE_BYTES + 32 * OUT_PITCH + 16);
};

/* dec_flush_ragged for rows of any pitch (planar mono rows only) */
template <int PITCH>
__device__ __forceinline__ void dec_flush_ragged_pitch(const unsigned char *out_rows, uint32_t todo, uint32_t n_row, uint32_t produced,
                                                       int16_t *grow, uint32_t out_base, uint32_t lane)
{
  uint32_t live = __ballot_sync(0xFFFFFFFFu, n_row > out_base) & todo;
  for (; live != 0u; live &= live - 1u) {
    const uint32_t rr = (uint32_t)__ffs((int)live) - 1u;
    const uint32_t n_rr = __shfl_sync(0xFFFFFFFFu, n_row, rr);
    const uint64_t gp = __shfl_sync(0xFFFFFFFFu, (unsigned long long)(uintptr_t)grow, rr);
    const uint32_t count = min(produced, n_rr - out_base);
    int16_t *dst = reinterpret_cast<int16_t *>((uintptr_t)gp) + out_base;
    const unsigned char *srow = out_rows + rr * PITCH;
    const uint32_t s0 = lane * 4u;
    if (s0 + 4u <= count) {
      *reinterpret_cast<uint2 *>(dst + s0) = *reinterpret_cast<const uint2 *>(srow + 2u * s0);
    } else {
      for (uint32_t k = s0; k < count; k++) dst[k] = *reinterpret_cast<const int16_t *>(srow + 2u * k);
    }
  }
}

template <int BITS, int HALF>
__global__ void __launch_bounds__(DecTmaGeom<BITS, HALF>::WARPS * 32, 1) aad_decode_tma(const aadk_decode_params p, const __grid_constant__ CUtensorMap tmap)
{
  static_assert(BITS == 4 || BITS == 2, "mono 4-bit / 2-bit");
  using G = DecGeom<BITS, 1>;
  using T = DecTmaGeom<BITS, HALF>;
  constexpr int kWarps = T::WARPS;
  extern __shared__ __align__(128) unsigned char dec_smem[];
  DecTables &tab = *reinterpret_cast<DecTables *>(dec_smem);
  dec_load_tables<BITS>(tab);

  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = threadIdx.x >> 5;
  /* the stages sit at 1024-byte aligned shared addresses: the swizzle is a function of the address */
  const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(dec_smem);
  const uint32_t stages0 = (smem0 + (uint32_t)sizeof(DecTables) + 1023u) & ~1023u;
  unsigned char *stage_base = dec_smem + (stages0 - smem0) + (size_t)warp * (2 * T::STAGE_BYTES);
  unsigned char *out_rows = dec_smem + (stages0 - smem0) + (size_t)kWarps * (2 * T::STAGE_BYTES) + (size_t)warp * (32 * T::OUT_PITCH);
  const uint32_t stage_addr = stages0 + warp * (2u * T::STAGE_BYTES);
  const uint32_t bar0 = stages0 + kWarps * (2u * T::STAGE_BYTES + 32u * T::OUT_PITCH) + warp * 16u;
  if (lane == 0) {
    dec_mbar_init(bar0, 1u);
    dec_mbar_init(bar0 + 8u, 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  const uint32_t spb = p.geo.samples_per_block;
  const uint32_t bs = p.geo.block_size;
  const uint32_t nblocks = p.block_end - p.block_begin;
  const uint32_t warps_per_stream = (nblocks + 31u) / 32u;
  const uint64_t total_warps = (uint64_t)p.num_streams * warps_per_stream;
  const uint32_t windows = (bs + T::TB - 1) / T::TB;
  const uint32_t xr = ((lane >> T::SWZ_SHIFT) & T::SWZ_MASK) << 4;   /* this lane's row: byte o of its window sits at (lane * TB + o) ^ xr */
  unsigned char *orow = out_rows + lane * T::OUT_PITCH;
  uint32_t it = 0;                                                   /* windows this warp has waited for: stage it & 1, parity (it >> 1) & 1 */

  /* The warp's windows -- of all its tasks, in order -- form ONE pipeline two windows deep: the first windows of the next
   * task are asked for while the last windows of the current one are decoded (with loads held in registers that
   * look-ahead would cost registers; here it is two integers in lane 0). */
  const uint64_t gw_step = (uint64_t)gridDim.x * kWarps;
  uint64_t next_gw = (uint64_t)blockIdx.x * kWarps + warp;        /* the window to ask for next: task, */
  uint32_t next_w = 0;                                               /* window of that task */
  auto issue_next = [&](uint32_t st) {
    if (lane == 0 && next_gw < total_warps) {
      const uint64_t nstream = next_gw / warps_per_stream;
      const uint32_t nbo0 = (uint32_t)(next_gw - nstream * warps_per_stream) * 32u;
      dec_mbar_expect_tx(bar0 + 8u * st, T::STAGE_BYTES);
      dec_tma_load_3d(stage_addr + st * T::STAGE_BYTES, &tmap, bar0 + 8u * st, (int32_t)(next_w * T::TB),
                      (int32_t)(p.block_begin + nbo0), (int32_t)nstream);
      if (++next_w == windows) {
        next_w = 0;
        next_gw += gw_step;
      }
    }
  };
  issue_next(0u);
  issue_next(1u);

  for (uint64_t gw = (uint64_t)blockIdx.x * kWarps + warp; gw < total_warps; gw += gw_step) {
    const uint64_t stream = gw / warps_per_stream;
    const uint32_t bo0 = (uint32_t)(gw % warps_per_stream) * 32u;
    const uint32_t b = p.block_begin + bo0 + lane;

    const uint8_t *slot = p.aad + stream * p.aad_stride;
    const uint32_t size = p.uniform_size;
    uint32_t ns = p.uniform_samples;
    if (p.read_headers) ns = dec_header_samples(slot, size, p.uniform_samples);
    const uint32_t buf = p.buf_samples ? p.buf_samples : ns;
    const uint64_t blk_off = AADF_FILE_HEADER_BYTES + (uint64_t)b * bs;
    const bool have = b < p.block_end && (uint64_t)b * spb < ns && blk_off + (uint64_t)AADF_CHANNEL_HEADER_BYTES <= size;
    const uint32_t n_row = (have && (uint64_t)b * spb < buf) ? min(spb, buf - b * spb) : 0u;
    const uint64_t srel = (uint64_t)b * spb - p.sample_base;
    int16_t *grow = (int16_t *)p.pcm + stream * p.pcm_clip_stride + srel;
    const uint32_t full_mask = __ballot_sync(0xFFFFFFFFu, n_row == spb);
    const uint32_t nfull = min((uint32_t)__ffs((int)~full_mask) - 1u, 32u);
    const bool all_full = nfull >= 32u;
    int16_t *grow0 = (int16_t *)p.pcm + stream * p.pcm_clip_stride + ((uint64_t)(p.block_begin + bo0) * spb - p.sample_base);
    const uint32_t ragged = ~dec_lanes(0u, nfull);

    DecChainT<1> c;   /* complete shared-space table addresses: with generic pointers ptxas re-adds the table base per lookup here */
    c.h0 = c.h1 = c.h2 = c.h3 = c.w0 = c.w1 = c.w2 = c.w3 = c.idx = 0;
    c.tables(tab);

    uint32_t out_base = 0;
    for (uint32_t w = 0; w < windows; w++, it++) {
      const uint32_t st = it & 1u;
      dec_mbar_wait(bar0 + 8u * st, (it >> 1) & 1u);
      const unsigned char *irow = stage_base + st * T::STAGE_BYTES + lane * T::TB;
      auto in_u8 = [&](uint32_t o) -> uint32_t { return irow[o ^ xr]; };
      auto in_u32 = [&](uint32_t o) -> uint32_t { return *reinterpret_cast<const uint32_t *>(irow + (o ^ xr)); };

      uint32_t produced = 0, pos = 0;
      if (w == 0) {
        /* block header, src/aad_decoder.c:364-380 (see aad_decode_fast) */
        const uint32_t head = (in_u8(0) << 8) | in_u8(1);
        c.idx = (int32_t)(int16_t)(head >> 4);
        c.idx = max(0, min(c.idx, AADF_INDEX_MAX));
        const uint32_t shift = head & 0xFu;
        int32_t wv[4], hv[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          wv[k] = (int32_t)((uint32_t)(int32_t)(int16_t)((in_u8(2 + 4 * k) << 8) | in_u8(3 + 4 * k)) << shift);
          hv[k] = (int32_t)(int16_t)((in_u8(4 + 4 * k) << 8) | in_u8(5 + 4 * k));
        }
        c.w0 = wv[0]; c.w1 = wv[1]; c.w2 = wv[2]; c.w3 = wv[3];
        c.h0 = hv[0]; c.h1 = hv[1]; c.h2 = hv[2]; c.h3 = hv[3];
        const int32_t first4[4] = { c.h3, c.h2, c.h1, c.h0 };   /* src/aad_decoder.c:386-391 */
        dec_emit<4>(orow, 0, first4);
        produced = 4;
        pos = AADF_CHANNEL_HEADER_BYTES;
        /* half a step brings the read position to a word boundary (the window itself is 16-byte aligned here) */
        const uint32_t v = in_u8(pos) | (in_u8(pos + 1) << 8);
        constexpr int PB = (BITS == 4) ? 2 : 4;
        int32_t o[2 * PB];
        dec_byte<BITS, 0>(c, v, tab, o);
        dec_byte<BITS, 1>(c, v, tab, o + PB);
        dec_emit<2 * PB>(orow, produced, o);
        produced += 2 * PB;
        pos += G::HALF_BYTES;
      }
      /* the window's whole steps, in one go (HALF = 0) or in two (HALF = 1: the output row holds 64 samples), a flush
       * after each */
#pragma unroll 1
      for (int part = 0; part < (HALF ? 2 : 1); part++) {
        const uint32_t left = (spb > out_base + produced) ? spb - out_base - produced : 0u;
        uint32_t steps = min((uint32_t)(T::TB - pos) / G::STEP_BYTES, (left + G::SPS - 1u) / G::SPS);
        if (HALF) steps = min(steps, ((uint32_t)T::OUT_SAMPLES - produced) / G::SPS);
#pragma unroll 2
        for (uint32_t s = 0; s < steps; s++) {
          const uint32_t v = in_u32(pos);
          pos += 4;
          constexpr int PER_BYTE = (BITS == 4) ? 2 : 4;
          int32_t o[G::SPS];
          dec_byte<BITS, 0>(c, v, tab, o);
          dec_byte<BITS, 1>(c, v, tab, o + PER_BYTE);
          dec_byte<BITS, 2>(c, v, tab, o + 2 * PER_BYTE);
          dec_byte<BITS, 3>(c, v, tab, o + 3 * PER_BYTE);
          dec_emit<G::SPS>(orow, produced, o);
          produced += G::SPS;
        }
        __syncwarp();   /* every lane has written its output row (and, after the last part, read its row of this stage) */
        if (part == (HALF ? 1 : 0))
          issue_next(st);   /* the window after next -- of this task or of the warp's next one -- into the stage just read */

        /* flush: as aad_decode_fast's mono planar paths; HALF: rows of 64 samples, two of them per store instruction */
        const uint32_t count = min(produced, spb - out_base);                 /* identical in every lane, a multiple of 4 */
        constexpr uint32_t LPR = HALF ? 16u : 32u;                            /* lanes per row */
        const uint32_t col = (lane % LPR) * 4u, sub = lane / LPR;             /* this lane's 4 samples, its row of a pair */
        const bool mine = col + 4u <= count;
        if (mine) {
          const uint32_t rows_arith = all_full ? 32u : nfull;
          const unsigned char *srow = out_rows + 2u * col + sub * T::OUT_PITCH;
          int16_t *dst = grow0 + out_base + col + (uint64_t)sub * spb;
          if (all_full) {
#pragma unroll
            for (uint32_t rr = 0; rr < 32; rr += 32u / LPR)
              *reinterpret_cast<uint2 *>(dst + (uint64_t)rr * spb) = *reinterpret_cast<const uint2 *>(srow + rr * T::OUT_PITCH);
          } else {
            for (uint32_t rr = sub; rr < rows_arith; rr += 32u / LPR)
              *reinterpret_cast<uint2 *>(dst + (uint64_t)(rr - sub) * spb) = *reinterpret_cast<const uint2 *>(srow + (rr - sub) * T::OUT_PITCH);
          }
        }
        if (!all_full) dec_flush_ragged_pitch<T::OUT_PITCH>(out_rows, ragged, n_row, count, grow, out_base, lane);
        out_base += produced;
        produced = 0;
        __syncwarp();
      }
    }
  }
}


inline bool dec_fast_eligible(const aadk_decode_params &p)
{
  if (p.geo.channels < 1 || p.geo.channels > 32) return false;
  if (p.geo.samples_per_block % 4u) return false;
  if (((uintptr_t)p.pcm & 7u) || (p.pcm_clip_stride % 4u) || (!p.interleaved && (p.pcm_ch_stride % 4u))) return false;
  if (p.sample_base % 4u) return false;
  if (p.interleaved && p.geo.channels > 1 && (((uintptr_t)p.pcm & 15u) || (p.pcm_clip_stride % 8u))) return false;
  /* the window arithmetic assumes the canonical block layout: header, then whole groups */
  const uint32_t gs = aadf_group_samples(p.geo.bits), gb = aadf_group_bytes(p.geo.bits);
  if (p.geo.samples_per_block < AADF_TAPS || (p.geo.samples_per_block - AADF_TAPS) % gs) return false;
  if (p.geo.block_size != p.geo.channels * (AADF_CHANNEL_HEADER_BYTES + (p.geo.samples_per_block - AADF_TAPS) / gs * gb))
    return false;
  return true;
}

#ifndef AAD_DEC_PERSIST
#define AAD_DEC_PERSIST 1
#endif
/* One wave of CTAs (SMs x resident CTAs per SM), each warp looping over its share of the warp tasks,
 * so the shared-memory tables are built once per resident CTA instead of once per 4 tasks. */
template <typename K>
int dec_persistent_grid(K kernel, size_t smem, uint64_t warps, unsigned *grid)
{
  const uint64_t ctas = (warps + kDecWarps - 1) / kDecWarps;
  int dev = 0, sms = 148, per_sm = 1;
  if (int rc = device_sm_count(&dev, &sms)) return rc;
  if (int rc = allow_dynamic_smem(kernel, dev, smem)) return rc;
  if (int rc = resident_ctas(kernel, dev, kDecWarps * 32, smem, &per_sm)) return rc;
  const uint64_t wave = (uint64_t)sms * (uint64_t)(per_sm > 0 ? per_sm : 1);
  *grid = (unsigned)((AAD_DEC_PERSIST && ctas > wave) ? wave : ctas);
  return 0;
}

template <int BITS, int C, int IL, int BULK, int SPAN>
int dec_fast_launch_as(const aadk_decode_params &p, cudaStream_t s)
{
  using G = DecGeom<BITS, C>;
  const size_t smem = ((sizeof(DecTables) + 15) & ~(size_t)15) + (size_t)kDecWarps * (BULK ? G::WARP_BYTES_BULK : G::WARP_BYTES);
  const uint32_t nblocks = p.block_end - p.block_begin;
  const uint64_t warps = SPAN ? ((uint64_t)p.num_streams * nblocks + G::IN_ROWS - 1) / G::IN_ROWS
                              : (uint64_t)p.num_streams * ((nblocks + G::IN_ROWS - 1) / G::IN_ROWS);
  unsigned grid = 0;
  if (int rc = dec_persistent_grid(aad_decode_fast<BITS, C, IL, BULK, SPAN>, smem, warps, &grid)) return rc;
  aad_decode_fast<BITS, C, IL, BULK, SPAN><<<grid, kDecWarps * 32, smem, s>>>(p);
  return (int)cudaGetLastError();
}

/* Tasks that span streams (SPAN) cost a little per task (per-lane stream bookkeeping, shuffles instead of arithmetic
 * for the loader's row pointers: +1..3 % on batches that need none of it), so they are used where per-stream tasks
 * would idle at least 1 lane in 16: g_dec_span = 1 (default) by shape, 0 never, 2 always (tests). */
template <int BITS, int C, int IL, int BULK = 0>
int dec_fast_launch_bc(const aadk_decode_params &p, cudaStream_t s)
{
  using G = DecGeom<BITS, C>;
  const uint32_t nblocks = p.block_end - p.block_begin;
  const uint32_t per_stream = (nblocks + G::IN_ROWS - 1) / G::IN_ROWS * G::IN_ROWS;
  const bool wasteful = p.num_streams > 1u && (uint64_t)(per_stream - nblocks) * 16u >= per_stream;
  if (!BULK && (g_dec_span == 2 || (g_dec_span == 1 && wasteful))) return dec_fast_launch_as<BITS, C, IL, 0, 1>(p, s);
  return dec_fast_launch_as<BITS, C, IL, BULK, 0>(p, s);
}

/* the TMA flush needs every window of every block to start at a 16-byte aligned place of its PCM row */
inline bool dec_bulk_eligible(const aadk_decode_params &p)
{
  return g_dec_bulk != 0 && p.geo.bits == 4 && p.geo.channels == 1 && !p.interleaved && p.geo.samples_per_block % 8u == 0 &&
         ((uintptr_t)p.pcm & 15u) == 0 && p.pcm_clip_stride % 8u == 0 && p.sample_base % 8u == 0;
}

template <int BITS>
int dec_wide_launch(const aadk_decode_params &p, cudaStream_t s)
{
  const uint32_t C = p.geo.channels, rows = 32u / C;
  const size_t in_bytes = ((size_t)rows * 16u * ((BITS * C + 1u) | 1u) + 16u + 15u) & ~(size_t)15;
  const size_t smem = ((sizeof(DecTables) + 15) & ~(size_t)15) + (size_t)kDecWarps * (in_bytes + 32u * kDecOutPitch);
  auto kernel = p.interleaved ? aad_decode_wide<BITS, 1> : aad_decode_wide<BITS, 0>;
  const uint32_t nblocks = p.block_end - p.block_begin;
  const uint64_t warps = (uint64_t)p.num_streams * ((nblocks + rows - 1) / rows);
  unsigned grid = 0;
  if (int rc = dec_persistent_grid(kernel, smem, warps, &grid)) return rc;
  kernel<<<grid, kDecWarps * 32, smem, s>>>(p);
  return (int)cudaGetLastError();
}

/* aad_decode_tma: what the tensor map and the kernel's shortcuts need (everything else takes the default kernel) */
inline bool dec_tma_eligible(const aadk_decode_params &p)
{
  if (g_dec_tma == 0 || p.geo.channels != 1 || (p.geo.bits != 4 && p.geo.bits != 2) || p.interleaved) return false;
  if (p.sizes != nullptr || p.byte_base != 0 || p.uniform_samples == 0) return false;
  if (((uintptr_t)(p.aad + AADF_FILE_HEADER_BYTES) & 15u) || (p.aad_stride % 16u) || (p.geo.block_size % 16u)) return false;
  if (p.aad_stride >= (1ull << 40) || p.num_streams == 0) return false;
  /* every block a stream of at most uniform_samples samples can deliver lies wholly inside uniform_size (and the slot) */
  const uint64_t need = aadf_stream_bytes_bound(p.uniform_samples, p.geo.block_size, p.geo.samples_per_block);
  return need <= p.uniform_size && p.uniform_size <= p.aad_stride;
}

typedef CUresult (*dec_tma_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                      const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BITS, int HALF>
int dec_tma_launch(const aadk_decode_params &p, cudaStream_t s)
{
  using T = DecTmaGeom<BITS, HALF>;
  static std::atomic<dec_tma_encode_fn> encode_fn{nullptr};   /* contexts on different host threads launch concurrently */
  dec_tma_encode_fn encode = encode_fn.load();
  if (!encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    const cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return (int)(e != cudaSuccess ? e : cudaErrorNotSupported);
    encode = (dec_tma_encode_fn)fn;
    encode_fn.store(encode);
  }
  /* [stream][block][byte]: the whole blocks inside every stream's uniform_size bytes */
  const uint64_t rows = (p.uniform_size - AADF_FILE_HEADER_BYTES) / p.geo.block_size;
  const cuuint64_t dims[3] = { p.geo.block_size, rows, p.num_streams };
  const cuuint64_t strides[2] = { p.geo.block_size, p.aad_stride };
  const cuuint32_t box[3] = { (cuuint32_t)T::TB, 32u, 1u };
  const cuuint32_t estr[3] = { 1u, 1u, 1u };
  CUtensorMap tmap;
  const CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)(p.aad + AADF_FILE_HEADER_BYTES), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, T::TB == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return (int)cudaErrorInvalidValue;
  const uint32_t nblocks = p.block_end - p.block_begin;
  const uint64_t warps = (uint64_t)p.num_streams * ((nblocks + 31u) / 32u);
  /* one persistent CTA per SM (the shared memory of one fills it) */
  int dev = 0, sms = 148;
  if (int rc = device_sm_count(&dev, &sms)) return rc;
  if (int rc = allow_dynamic_smem(aad_decode_tma<BITS, HALF>, dev, T::SMEM)) return rc;
  const uint64_t ctas = (warps + T::WARPS - 1) / T::WARPS;
  const unsigned grid = (unsigned)(ctas > (uint64_t)sms ? (uint64_t)sms : ctas);
  aad_decode_tma<BITS, HALF><<<grid, T::WARPS * 32, T::SMEM, s>>>(p, tmap);
  g_tma_launches++;
  return (int)cudaGetLastError();
}

/* g_dec_wide_all (tests): 1 = mono and stereo streams also go through aad_decode_wide */
template <int BITS>
int dec_fast_launch(const aadk_decode_params &p, cudaStream_t s)
{
  if (p.geo.channels > 2 || (g_dec_wide_all && !p.interleaved)) return dec_wide_launch<BITS>(p, s);
  if (p.geo.channels == 1) {   /* mono: WAV order is the plane itself */
    if ((BITS == 4 || BITS == 2) && dec_tma_eligible(p))
      return g_dec_tma == 2 ? dec_tma_launch<(BITS == 3 ? 4 : BITS), 1>(p, s) : dec_tma_launch<(BITS == 3 ? 4 : BITS), 0>(p, s);
    if (BITS == 4 && dec_bulk_eligible(p)) return dec_fast_launch_bc<(BITS == 4 ? 4 : BITS), 1, 0, (BITS == 4 ? 1 : 0)>(p, s);
    return dec_fast_launch_bc<BITS, 1, 0>(p, s);
  }
  return p.interleaved ? dec_fast_launch_bc<BITS, 2, 1>(p, s) : dec_fast_launch_bc<BITS, 2, 0>(p, s);
}

}  // namespace
