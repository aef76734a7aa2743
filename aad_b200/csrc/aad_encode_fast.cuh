/*
 * aad_encode_fast.cuh -- the production encoder kernel (included by aad_kernels.cu).
 *
 * A chain = one (stream, channel) -- or, in the segment-parallel extension (aad_kernels.h:
 * segment_blocks), one (stream, segment, channel), a segment being a run of blocks that is encoded
 * like a stream of its own.  The reference carries the predictor weights and the step
 * index from block to block (processor[], src/aad_encoder.c:21,853-886), so a chain is serial over
 * its blocks; per block it runs the start-state search (src/aad_encoder.c:470-562: 1 + 2*trials
 * dry-run passes) and then the emitting pass (src/aad_encoder.c:565-727).  One thread per chain;
 * dry and emitting passes each have one copy of the 16-sample loop (enc_job_units<KIND>).  With few
 * chains (PAIR = 1) the baseline pass and the first trial pass of a block, which both start from the
 * carried state, run interleaved in the same thread (enc_run_pair).
 *
 * What makes a pass fast:
 *  - no integer divide and no shift around it: min((|diff| << (b-2)) / step, 2^(b-1) - 1) is
 *    min(umulhi(|diff|, D[step]), 2^(b-1) - 1) with D = ceil(2^(30+b) / step), exact wherever the clamp does
 *    not hide the quotient (tools/gen_tables.py).  D does not fit 32 bits for the first few steps
 *    (step <= 2^(b-2): "tiny" rows, near-silent signals): a 16-sample unit that could reach them takes the
 *    SAFE copy of the loop, which pre-shifts the operand there (EncQuant);
 *  - one 8-byte shared-memory load per sample fetches {step, D} from a table indexed directly
 *    by the Q4 step index (kept pre-multiplied by 8 in a register), the index update is one
 *    VIADDMNMX.RELU, the index-delta table is a register-resident byte LUT (3 PRMT);
 *  - short recurrence: |x - p| is one VABSDIFF, the sign of the difference is folded into the
 *    dequantiser's multiplier and rounding term so q follows the magnitude by IMAD + SHF only,
 *    the predictor sum is two multiply-add chains joined by one add;
 *  - the squared error is accumulated in FP64 like the reference does (exact: integers < 2^47),
 *    on the conversion / FP64 pipes this kernel otherwise leaves idle;
 *  - PCM is staged through a per-lane shared-memory ring with cp.async, three 16-sample units
 *    ahead of use; completion is tracked per commit group, so waiting for one unit never waits
 *    for the units requested after it (with plain loads the scoreboard made it so: 8 % of the
 *    v3 kernel's time, profiles/r01_v3_fast.md).
 *
 * Tried and dropped (profiles/r01_v4_encoder_experiments.md): speculative variants that ran the
 * EMITTING passes next to the trial passes (second lane, or second interleaved chain in the same
 * thread) to cut the critical path from 2 + 2*trials to 2*trials passes per block.  Bit-exact, but
 * not faster: they issue 8 chain-passes per block instead of 6, and each candidate start state
 * wins about a third of the blocks, so an emission run early is wasted two times out of three.
 */
#pragma once

namespace {

constexpr int kEncLutEntries = AADF_INDEX_MAX + 1;

typedef uint2 EncLutEntry;     /* x = (step << 16) | tiny, y = D (tiny rows: ceil(2^31 / step)) */
struct EncShared {
  EncLutEntry lut[kEncLutEntries];   /* indexed by stepsize_index */
};
constexpr uint32_t kEncLutBytes = (sizeof(EncShared) + 15u) & ~15u;
constexpr int kEncIdxScale = (int)sizeof(EncLutEntry);   /* the step index lives in a register as the table byte offset */

/* kEncIdxScale * (Q4 index delta) per magnitude code as a register-resident byte LUT: three PRMTs instead
 * of a shared-memory load on the step-index dependency chain (src/aad_tables.c:8-45). */
template <int BITS>
struct EncDelta {
  static __host__ __device__ constexpr int v(int k)
  {
    constexpr int d4[8] = AADK_DELTA4_INIT;
    constexpr int d3[4] = AADK_DELTA3_INIT;
    constexpr int d2[2] = AADK_DELTA2_INIT;
    return kEncIdxScale * (BITS == 4 ? d4[k & 7] : (BITS == 3 ? d3[k & 3] : d2[k & 1]));
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

__device__ const uint32_t g_step_direct[3][256] = {AADK_STEP_DIRECT2_INIT, AADK_STEP_DIRECT3_INIT, AADK_STEP_DIRECT4_INIT};

/* The shift-free quantiser (tools/gen_tables.py): rows [0, kTinyRows) of the step table need the operand pre-shifted
 * by b-1.  A 16-sample unit may use the copy of the loop without that (SAFE = 0) when the step index cannot come down
 * to a tiny row before the unit's last table lookup: index >= kFastIndex at its start (the index falls by at most
 * kMaxDrop per sample, src/aad_tables.c:8-45). */
template <int BITS>
struct EncQuant {
  static constexpr int kTinyRows = (BITS == 4) ? AADK_TINY_ROWS4 : (BITS == 3 ? AADK_TINY_ROWS3 : AADK_TINY_ROWS2);
  static constexpr int kMaxDrop = (BITS == 4) ? 18 : (BITS == 3 ? 16 : 14);
  static constexpr int kFastIndex = 16 * kTinyRows - 8 + 16 * kMaxDrop;   /* units are 16 samples (kEncUnit) */
};

template <int BITS>
__device__ __forceinline__ void enc_load_shared(EncShared &s)
{
  for (int i = threadIdx.x; i < kEncLutEntries; i += blockDim.x) {
    const int e = (i + 8) >> 4;
    s.lut[i] = make_uint2(((uint32_t)g_step_table[e] << 16) | (e < EncQuant<BITS>::kTinyRows ? 1u : 0u), g_step_direct[BITS - 2][e]);
  }
  __syncthreads();
}

/* what the reference keeps per channel between blocks: weight[4], stepsize_index */
struct EncState {
  int32_t w0, w1, w2, w3, idx8;   /* idx8 = kEncIdxScale * stepsize_index */
};

struct EncChain {
  int32_t h0, h1, h2, h3;   /* h0 newest */
  int32_t w0, w1, w2, w3;
  int32_t idx8;             /* kEncIdxScale * stepsize_index */
  EncLutEntry e;            /* the table entry of idx8, fetched as soon as idx8 is known (prime / the previous sample) */
  __device__ __forceinline__ void prime(const EncShared &s)
  {
    e = *reinterpret_cast<const EncLutEntry *>(reinterpret_cast<const char *>(s.lut) + idx8);
  }
  __device__ __forceinline__ void set(const EncState &s) { w0 = s.w0; w1 = s.w1; w2 = s.w2; w3 = s.w3; idx8 = s.idx8; }
  __device__ __forceinline__ EncState state() const { return EncState{w0, w1, w2, w3, idx8}; }
};

/* src/aad_encoder.c:343-410, one sample.  Returns the magnitude|sign code; q = signed dequantised diff.
 * SAFE = 0 only where the step index is known to stay clear of the tiny rows (EncQuant). */
template <int BITS, int SAFE = 1>
__device__ __forceinline__ uint32_t enc_sample(EncChain &c, int32_t x, const EncShared &s, int32_t &q)
{
  constexpr uint32_t kMaxMag = (1u << (BITS - 1)) - 1u;
  constexpr int kSh = BITS - 1;
  const EncLutEntry e = c.e;
  const int32_t step = (int32_t)(e.x >> 16), step2 = 2 * step;
  /* predictor: two independent multiply-add chains, joined by one add */
  int32_t acc_a, acc_b;   /* asm keeps the two chains apart (the compiler would re-serialise them) */
  asm("mad.lo.s32 %0, %1, %2, 16384;" : "=r"(acc_a) : "r"(c.h1), "r"(c.w1));
  asm("mul.lo.s32 %0, %1, %2;" : "=r"(acc_b) : "r"(c.h2), "r"(c.w2));
  asm("mad.lo.s32 %0, %1, %2, %0;" : "+r"(acc_a) : "r"(c.h0), "r"(c.w0));
  asm("mad.lo.s32 %0, %1, %2, %0;" : "+r"(acc_b) : "r"(c.h3), "r"(c.w3));
  const int32_t p = (int32_t)((uint32_t)acc_a + (uint32_t)acc_b) >> 15;
  /* |x| <= 2^15 and |p| <= 2^16: the difference cannot wrap, so sign and magnitude come from
   * a compare and a VABSDIFF instead of subtract + abs */
  const bool neg = x < p;
  uint32_t a = __sad(x, p, 0u);
  if (SAFE) a = (e.x & 1u) ? a << kSh : a;
  const uint32_t mag = min(__umulhi(a, e.y), kMaxMag);
  /* q = +-((step * (2 mag + 1)) >> kSh): for the negative branch -(v >> k) == (-v + 2^k - 1) >> k,
   * so the sign goes into the multiplier and the addend, both ready before mag is */
  const int32_t s2 = neg ? -step2 : step2;
  const int32_t s1 = neg ? ((1 << kSh) - 1) - step : step;
  q = ((int32_t)mag * s2 + s1) >> kSh;
  const int32_t r = max(__viaddmin_s32(q, p, 32767), -32768);
  c.w0 += (int32_t)((uint32_t)q * (uint32_t)c.h0 + (1u << 14)) >> 18;
  c.w1 += (int32_t)((uint32_t)q * (uint32_t)c.h1 + (1u << 14)) >> 18;
  c.w2 += (int32_t)((uint32_t)q * (uint32_t)c.h2 + (1u << 14)) >> 18;
  c.w3 += (int32_t)((uint32_t)q * (uint32_t)c.h3 + (1u << 14)) >> 18;
  c.idx8 = __viaddmin_s32_relu(c.idx8, EncDelta<BITS>::lookup(mag), kEncIdxScale * AADF_INDEX_MAX);
  c.prime(s);   /* the next sample's entry: its latency overlaps the rest of this sample, also across units */
  c.h3 = c.h2;
  c.h2 = c.h1;
  c.h1 = c.h0;
  c.h0 = r;
  return mag | (neg ? (1u << (BITS - 1)) : 0u);
}

/* sum += (double)(int32)(q * q): the reference squares in wrapping int32 and accumulates in double
 * (src/aad_encoder.c:461); every partial sum is an integer below 2^47, so the double sum is exact
 * and order independent.  I2F.F64 + DADD run on the conversion / FP64 pipes, which this kernel
 * leaves idle, instead of three more instructions on the saturated integer pipes. */
__device__ __forceinline__ void enc_add_square(double &sum, int32_t q)
{
  sum += (double)(int32_t)((uint32_t)q * (uint32_t)q);
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

constexpr int kEncUnit = 16;    /* samples per staged unit */
constexpr int kEncSlots = 4;    /* ring depth, in units */
constexpr int kEncAhead = 3;    /* units in flight ahead of the one being encoded */

/* Per-lane shared-memory ring of 16-sample units, filled by cp.async.  Layout per warp:
 * [slot][part][lane] of 16 bytes (part 0/1 = samples 0-7 / 8-15 of the a row, 2/3 = the b row
 * when mid/side is on), so a lane's LDS.128 and its 31 neighbours' are conflict free. */
template <int MS>
struct EncRing {
  static constexpr uint32_t kParts = MS ? 4u : 2u;
  static constexpr uint32_t kSlotBytes = kParts * 32u * 16u;
  static constexpr uint32_t kWarpBytes = (kEncSlots + 1) * kSlotBytes;   /* slot kEncSlots: a pass's last, partial unit */
  uint32_t base;   /* shared-space address of this lane's 16 bytes in slot 0, part 0 */

  static __device__ __forceinline__ void cp8(uint32_t dst, const void *src)
  {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
  }
  /* unit of 16 samples starting at sample i (i % 4 == 0, rows 8-byte aligned there) */
  __device__ __forceinline__ void issue(uint32_t slot, const EncSource<MS> &src, uint32_t i) const
  {
    const uint32_t d = base + slot * kSlotBytes;
    cp8(d, src.a + i);
    cp8(d + 8u, src.a + i + 4);
    cp8(d + 512u, src.a + i + 8);
    cp8(d + 520u, src.a + i + 12);
    if (MS && src.mode != 0) {
      cp8(d + 1024u, src.b + i);
      cp8(d + 1032u, src.b + i + 4);
      cp8(d + 1536u, src.b + i + 8);
      cp8(d + 1544u, src.b + i + 12);
    }
  }
  /* the first `pieces` (1..3) 4-sample pieces of the unit that starts at sample i: a pass's last, partial unit */
  __device__ __forceinline__ void issue_pieces(uint32_t slot, const EncSource<MS> &src, uint32_t i, uint32_t pieces) const
  {
    const uint32_t d = base + slot * kSlotBytes;
    const bool both = MS && src.mode != 0;
    cp8(d, src.a + i);
    if (both) cp8(d + 1024u, src.b + i);
    if (pieces > 1u) {
      cp8(d + 8u, src.a + i + 4);
      if (both) cp8(d + 1032u, src.b + i + 4);
    }
    if (pieces > 2u) {
      cp8(d + 512u, src.a + i + 8);
      if (both) cp8(d + 1536u, src.b + i + 8);
    }
  }
  static __device__ __forceinline__ void commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
  template <int N>
  static __device__ __forceinline__ void wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
  static __device__ __forceinline__ uint4 lds128(uint32_t addr)
  {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
  }
};

/* the 16 samples of a unit, two packed int16 per register */
template <int MS>
struct EncUnit {
  uint32_t a[8], b[8];
  __device__ __forceinline__ void read(const EncRing<MS> &ring, uint32_t slot, const EncSource<MS> &src)
  {
    const uint32_t d = ring.base + slot * EncRing<MS>::kSlotBytes;
    const uint4 v0 = EncRing<MS>::lds128(d), v1 = EncRing<MS>::lds128(d + 512u);
    a[0] = v0.x; a[1] = v0.y; a[2] = v0.z; a[3] = v0.w; a[4] = v1.x; a[5] = v1.y; a[6] = v1.z; a[7] = v1.w;
    if (MS) {
      if (src.mode != 0) {
        const uint4 u0 = EncRing<MS>::lds128(d + 1024u), u1 = EncRing<MS>::lds128(d + 1536u);
        b[0] = u0.x; b[1] = u0.y; b[2] = u0.z; b[3] = u0.w; b[4] = u1.x; b[5] = u1.y; b[6] = u1.z; b[7] = u1.w;
      } else {
#pragma unroll
        for (int k = 0; k < 8; k++) b[k] = 0u;
      }
    }
  }
  __device__ __forceinline__ int32_t get(const EncSource<MS> &src, int j) const
  {
    const uint32_t wa = a[j >> 1];
    const int32_t x = (j & 1) ? ((int32_t)wa >> 16) : (int32_t)(int16_t)(wa & 0xFFFFu);
    if (!MS) return x;
    const uint32_t wb = b[j >> 1];
    const int32_t y = (j & 1) ? ((int32_t)wb >> 16) : (int32_t)(int16_t)(wb & 0xFFFFu);
    return src.combine(x, y);
  }
};

/* history[3 - i] = x[i], missing taps zero (src/aad_encoder.c:606-616, :450-453) */
template <int MS>
__device__ __forceinline__ void enc_load_history(EncChain &c, const EncSource<MS> &src, uint32_t first, uint32_t n)
{
  c.h3 = (n > 0) ? src.at(first) : 0;
  c.h2 = (n > 1) ? src.at(first + 1) : 0;
  c.h1 = (n > 2) ? src.at(first + 2) : 0;
  c.h0 = (n > 3) ? src.at(first + 3) : 0;
}

/* Per-lane output stream for the mono case, where a chain's code bytes are contiguous: bytes are
 * collected in a 64-bit register and leave as aligned 4-byte stores (single bytes only up to the
 * first 4-byte boundary and for the last partial word): one store per 8 samples instead of four
 * lane-strided byte stores. */
struct EncByteStream {
  uint8_t *ptr;          /* 4-byte aligned: where the low word of acc goes */
  unsigned long long acc;
  uint32_t cnt;          /* bytes held in acc, counting the `skip` bytes in front of the first word */
  uint32_t skip;         /* bytes of the next word store that lie before the stream's start */
  __device__ __forceinline__ void begin(uint8_t *p)
  {
    skip = (uint32_t)((uintptr_t)p & 3u);
    ptr = p - skip;
    acc = 0ull;
    cnt = skip;
  }
  __device__ __forceinline__ void flush_word()
  {
    if (skip == 0u) {
      *reinterpret_cast<uint32_t *>(ptr) = (uint32_t)acc;
    } else {   /* once per block: the word that straddles the stream's start */
      for (uint32_t k = skip; k < 4u; k++) ptr[k] = (uint8_t)(acc >> (8u * k));
      skip = 0u;
    }
    ptr += 4; acc >>= 32; cnt -= 4u;
  }
  __device__ __forceinline__ void put(uint32_t value, uint32_t nbytes)   /* value: first byte in bits 0-7; nbytes <= 4 */
  {
    acc |= (unsigned long long)value << (8u * cnt);
    cnt += nbytes;
    if (cnt >= 4u) flush_word();
  }
  __device__ __forceinline__ void end()
  {
    for (uint32_t k = skip; k < cnt; k++) ptr[k] = (uint8_t)(acc >> (8u * k));
    cnt = 0u;
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

/*
 * A job = one pass of one chain over the samples [first, first + n) of its row, starting from the
 * weights / step index in c; `sum` collects the wrapped squares of the dequantised differences
 * (exact: |term| < 2^31, n < 2^16).
 *   emit == false  dry run, src/aad_encoder.c:431-467: exactly n samples; with n < 4 nothing
 *                  happens at all (state untouched).
 *   emit == true   src/aad_encoder.c:565-727 for this channel: weights masked to what the block
 *                  header can carry, header written at blk, codes written in whole groups, the
 *                  last group zero padded past n.
 * run == false makes the job a no-op.
 */
template <int MS>
struct EncJob {
  EncChain c;
  uint32_t first, n;
  bool run, emit;
  uint8_t *blk;
  double sum;
  /* progress */
  uint32_t units;        /* full 16-sample units after the 4 history samples */
  uint32_t tail4;        /* whole 4-sample pieces after the last full unit (0..3): they come through the ring too */
  uint8_t *dp;           /* next code group (channel-interleaved streams) */
  EncByteStream bytes;   /* next code bytes (mono streams) */
  EncRing<MS> ring;
};

template <int BITS, int MS>
__device__ __forceinline__ void enc_job_begin(EncJob<MS> &j, const EncSource<MS> &src, uint32_t ch, uint32_t C, const EncShared &sh)
{
  constexpr uint32_t GB = (BITS == 3) ? 3 : 1;
  j.sum = 0.0;
  j.units = 0;
  j.tail4 = 0;
  if (!j.emit && j.n < AADF_TAPS) j.run = false;
  j.dp = j.blk + C * AADF_CHANNEL_HEADER_BYTES + ch * GB;
  j.bytes.begin(j.dp);
  if (!j.run) return;
  EncChain &c = j.c;
  c.prime(sh);
  enc_load_history<MS>(c, src, j.first, j.n);
  if (j.emit) {   /* block header, src/aad_encoder.c:619-655 */
    int32_t maxabs = 0;
    maxabs = max(maxabs, c.w0 >= 0 ? c.w0 : (int32_t)(0u - (uint32_t)c.w0));
    maxabs = max(maxabs, c.w1 >= 0 ? c.w1 : (int32_t)(0u - (uint32_t)c.w1));
    maxabs = max(maxabs, c.w2 >= 0 ? c.w2 : (int32_t)(0u - (uint32_t)c.w2));
    maxabs = max(maxabs, c.w3 >= 0 ? c.w3 : (int32_t)(0u - (uint32_t)c.w3));
    uint32_t shift = 0;
    while (maxabs > 32767) { maxabs >>= 1; shift++; }
    const int32_t keep = (int32_t)~((1u << shift) - 1u);
    c.w0 &= keep; c.w1 &= keep; c.w2 &= keep; c.w3 &= keep;
    uint8_t *hp = j.blk + ch * AADF_CHANNEL_HEADER_BYTES;
    aadf_put_be16(hp, ((((uint32_t)c.idx8 / kEncIdxScale) << 4) | (shift & 0xFu)) & 0xFFFFu);
    aadf_put_be16(hp + 2, (uint32_t)(c.w0 >> shift) & 0xFFFFu);  aadf_put_be16(hp + 4, (uint32_t)c.h0 & 0xFFFFu);
    aadf_put_be16(hp + 6, (uint32_t)(c.w1 >> shift) & 0xFFFFu);  aadf_put_be16(hp + 8, (uint32_t)c.h1 & 0xFFFFu);
    aadf_put_be16(hp + 10, (uint32_t)(c.w2 >> shift) & 0xFFFFu); aadf_put_be16(hp + 12, (uint32_t)c.h2 & 0xFFFFu);
    aadf_put_be16(hp + 14, (uint32_t)(c.w3 >> shift) & 0xFFFFu); aadf_put_be16(hp + 16, (uint32_t)c.h3 & 0xFFFFu);
  }
  j.units = ((j.n > AADF_TAPS) ? j.n - AADF_TAPS : 0u) / kEncUnit;
  j.tail4 = (((j.n > AADF_TAPS) ? j.n - AADF_TAPS : 0u) % kEncUnit) / 4u;
  /* the samples after the last whole 4-sample piece are read one by one at the end: pull their line into L1 now */
  const uint32_t tail = j.first + AADF_TAPS + j.units * kEncUnit + 4u * j.tail4;
  /* the whole 4-sample pieces behind the last full unit: requested now, into a slot of their own, read at the very end
   * (they join the pass's first commit group) */
  if (j.tail4 != 0u) j.ring.issue_pieces(kEncSlots, src, j.first + AADF_TAPS + j.units * kEncUnit, j.tail4);
  if (tail < j.first + j.n) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(src.a + tail));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(src.a + j.first + j.n - 1));
    if (MS && src.mode != 0) {
      asm volatile("prefetch.global.L1 [%0];" ::"l"(src.b + tail));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(src.b + j.first + j.n - 1));
    }
  }
}

/* request unit u of the job (a no-op past its last full unit) */
template <int MS>
__device__ __forceinline__ void enc_job_request(const EncJob<MS> &j, const EncSource<MS> &src, uint32_t u)
{
  if (u < j.units) j.ring.issue(u & (kEncSlots - 1), src, j.first + AADF_TAPS + u * kEncUnit);
}

/* a unit's codes, two words of 8 codes each (first sample in the most significant bits), to the stream:
 * src/aad_encoder.c:661-722 */
template <int BITS, int MS>
__device__ __forceinline__ void enc_job_put_unit(EncJob<MS> &j, const uint32_t half[2], bool mono, uint32_t gstride)
{
  constexpr uint32_t GS = (BITS == 4) ? 2 : (BITS == 3 ? 8 : 4);
  constexpr uint32_t GB = (BITS == 3) ? 3 : 1;
  if (mono) {
    /* stream order = most significant byte first: byte-reverse, then feed little-endian */
#pragma unroll
    for (int h = 0; h < 2; h++) {
      if (BITS == 4) j.bytes.put(__byte_perm(half[h], 0u, 0x0123), 4u);
      else if (BITS == 3) j.bytes.put(__byte_perm(half[h], 0u, 0x4012), 3u);
      else j.bytes.put(__byte_perm(half[h], 0u, 0x4401), 2u);
    }
  } else {
    constexpr int kGroups = 8 / (int)GS;   /* groups per 8 codes: 4 / 1 / 2 */
#pragma unroll
    for (int h = 0; h < 2; h++) {
#pragma unroll
      for (int g = 0; g < kGroups; g++) {
        enc_store_group<BITS>(j.dp, (half[h] >> (8 * GB * (kGroups - 1 - g))) & ((1u << (8 * GB)) - 1u));
        j.dp += gstride;
      }
    }
  }
}

/* one full 16-sample unit u of a job: wait for its samples, request the unit kEncAhead further on, run the 16 samples.
 * KIND 0: dry run (error sum only), 1: emitting (codes only): two copies of the loop, 3 instructions per sample
 * shorter each than one that does both.  SAFE: see EncQuant.  The whole unit is ONE basic block on purpose: the ring
 * bookkeeping and the next unit's request issue in the stall slots of the sample recurrence. */
template <int BITS, int MS, int KIND, int SAFE>
__device__ __forceinline__ void enc_job_unit(EncJob<MS> &j, const EncSource<MS> &src, uint32_t u, bool mono, uint32_t gstride,
                                             const EncShared &sh)
{
  EncChain &c = j.c;
  EncRing<MS>::template wait<kEncAhead - 1>();
  EncUnit<MS> cur;
  cur.read(j.ring, u & (kEncSlots - 1), src);
  enc_job_request<MS>(j, src, u + kEncAhead);
  EncRing<MS>::commit();
  uint32_t half[2] = {0u, 0u};
#pragma unroll
  for (int k = 0; k < kEncUnit; k++) {
    int32_t q;
    const uint32_t code = enc_sample<BITS, SAFE>(c, cur.get(src, k), sh, q);
    if (KIND != 0) half[k >> 3] = (half[k >> 3] << BITS) + code;
    if (KIND != 1) enc_add_square(j.sum, q);
  }
  if (KIND == 1) enc_job_put_unit<BITS, MS>(j, half, mono, gstride);
}

/* the full 16-sample units [u0, units) of one job.
 * BY2: the emitting passes of aad_encode_fast<..., PAIR = 0> take two units per turn while the step index stays clear of
 * the tiny rows for 32 samples.  Measured, tools/enc_variants.py: 12,500 chains at 0 trials (one warp per scheduler,
 * emitting passes only) 24.4 -> 21.2 ms; no change with 50,000 chains; but the helper-lane and pairing kernels LOSE with
 * it (98.6 -> 104.0 ms emitting passes only, -> 118.5 ms dry passes too: the slowest warps take 19 % longer while the
 * average takes 3 %), so nothing else uses it. */
template <int BITS, int MS, int KIND, int BY2>
__device__ __forceinline__ void enc_job_units(EncJob<MS> &j, const EncSource<MS> &src, uint32_t u0, bool mono,
                                              uint32_t gstride, const EncShared &sh)
{
  uint32_t u = u0;
  if (BY2 && KIND == 1) {
    for (; u + 1u < j.units && j.c.idx8 >= kEncIdxScale * (EncQuant<BITS>::kFastIndex + 16 * EncQuant<BITS>::kMaxDrop); u += 2u) {
      enc_job_unit<BITS, MS, KIND, 0>(j, src, u, mono, gstride, sh);
      enc_job_unit<BITS, MS, KIND, 0>(j, src, u + 1u, mono, gstride, sh);
    }
  }
  for (; u < j.units; u++) {
    if (j.c.idx8 >= kEncIdxScale * EncQuant<BITS>::kFastIndex) enc_job_unit<BITS, MS, KIND, 0>(j, src, u, mono, gstride, sh);
    else enc_job_unit<BITS, MS, KIND, 1>(j, src, u, mono, gstride, sh);   /* near silence: the index may reach the tiny rows */
  }
}

/* units [u0, units) of one job, then its last samples one by one; the emitting pass rounds up
 * to whole groups with zero samples (src/aad_encoder.c:592-593).  Units u0 .. u0+kEncAhead-1
 * have been requested already. */
template <int BITS, int MS, int BY2 = 0>
__device__ __forceinline__ void enc_job_finish(EncJob<MS> &j, const EncSource<MS> &src, uint32_t u0, uint32_t C,
                                               const EncShared &sh)
{
  constexpr uint32_t GS = (BITS == 4) ? 2 : (BITS == 3 ? 8 : 4);
  constexpr uint32_t GB = (BITS == 3) ? 3 : 1;
  if (!j.run) return;
  const bool mono = (C == 1);                      /* uniform: contiguous code bytes -> word stores */
  const uint32_t gstride = C * GB;
  EncChain &c = j.c;
  if (j.emit) enc_job_units<BITS, MS, 1, BY2>(j, src, u0, mono, gstride, sh);
  else enc_job_units<BITS, MS, 0, BY2>(j, src, u0, mono, gstride, sh);
  const uint32_t total = (j.n > AADF_TAPS) ? j.n - AADF_TAPS : 0u;
  const uint32_t limit = j.first + j.n;
  const uint32_t rest = total - j.units * kEncUnit;
  const uint32_t cnt = j.emit ? (rest + GS - 1u) / GS * GS : rest;
  uint32_t i = j.first + AADF_TAPS + j.units * kEncUnit;
  uint32_t packed = 0, inpack = 0;
  auto one = [&](int32_t xs) {
    int32_t q;
    packed = (packed << BITS) | enc_sample<BITS>(c, xs, sh, q);
    enc_add_square(j.sum, q);
    if (j.emit && ++inpack == GS) {
      if (mono) {
        if (BITS == 3) j.bytes.put(__byte_perm(packed, 0u, 0x4012), 3u);
        else j.bytes.put(packed & 0xFFu, 1u);
      } else {
        enc_store_group<BITS>(j.dp, packed);
        j.dp += gstride;
      }
      packed = 0;
      inpack = 0;
    }
  };
  uint32_t k = 0;
  if (j.tail4 != 0u) {   /* the whole 4-sample pieces behind the last full unit (requested in enc_job_begin) */
    EncRing<MS>::template wait<0>();
    EncUnit<MS> cur;
    cur.read(j.ring, kEncSlots, src);
#pragma unroll
    for (int g = 0; g < 3; g++) {
      if ((uint32_t)g < j.tail4) {
#pragma unroll
        for (int t = 0; t < 4; t++) one(cur.get(src, 4 * g + t));
      }
    }
    k = 4u * j.tail4;
    i += k;
  }
  for (; k < cnt; k++, i++) one((i < limit) ? src.at(i) : 0);
  if (j.emit && mono) j.bytes.end();
}

/* one whole pass: begin, request the first units, run */
template <int BITS, int MS, int BY2 = 0>
__device__ __forceinline__ void enc_run_job(EncJob<MS> &j, const EncSource<MS> &src, uint32_t ch, uint32_t C,
                                            const EncShared &sh)
{
  enc_job_begin<BITS, MS>(j, src, ch, C, sh);
#pragma unroll
  for (int d = 0; d < kEncAhead; d++) {
    enc_job_request<MS>(j, src, d);
    EncRing<MS>::commit();
  }
  enc_job_finish<BITS, MS, BY2>(j, src, 0u, C, sh);
}

/* one 16-sample unit of two DRY passes interleaved instruction by instruction (enc_run_pair) */
template <int BITS, int MS, int SAFE>
__device__ __forceinline__ void enc_pair_unit(EncJob<MS> &x, EncJob<MS> &y, const EncSource<MS> &src, uint32_t u, const EncShared &sh)
{
  EncRing<MS>::template wait<kEncAhead - 1>();
  EncUnit<MS> ux, uy;
  ux.read(x.ring, u & (kEncSlots - 1), src);
  uy.read(y.ring, u & (kEncSlots - 1), src);
  enc_job_request<MS>(x, src, u + kEncAhead);
  enc_job_request<MS>(y, src, u + kEncAhead);
  EncRing<MS>::commit();
#pragma unroll
  for (int k = 0; k < kEncUnit; k++) {
    int32_t qx, qy;
    (void)enc_sample<BITS, SAFE>(x.c, ux.get(src, k), sh, qx);
    (void)enc_sample<BITS, SAFE>(y.c, uy.get(src, k), sh, qy);
    enc_add_square(x.sum, qx);
  }
}

/* Two DRY passes of the same thread at once (the baseline pass and the first trial pass of a block
 * both start from the carried state, so they are independent).  While both have full units left
 * their sample recurrences are interleaved instruction by instruction: one fills the other's
 * dependency stalls (a lone chain issues on ~45 % of its cycles), ~130 cycles per sample pair
 * instead of 2 x 95.  x keeps its error sum; y only its state.  What is left of either is finished
 * alone. */
template <int BITS, int MS>
__device__ __forceinline__ void enc_run_pair(EncJob<MS> &x, EncJob<MS> &y, const EncSource<MS> &src, uint32_t ch,
                                             uint32_t C, const EncShared &sh)
{
  enc_job_begin<BITS, MS>(x, src, ch, C, sh);
  enc_job_begin<BITS, MS>(y, src, ch, C, sh);
#pragma unroll
  for (int d = 0; d < kEncAhead; d++) {
    enc_job_request<MS>(x, src, d);
    enc_job_request<MS>(y, src, d);
    EncRing<MS>::commit();
  }
  const uint32_t common = (x.run && y.run) ? min(x.units, y.units) : 0u;
  for (uint32_t u = 0; u < common; u++) {
    if (min(x.c.idx8, y.c.idx8) >= kEncIdxScale * EncQuant<BITS>::kFastIndex) enc_pair_unit<BITS, MS, 0>(x, y, src, u, sh);
    else enc_pair_unit<BITS, MS, 1>(x, y, src, u, sh);
  }
#pragma unroll 1
  for (int which = 0; which < 2; which++) enc_job_finish<BITS, MS>(which ? y : x, src, common, C, sh);
}

/* src/aad_encoder.c:465: sqrt(sum / n) in double; 0 for n < 4 (src/aad_encoder.c:444-447) */
__device__ __forceinline__ double enc_rmse(double sum, uint32_t n)
{
  return (n < AADF_TAPS) ? 0.0 : sqrt(sum / (double)n);
}

/* PAIR: run the two independent dry passes of a block interleaved in one thread (enc_run_pair).
 * Pays when chains are scarce and every warp has a scheduler to itself; costs registers (a second
 * chain and sample unit), so launches with many chains use PAIR = 0. */
template <int BITS, int MS, int PAIR>
__global__ void __launch_bounds__(256) aad_encode_fast(const aadk_encode_params p)
{
  extern __shared__ __align__(16) unsigned char enc_smem[];
  EncShared &sh = *reinterpret_cast<EncShared *>(enc_smem);
  enc_load_shared<BITS>(sh);

  const uint32_t C = p.geo.channels;
  const uint32_t spb = p.geo.samples_per_block;
  const uint32_t bs = p.geo.block_size;

  /* chain = (stream, segment, channel); one segment spanning the stream unless segment mode is on */
  const uint32_t segs = p.segment_blocks ? p.num_segments : 1u;
  const uint64_t chain = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stream = chain / ((uint64_t)C * segs);
  if (stream >= p.num_streams) return;
  const uint32_t seg = (uint32_t)(chain / C % segs);
  const uint32_t ch = (uint32_t)(chain % C);
  const uint32_t ns = p.num_samples ? p.num_samples[stream] : p.uniform_samples;

  uint8_t *out = p.aad + stream * p.aad_stride;
  if (ch == 0 && seg == 0 && p.block_begin == 0 && p.byte_base == 0) {
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

  const uint64_t st = chain * AADK_STATE_WORDS;
  EncState S;   /* the carried state */
  S.w0 = p.state_in ? p.state_in[st + 0] : 0;
  S.w1 = p.state_in ? p.state_in[st + 1] : 0;
  S.w2 = p.state_in ? p.state_in[st + 2] : 0;
  S.w3 = p.state_in ? p.state_in[st + 3] : 0;
  S.idx8 = kEncIdxScale * (p.state_in ? p.state_in[st + 4] : 0);

  const uint32_t trials = p.trials;
  /* this chain's blocks: [seg_first, seg_end) cut to the launch's block range */
  const uint32_t seg_first = p.segment_blocks ? seg * p.segment_blocks : 0u;
  const uint32_t seg_end = p.segment_blocks ? seg_first + p.segment_blocks : 0xFFFFFFFFu;
  /* the launch's block range counts from the stream's first block, or (segment_relative) from every segment's */
  const uint32_t rel = (p.segment_blocks && p.segment_relative) ? seg_first : 0u;
  if (p.segment_blocks && p.segment_relative && p.segment_end != 0u && (seg < p.segment_begin || seg >= p.segment_end))
    return;   /* another device's segment */
  const uint32_t nblk = min(min(aadf_num_blocks(ns, spb), (uint32_t)min((uint64_t)rel + p.block_end, (uint64_t)0xFFFFFFFFu)), seg_end);
  const uint32_t b_begin = max(rel + p.block_begin, seg_first);
  if (p.segment_blocks && b_begin == seg_first) S.w0 = S.w1 = S.w2 = S.w3 = S.idx8 = 0;   /* a segment starts like a new stream */

  EncJob<MS> job, job2;
  job.ring.base = (uint32_t)__cvta_generic_to_shared(enc_smem) + kEncLutBytes +
                  (threadIdx.x >> 5) * ((PAIR ? 2u : 1u) * EncRing<MS>::kWarpBytes) + (threadIdx.x & 31u) * 16u;
  job2.ring.base = job.ring.base + EncRing<MS>::kWarpBytes;

  for (uint32_t b = b_begin; b < nblk; b++) {
    const uint32_t n = min(spb, ns - b * spb);
    const uint32_t first = b * spb - (uint32_t)p.sample_base;   /* the rows of p.pcm start at sample sample_base */
    /* src/aad_encoder.c:470-562 then :565-727, as one loop over passes (a single copy of the
     * sample loop in the instruction stream):
     *   pass 0        baseline: current block from the carried state
     *   pass 2t+1     trial t: previous block (skipped in the first block), state keeps running
     *   pass 2t+2     trial t: snapshot = candidate, current block, keep candidate if strictly better
     *   last pass     emit from the best start state */
    EncState run = S, best = S, cand = S;
    double best_rmse = 0.0;
    const uint32_t dry = trials ? 1u + 2u * trials : 0u;
    uint32_t k = 0;
    if (PAIR && dry && b > seg_first) {   /* passes 0 and 1 both start from the carried state: run them together */
      job.c.set(S);  job.first = first;        job.n = n;    job.run = true;  job.emit = false;
      job2.c.set(S); job2.first = first - spb; job2.n = spb; job2.run = true; job2.emit = false;
      job.blk = job2.blk = out;
      enc_run_pair<BITS, MS>(job, job2, src, ch, C, sh);
      best_rmse = enc_rmse(job.sum, n);
      if (job2.run) run = job2.c.state();
      k = 2;
    }
    for (; k <= dry; k++) {
      const bool emit = (k == dry);
      const bool on_prev = !emit && (k & 1u) != 0u;
      if (on_prev && b == seg_first) continue;
      if (!emit && !on_prev && k > 0) cand = run;
      job.c.set(emit ? best : run);
      job.first = on_prev ? first - spb : first;
      job.n = on_prev ? spb : n;
      job.run = true;
      job.emit = emit;
      job.blk = out + (AADF_FILE_HEADER_BYTES + (uint64_t)b * bs - p.byte_base);   /* p.aad points at byte byte_base */
      enc_run_job<BITS, MS, PAIR ? 0 : 1>(job, src, ch, C, sh);
      if (emit) {
        S = job.c.state();
      } else {
        if (job.run) run = job.c.state();
        const double rmse = enc_rmse(job.sum, job.n);
        if (k == 0) {
          best_rmse = rmse;
          run = S;                 /* the trial chain restarts from the carried state */
        } else if (!on_prev && best_rmse > rmse) {   /* NaN compares false, like the reference */
          best_rmse = rmse;
          best = cand;
        }
      }
    }
  }

  if (p.state_out && (!p.segment_blocks || b_begin < nblk)) {
    p.state_out[st + 0] = S.w0;
    p.state_out[st + 1] = S.w1;
    p.state_out[st + 2] = S.w2;
    p.state_out[st + 3] = S.w3;
    p.state_out[st + 4] = S.idx8 / kEncIdxScale;
  }
}

/* rows must be 8-byte aligned wherever a unit is staged */
inline bool enc_fast_eligible(const aadk_encode_params &p)
{
  if (p.geo.samples_per_block % 4u) return false;
  if (((uintptr_t)p.pcm & 7u) || (p.pcm_clip_stride % 4u) || (p.pcm_ch_stride % 4u) || (p.sample_base % 4u)) return false;
  return true;
}

template <int BITS, int MS, int PAIR>
int enc_fast_launch_as(const aadk_encode_params &p, cudaStream_t s)
{
  const uint64_t lanes = (uint64_t)p.num_streams * p.geo.channels * (p.segment_blocks ? p.num_segments : 1u);
  const size_t ring_bytes = (size_t)(PAIR ? 2 : 1) * EncRing<MS>::kWarpBytes;
  int dev = 0, sms = 148;
  if (int rc = device_sm_count(&dev, &sms)) return rc;
  if (int rc = allow_dynamic_smem(aad_encode_fast<BITS, MS, PAIR>, dev, kEncLutBytes + 8 * ring_bytes)) return rc;
  /* The step table is per CTA.  Smallest CTA that keeps every chain resident in one wave, so few
   * chains spread as single warps over all SMs and sub-partitions; otherwise the shape with the
   * most resident threads. */
  unsigned block = 0, best_block = 32;
  uint64_t best_resident = 0;
  for (unsigned cand = 32; cand <= 256 && block == 0; cand *= 2) {
    int per_sm = 0;
    if (int rc = resident_ctas(aad_encode_fast<BITS, MS, PAIR>, dev, (int)cand, kEncLutBytes + (cand / 32) * ring_bytes, &per_sm)) return rc;
    const uint64_t resident = (uint64_t)per_sm * sms * cand;
    if (resident >= lanes) block = cand;
    if (resident > best_resident) { best_resident = resident; best_block = cand; }
  }
  if (block == 0) block = best_block;
  const unsigned grid = (unsigned)((lanes + block - 1) / block);
  const size_t smem = kEncLutBytes + (size_t)(block / 32) * ring_bytes;
  aad_encode_fast<BITS, MS, PAIR><<<grid, block, smem, s>>>(p);
  return (int)cudaGetLastError();
}

}  // namespace
