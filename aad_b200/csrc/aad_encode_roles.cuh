/*
 * aad_encode_roles.cuh -- the encoder's schedule for SCARCE chains (included by aad_kernels.cu after
 * aad_encode_fast.cuh, whose per-sample code, sample ring and pass machinery it reuses unchanged).
 *
 * The reference's per-block schedule (src/aad_encoder.c:470-562, then :565-727) with t trials is
 *
 *     baseline(cur) | prev_1 -> cur_1 -> prev_2 -> cur_2 ... | decide | emit(best)
 *
 * of which only prev_1 -> cur_1 -> ... -> cur_t -> decide -> emit is a dependency chain: the baseline pass
 * starts from the carried state like prev_1 does, and the emitting pass of every candidate start state
 * (the carried state, and the state before each cur_k) could start as soon as that state exists.  A GPU
 * thread advances one pass by one sample every ~95 cycles whatever its neighbours do, so with few chains
 * the time of the kernel is (passes on the chain) x samples x 95 cycles -- and passes that are OFF the
 * chain are free if they run in lanes (or warps) that would otherwise idle:
 *
 *  roles, SPEC = 0   chains <= 25 per warp scheduler of the device (12,500 ten-second clips on a B200):
 *      every warp holds `ka` chain lanes and ceil(ka / 2t) helper lanes.  A helper lane runs the BASELINE pass
 *      of one of its chain lanes per dry slot, so the chain lanes run 2t dry slots + the emitting slot:
 *      5 slots per block instead of 6 (5.4 with the pass pairing of aad_encode_fast).  All lanes of a
 *      warp are in the same kind of pass at any time (dry, or emitting): no divergence.
 *
 *  roles, SPEC = 1   a handful of chains (ONE long stream: 2 or 8 chains, BASELINE configs[2], [3]):
 *      a second warp of the CTA runs the EMITTING pass of every candidate start state while the first
 *      warp is still searching (into shared-memory scratch blocks; the winner's bytes are committed to
 *      the stream when the search has decided): 2t slots per block instead of 2 + 2t.
 *
 * Nothing about the arithmetic changes: every pass is the same enc_run_job as in aad_encode_fast, only
 * WHERE and WHEN it runs differs, so the bytes are the reference's (tests/test_gpu_parity.py::
 * test_encoder_schedules_do_not_change_a_byte, the golden fixtures, bench.py's parity legs).
 */
#pragma once

namespace {

constexpr int kRolesMaxTrials = 2;             /* candidates kept in registers */
constexpr int kRolesWarps = 4;                 /* SPEC = 0: warps per CTA, one per scheduler of an SM */
constexpr int kRolesSpecStreams = 4;           /* SPEC = 1: streams per CTA at most */
constexpr int kRolesSpecChains = 16;           /* SPEC = 1: chains per CTA at most */

struct RolesChain {
  uint64_t stream;
  uint32_t ch;
  uint32_t ns;
};

template <int MS>
__device__ __forceinline__ RolesChain roles_locate(const aadk_encode_params &p, uint64_t chain, EncSource<MS> &src)
{
  const uint32_t C = p.geo.channels;
  RolesChain r;
  r.stream = chain / C;
  r.ch = (uint32_t)(chain % C);
  r.ns = p.num_samples ? p.num_samples[r.stream] : p.uniform_samples;
  const int16_t *base = (const int16_t *)p.pcm + r.stream * p.pcm_clip_stride;
  const bool pair = MS && r.ch < 2;
  src.a = base + (uint64_t)(pair ? 0 : r.ch) * p.pcm_ch_stride;
  src.b = base + p.pcm_ch_stride;
  src.mode = pair ? (r.ch == 0 ? 1 : 2) : 0;
  return r;
}

__device__ __forceinline__ EncState roles_shfl_state(const EncState &s, uint32_t lane)
{
  EncState r;
  r.w0 = __shfl_sync(0xFFFFFFFFu, s.w0, lane);
  r.w1 = __shfl_sync(0xFFFFFFFFu, s.w1, lane);
  r.w2 = __shfl_sync(0xFFFFFFFFu, s.w2, lane);
  r.w3 = __shfl_sync(0xFFFFFFFFu, s.w3, lane);
  r.idx8 = __shfl_sync(0xFFFFFFFFu, s.idx8, lane);
  return r;
}

/* what the second warp needs to know about a candidate start state, and what it reports back */
struct RolesMail {
  EncState cand[2][kRolesSpecChains];    /* start state of the emitting pass to run in the next slot (by slot parity) */
  EncState after[1 + kRolesMaxTrials][kRolesSpecChains];   /* chain state after the emitting pass from candidate k */
  int32_t winner[kRolesSpecChains];      /* 0 = carried state, k = state before cur_k */
};

/*
 * SPEC = 0: grid of kRolesWarps-warp CTAs, every warp on its own: lanes [0, ka) = chains gw * ka + lane,
 *           lanes [ka, ka + ceil(ka / per)) = helpers, per = 2 * trials.
 * SPEC = 1: 2-warp CTAs: warp 0 as above for the CTA's chains, warp 1 lane i = emitting passes of chain i.
 */
template <int BITS, int MS, int SPEC>
__global__ void __launch_bounds__(SPEC ? 64 : kRolesWarps * 32) aad_encode_roles(const aadk_encode_params p, uint32_t ka)
{
  extern __shared__ __align__(16) unsigned char enc_smem[];
  EncShared &sh = *reinterpret_cast<EncShared *>(enc_smem);
  enc_load_shared<BITS>(sh);

  const uint32_t C = p.geo.channels;
  const uint32_t spb = p.geo.samples_per_block;
  const uint32_t bs = p.geo.block_size;
  const uint32_t trials = p.trials;            /* 1 or 2 here */
  const uint32_t per = 2u * trials;            /* dry slots per block = chain lanes one helper can serve */
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = threadIdx.x >> 5;
  const uint64_t total = (uint64_t)p.num_streams * C;
  const uint64_t group = SPEC ? blockIdx.x : (uint64_t)blockIdx.x * kRolesWarps + warp;   /* ka chains each */
  const uint32_t kh = (ka + per - 1u) / per;

  /* SPEC = 1 shared state: mail, then (1 + trials) scratch blocks per stream of the CTA */
  constexpr uint32_t kRingWarps = SPEC ? 2u : (uint32_t)kRolesWarps;
  unsigned char *extra = enc_smem + kEncLutBytes + kRingWarps * EncRing<MS>::kWarpBytes;
  RolesMail &mail = *reinterpret_cast<RolesMail *>(extra);
  unsigned char *scratch = extra + ((sizeof(RolesMail) + 15u) & ~15u);
  const uint32_t scratch_pitch = (bs + 15u) & ~15u;

  const bool emitter = SPEC && warp == 1;          /* the warp of speculative emitting passes */
  const bool isA = !emitter && lane < ka && group * ka + lane < total;
  const bool isH = !emitter && lane >= ka && lane < ka + kh;
  const bool isE = emitter && lane < ka && group * ka + lane < total;

  EncSource<MS> src;
  RolesChain me;
  me.stream = 0; me.ch = 0; me.ns = 0;
  src.a = src.b = (const int16_t *)p.pcm;
  src.mode = 0;
  if (isA || isE) me = roles_locate<MS>(p, group * ka + lane, src);
  const uint64_t st = (group * ka + lane) * AADK_STATE_WORDS;

  uint8_t *out = p.aad + me.stream * p.aad_stride;
  if (isA && me.ch == 0 && p.block_begin == 0 && p.byte_base == 0) {
    if (me.ns > 0) aadf_write_file_header(out, C, me.ns, p.sampling_rate, BITS, bs, spb, p.geo.ms);
    if (p.out_sizes) p.out_sizes[me.stream] = me.ns ? (uint32_t)aadf_stream_bytes(me.ns, C, BITS, bs, spb) : 0u;
  }

  EncState S;   /* the carried state (chain lanes; emitter lanes keep a copy) */
  S.w0 = S.w1 = S.w2 = S.w3 = S.idx8 = 0;
  if ((isA || isE) && p.state_in) {
    S.w0 = p.state_in[st + 0]; S.w1 = p.state_in[st + 1]; S.w2 = p.state_in[st + 2]; S.w3 = p.state_in[st + 3];
    S.idx8 = kEncIdxScale * p.state_in[st + 4];
  }

  EncJob<MS> job;
  job.ring.base = (uint32_t)__cvta_generic_to_shared(enc_smem) + kEncLutBytes + warp * EncRing<MS>::kWarpBytes + lane * 16u;

  const uint32_t my_blocks = (isA || isE) ? min(aadf_num_blocks(me.ns, spb), p.block_end) : 0u;
  /* all lanes walk the launch's block range together; a chain that has ended (ragged batches) idles */
  const uint32_t last = min(aadf_num_blocks(p.uniform_samples, spb), p.block_end);
  /* the stream slot of this chain inside the CTA (SPEC = 1 scratch): streams are whole in a CTA when ka % C == 0 */
  const uint32_t local_stream = (uint32_t)((group * ka + lane) / C - (group * ka) / C);

  for (uint32_t b = p.block_begin; b < last; b++) {
    const bool live = b < my_blocks;                 /* this chain has a block b */
    const uint32_t n = live ? min(spb, me.ns - b * spb) : 0u;
    const uint32_t first = b * spb - (uint32_t)p.sample_base;
    const bool opening = (b == 0);                   /* first block of the stream: no previous-block passes */
    uint8_t *blk = out + (AADF_FILE_HEADER_BYTES + (uint64_t)b * bs - p.byte_base);

    EncState run = S, cand[kRolesMaxTrials] = {S, S};
    double rmse[kRolesMaxTrials] = {0.0, 0.0};
    double helped[2 * kRolesMaxTrials] = {0.0, 0.0, 0.0, 0.0};   /* helper lanes: baseline error of chain lane per * j + s */
    double rmse0 = 0.0;

    /* ---- dry slots ---------------------------------------------------------------------------------------
     * regular block: slot 2k = prev_k, slot 2k + 1 = cur_k; opening block: slot k = cur_k (and the baseline is
     * cur_1 itself: same start state, same samples).  Helper lanes: the baseline pass of chain lane per * j + s.
     * SPEC = 1, second warp: the emitting pass from the carried state (slot 0) / from the state before cur_k. */
    const uint32_t slots = opening ? trials : per;
    for (uint32_t s = 0; s < slots; s++) {
      const bool cur_slot = opening || (s & 1u);
      const uint32_t k = opening ? s : (s >> 1);     /* trial index of this slot */
      /* helper: fetch the carried state of the chain lane served in this slot */
      const uint32_t served = (lane - ka) * per + s;                      /* meaningful for helper lanes */
      const bool helping = isH && !opening && served < ka;
      const EncState theirs = roles_shfl_state(S, helping ? served : lane);
      job.run = false;
      job.emit = false;
      job.blk = blk;
      EncSource<MS> jsrc = src;
      uint32_t jch = me.ch;
      if (isA && live) {
        if (cur_slot) cand[k] = run;
        job.c.set(run);
        job.first = cur_slot ? first : first - spb;
        job.n = cur_slot ? n : spb;
        job.run = true;
      } else if (helping) {
        const uint64_t chain = group * ka + served;
        if (chain < total) {
          const RolesChain them = roles_locate<MS>(p, chain, jsrc);
          jch = them.ch;
          if (b < min(aadf_num_blocks(them.ns, spb), p.block_end)) {
            job.c.set(theirs);
            job.first = first;
            job.n = min(spb, them.ns - b * spb);
            job.run = true;
          }
        }
      } else if (SPEC && isE && live) {
        /* emitting pass from the carried state (slot 0, candidate 0) or from the state before cur_k (in the slot
         * of cur_k, candidate k + 1); on an opening block slot 0's candidate is the carried state itself */
        if (s == 0u || cur_slot) {
          const uint32_t which = (s == 0u) ? 0u : k + 1u;
          job.c.set(s == 0u ? S : mail.cand[s & 1u][lane]);
          job.first = first;
          job.n = n;
          job.run = true;
          job.emit = true;
          /* scratch block `which` of this chain's stream, laid out like the real block */
          job.blk = scratch + ((size_t)local_stream * (1u + kRolesMaxTrials) + which) * scratch_pitch;
        }
      }
      enc_run_job<BITS, MS>(job, jsrc, jch, C, sh);
      if (isA && live) {
        if (job.run) run = job.c.state();
        if (cur_slot) rmse[k] = enc_rmse(job.sum, job.n);
      } else if (helping) {
        const double e = enc_rmse(job.sum, job.n);
#pragma unroll
        for (int q = 0; q < 2 * kRolesMaxTrials; q++)
          if ((uint32_t)q == s) helped[q] = e;
      } else if (SPEC && isE && live && job.run) {
        mail.after[(s == 0u) ? 0u : k + 1u][lane] = job.c.state();
      }
      if (SPEC) {
        /* hand the next candidate to the second warp: the state the chain is in before its next cur pass */
        const bool next_is_cur = opening ? (s + 1u < slots) : ((s & 1u) == 0u);
        if (isA && live && next_is_cur) mail.cand[(s + 1u) & 1u][lane] = run;
        __syncthreads();
      }
    }

    /* ---- decide (src/aad_encoder.c:518-557): baseline first, then every trial in order, strictly better wins ---- */
    if (!opening) {
#pragma unroll
      for (int q = 0; q < 2 * kRolesMaxTrials; q++) {
        const double v = __shfl_sync(0xFFFFFFFFu, helped[q], ka + lane / per);
        if ((uint32_t)q == lane % per) rmse0 = v;
      }
    } else {
      rmse0 = rmse[0];
    }
    EncState best = S;
    int32_t winner = 0;
    if (isA && live) {
      double best_rmse = rmse0;
      for (uint32_t k = 0; k < trials; k++) {
        if (best_rmse > rmse[k]) {   /* NaN compares false, like the reference */
          best_rmse = rmse[k];
          best = cand[k];
          winner = (int32_t)k + 1;
        }
      }
      /* an opening block's first candidate IS the carried state (never strictly better than itself) */
    }

    if (!SPEC) {
      /* ---- the emitting slot ---- */
      job.run = isA && live;
      job.emit = true;
      job.blk = blk;
      job.c.set(best);
      job.first = first;
      job.n = n;
      enc_run_job<BITS, MS>(job, src, me.ch, C, sh);
      if (isA && live) S = job.c.state();
    } else {
      /* ---- commit: the winner's scratch bytes go to the stream, its end state becomes the carried state ---- */
      if (isA && live) mail.winner[lane] = winner;
      __syncthreads();
      if ((isA || isE) && live) S = mail.after[mail.winner[lane]][lane];
      if (emitter) {
        /* the second warp copies, stream by stream, every byte of the block from its channel's winning scratch block */
        const uint32_t streams_here = (uint32_t)min((uint64_t)((ka + C - 1u) / C), (uint64_t)p.num_streams - (group * ka) / C);
        for (uint32_t ls = 0; ls < streams_here; ls++) {
          const uint64_t stream = (group * ka) / C + ls;
          const uint32_t sns = p.num_samples ? p.num_samples[stream] : p.uniform_samples;
          if (b >= min(aadf_num_blocks(sns, spb), p.block_end)) continue;
          const uint32_t nn = min(spb, sns - b * spb);
          const uint32_t bytes = aadf_block_bytes(nn, C, BITS);
          uint8_t *dst = p.aad + stream * p.aad_stride + (AADF_FILE_HEADER_BYTES + (uint64_t)b * bs - p.byte_base);
          constexpr uint32_t GB = (BITS == 3) ? 3u : 1u;
          for (uint32_t i = lane; i < bytes; i += 32u) {
            const uint32_t chn = (i < AADF_CHANNEL_HEADER_BYTES * C) ? i / AADF_CHANNEL_HEADER_BYTES
                                                                     : ((i - AADF_CHANNEL_HEADER_BYTES * C) / GB) % C;
            const int32_t w = mail.winner[ls * C + chn];
            dst[i] = scratch[((size_t)ls * (1u + kRolesMaxTrials) + (uint32_t)w) * scratch_pitch + i];
          }
        }
      }
      __syncthreads();
    }
  }

  if (isA && p.state_out) {
    p.state_out[st + 0] = S.w0;
    p.state_out[st + 1] = S.w1;
    p.state_out[st + 2] = S.w2;
    p.state_out[st + 3] = S.w3;
    p.state_out[st + 4] = S.idx8 / kEncIdxScale;
  }
}

/* chain lanes per warp so that one wave of single warps per scheduler holds every chain; 0 = does not fit */
inline uint32_t roles_chain_lanes(uint64_t chains, uint32_t trials, int sms)
{
  const uint64_t schedulers = (uint64_t)sms * 4u;
  const uint32_t per = 2u * trials;
  uint32_t ka = (uint32_t)((chains + schedulers - 1) / schedulers);
  if (ka == 0) ka = 1;
  return (ka + (ka + per - 1) / per <= 32u) ? ka : 0u;
}

inline bool roles_eligible(const aadk_encode_params &p)
{
  return enc_fast_eligible(p) && p.segment_blocks == 0 && p.trials >= 1 && p.trials <= (uint32_t)kRolesMaxTrials;
}

template <int BITS, int MS>
int roles_launch(const aadk_encode_params &p, uint32_t ka, cudaStream_t s)
{
  const uint64_t chains = (uint64_t)p.num_streams * p.geo.channels;
  const uint64_t warps = (chains + ka - 1) / ka;
  const size_t smem = kEncLutBytes + (size_t)kRolesWarps * EncRing<MS>::kWarpBytes;
  int dev = 0, sms = 0;
  if (int rc = device_sm_count(&dev, &sms)) return rc;
  if (int rc = allow_dynamic_smem(aad_encode_roles<BITS, MS, 0>, dev, smem)) return rc;
  aad_encode_roles<BITS, MS, 0><<<(unsigned)((warps + kRolesWarps - 1) / kRolesWarps), kRolesWarps * 32, smem, s>>>(p, ka);
  return (int)cudaGetLastError();
}

/* SPEC = 1: chains per CTA = whole streams, at most kRolesSpecStreams of them and kRolesSpecChains chains */
inline uint32_t roles_spec_chain_lanes(const aadk_encode_params &p, int sms)
{
  const uint32_t C = p.geo.channels;
  if (C > (uint32_t)kRolesSpecChains) return 0u;
  uint32_t streams = (uint32_t)kRolesSpecChains / C;
  if (streams > (uint32_t)kRolesSpecStreams) streams = kRolesSpecStreams;
  /* spread the streams over the SMs first: one CTA per SM keeps each warp on a scheduler of its own */
  const uint32_t want = (uint32_t)(((uint64_t)p.num_streams + sms - 1) / sms);
  if (want > streams) return (p.num_streams <= 2u * (uint64_t)sms * streams) ? streams * C : 0u;
  return (want ? want : 1u) * C;
}

template <int BITS, int MS>
int roles_spec_launch(const aadk_encode_params &p, uint32_t ka, cudaStream_t s)
{
  const uint32_t C = p.geo.channels;
  const uint64_t chains = (uint64_t)p.num_streams * C;
  const uint64_t ctas = (chains + ka - 1) / ka;
  const size_t scratch = (size_t)(ka / C) * (1u + kRolesMaxTrials) * ((p.geo.block_size + 15u) & ~15u);
  const size_t smem = kEncLutBytes + 2u * EncRing<MS>::kWarpBytes + ((sizeof(RolesMail) + 15u) & ~15u) + scratch;
  int dev = 0, sms = 0;
  if (int rc = device_sm_count(&dev, &sms)) return rc;
  if (int rc = allow_dynamic_smem(aad_encode_roles<BITS, MS, 1>, dev, smem)) return rc;
  aad_encode_roles<BITS, MS, 1><<<(unsigned)ctas, 64, smem, s>>>(p, ka);
  return (int)cudaGetLastError();
}

/* Which schedule a launch gets (g_enc_schedule, AADGpu_SetEncoderSchedule): 1 = by shape (default), 0 = one pass at a
 * time in every thread, 2 = pass pairing (aad_encode_fast PAIR), 3 = helper lanes, 4 = helper lanes + emitting warp.
 * A forced schedule falls back to `by shape` rules where the launch does not qualify for it.  All bit-exact. */
template <int BITS>
int enc_fast_launch(const aadk_encode_params &p, cudaStream_t s)
{
  const uint64_t chains = (uint64_t)p.num_streams * p.geo.channels * (p.segment_blocks ? p.num_segments : 1u);
  const bool ms = p.geo.ms && p.geo.channels >= 2;
  const int mode = g_enc_schedule;
  int dev = 0, sms = 148;
  if (int rc = device_sm_count(&dev, &sms)) return rc;
  if (mode != 0 && mode != 2 && roles_eligible(p)) {
    const uint32_t ka_spec = roles_spec_chain_lanes(p, sms);
    if (ka_spec != 0 && mode != 3)
      return ms ? roles_spec_launch<BITS, 1>(p, ka_spec, s) : roles_spec_launch<BITS, 0>(p, ka_spec, s);
    const uint32_t ka = roles_chain_lanes(chains, p.trials, sms);
    if (ka != 0 && mode != 4) return ms ? roles_launch<BITS, 1>(p, ka, s) : roles_launch<BITS, 0>(p, ka, s);
  }
  /* up to one warp per scheduler: pair the two independent dry passes of a block inside each thread */
  const bool pair = p.trials >= 1 && chains <= (uint64_t)sms * 4 * 32 && mode != 0;
  if (pair) return ms ? enc_fast_launch_as<BITS, 1, 1>(p, s) : enc_fast_launch_as<BITS, 0, 1>(p, s);
  return ms ? enc_fast_launch_as<BITS, 1, 0>(p, s) : enc_fast_launch_as<BITS, 0, 0>(p, s);
}

}  // namespace
