/*
 * aad_kernels.h -- the thin C ABI between the C host code and the sm_100a kernels.
 * Plain pointers and sizes only; every pointer in the *_params structs is a DEVICE
 * pointer, `stream` is a cudaStream_t passed as void*.  Launchers return a cudaError_t
 * value as int (0 = success) and never synchronise.
 */
#ifndef AAD_KERNELS_H
#define AAD_KERNELS_H

#include <stdint.h>
#include "aad_format.h"

#ifdef __cplusplus
extern "C" {
#endif

/* One chain = one (stream, channel).  Encoder state carried between blocks / calls:
 * weight[4] then stepsize_index (src/aad_encoder.c:10-15). */
#define AADK_STATE_WORDS 5

struct aadk_decode_params {
  const uint8_t *aad;          /* stream i starts (file header included) at aad + i*aad_stride */
  uint64_t aad_stride;
  const uint32_t *sizes;       /* valid bytes per stream, or NULL -> uniform_size */
  uint32_t uniform_size;
  uint32_t num_streams;
  struct aadf_geometry geo;
  uint32_t block_begin;        /* blocks [block_begin, block_end) of every stream are decoded */
  uint32_t block_end;
  uint32_t uniform_samples;    /* samples per channel when read_headers == 0; with read_headers: cap on what a header may claim (0 = none) */
  uint32_t read_headers;       /* 1: take num_samples from each stream's own 31-byte header */
  uint32_t buf_samples;        /* output capacity per channel (reference DecodeWhole semantics); 0 = num_samples */
  void *pcm;                   /* planar int16: sample s of channel c of stream i at i*clip_stride + c*ch_stride + s */
  uint64_t pcm_clip_stride;    /* in samples */
  uint64_t pcm_ch_stride;      /* in samples */
  /* Shards of one stream (a device holds only its own byte / sample range): `aad` points at byte byte_base of
   * every stream and the rows of `pcm` start at sample sample_base.  Sizes, block and sample numbers stay
   * absolute; the kernels subtract the bases when they form addresses, so no pointer ever lies outside its
   * allocation.  read_headers needs byte_base == 0. */
  uint64_t byte_base;
  uint64_t sample_base;
  /* 1: WAV order -- sample s of channel c of stream i at i*clip_stride + (s - sample_base)*channels + c
   * (pcm_ch_stride unused); only where aadk_decode_interleaved_ok() says so */
  uint32_t interleaved;
};

/* 1 when aadk_launch_decode can write WAV-order output itself for this shape (mono: the same thing;
 * 2 / 4 / 8 channels on the staged kernels); otherwise decode planar and run aadk_launch_interleave16 */
int aadk_decode_interleaved_ok(const struct aadk_decode_params *p);

struct aadk_encode_params {
  const void *pcm;             /* planar int16, same addressing as the decoder's output */
  uint64_t pcm_clip_stride;
  uint64_t pcm_ch_stride;
  const uint32_t *num_samples; /* per stream, or NULL -> uniform_samples */
  uint32_t uniform_samples;
  uint32_t num_streams;
  struct aadf_geometry geo;
  uint32_t sampling_rate;
  uint32_t trials;             /* num_encode_trials, src/aad_encoder.h:14 */
  uint8_t *aad;                /* stream i written (file header included) at aad + i*aad_stride */
  uint64_t aad_stride;         /* >= aadf_stream_bytes_bound() */
  uint32_t *out_sizes;         /* bytes written per stream, nullable */
  const int32_t *state_in;     /* [stream][channel][AADK_STATE_WORDS], NULL = all zero */
  int32_t *state_out;          /* same shape, nullable */
  uint32_t block_begin;        /* blocks [block_begin, block_end) of every stream are encoded; the */
  uint32_t block_end;          /* chain state enters through state_in and leaves through state_out */
  /* Segment mode (NOT byte-identical to the reference encoder, which carries its state through the whole
   * stream; off when segment_blocks == 0): every run of segment_blocks blocks is encoded as a stream of
   * its own -- zero weights and step index at its first block, no previous-block trial pass there -- so a
   * stream is num_segments independent chains per channel.  The state arrays are then
   * [stream][segment][channel][AADK_STATE_WORDS]; num_segments must be the same for every launch that
   * shares them. */
  uint32_t segment_blocks;
  uint32_t num_segments;       /* >= 1 when segment_blocks != 0 */
  /* segment mode only, 1: block_begin / block_end count from the first block of EVERY segment (blocks
   * [block_begin, block_end) of each segment are encoded), so a host pipeline can cut a stream of many short
   * chains into slices that keep every chain busy */
  uint32_t segment_relative;
  /* with segment_relative: only segments [segment_begin, segment_end) take part (a device's shard of one stream);
   * segment_end == 0 means all of them */
  uint32_t segment_begin;
  uint32_t segment_end;
  /* shards of one stream: `aad` points at byte byte_base of every stream, the rows of `pcm` start at sample
   * sample_base (see aadk_decode_params) */
  uint64_t byte_base;
  uint64_t sample_base;
};

int aadk_launch_decode(const struct aadk_decode_params *p, void *stream);
int aadk_launch_encode(const struct aadk_encode_params *p, void *stream);

/* Deterministic integer-only synthetic PCM (bench / tests): SURVEY.md 8(d).
 * lut = 1024-entry int16 sine table (device). */
struct aadk_synth_params {
  int16_t *pcm;
  uint64_t pcm_clip_stride;
  uint64_t pcm_ch_stride;
  uint32_t num_streams;
  uint32_t channels;
  uint32_t num_samples;
  uint32_t sampling_rate;
  uint32_t first_stream;       /* global index of stream 0 (so shards generate their own slice) */
  const int16_t *lut;
};
int aadk_launch_synth(const struct aadk_synth_params *p, void *stream);

/* PCM16 interleaved (WAV order) <-> planar, src/main.c:122-126,175-179 done on the device */
int aadk_launch_deinterleave16(const int16_t *interleaved, int16_t *planar, uint64_t ch_stride, uint32_t channels,
                               uint32_t num_samples, void *stream);
int aadk_launch_interleave16(const int16_t *planar, uint64_t ch_stride, int16_t *interleaved, uint32_t channels,
                             uint32_t num_samples, void *stream);
/* de-interleave `rows` runs of `width` frames, run r starting at frame r * row_frames + first of the (whole-stream)
 * interleaved / planar buffers; frames at or past `limit` are skipped */
int aadk_launch_deinterleave16_rows(const int16_t *interleaved, int16_t *planar, uint64_t ch_stride, uint32_t channels,
                                    uint64_t row_frames, uint64_t first, uint32_t width, uint32_t rows, uint64_t limit, void *stream);

/* planar int32 rows (the reference API's sample type, int16-range values) <-> planar int16 rows;
 * pitches in elements; widen converts elements [first, first + n) of every row */
int aadk_launch_narrow32(const int32_t *in, uint64_t in_pitch, int16_t *out, uint64_t out_pitch, uint32_t rows, uint64_t n,
                         void *stream);
int aadk_launch_widen16(const int16_t *in, uint64_t in_pitch, int32_t *out, uint64_t out_pitch, uint32_t rows, uint64_t first,
                        uint64_t n, void *stream);

/* WAV data chunk (8 / 16 / 24 / 32 bits per sample, interleaved) -> planar int16 rows: the (int16_t)(PCM >> 16) of
 * src/main.c:175-179 with the widening of src/wav.c:391-415 */
int aadk_launch_wav_to_planar16(const uint8_t *data, uint32_t bits, int16_t *planar, uint64_t ch_stride, uint32_t channels,
                                uint32_t num_samples, void *stream);
/* src/main.c:372-381 (-r) / :418-428 (-g): overwrite the WAV data chunk `data` (count samples) with the reconstruction
 * `decoded` (interleaved int16) or with input minus reconstruction, in the chunk's own sample format */
int aadk_launch_analysis_image(uint8_t *data, uint32_t bits, const int16_t *decoded, uint64_t count, int gap, void *stream);
/* src/main.c:470-497 (-c): per-block partial sums {sum of squares, sum of magnitudes, maximum} of the reference's
 * per-sample error term; partials holds 3 * AADK_STATS_BLOCKS doubles, to be added up in order on the host */
#define AADK_STATS_BLOCKS 592
int aadk_launch_analysis_stats(const uint8_t *data, uint32_t bits, const int16_t *decoded, uint64_t count, double *partials,
                               void *stream);

/* number of kernels launched by this library since load (bench.py's gpu_launches) */
uint64_t aadk_launch_count(void);
uint64_t aadk_tma_launch_count(void);   /* launches of aad_decode_tma (kernel path 7) */
/* tests only: 1 = always use the generic (any-shape) kernels instead of the fast paths;
 * 2 = decode mono / stereo streams with the any-channel-count staged kernel (aad_decode_wide) too;
 * 4 = mono 4-bit decode flushes its output rows through the TMA unit (cp.async.bulk shared -> global; the A/B of
 *     profiles/r02_decoder_experiments.md: bit-exact, 9 % slower than the register flush, hence not the default) */
void aadk_force_generic(int on);
/* tests / measurement: the fast encoder's pass schedule.  1 (default) = chosen by shape; 0 = one pass at a time in
 * every thread; 2 = the two independent dry passes of a block interleaved in one thread; 3 = helper lanes run the
 * baseline passes; 4 = helper lanes + a second warp that runs the emitting passes ahead of the decision.  A forced
 * schedule applies where the launch qualifies for it.  All of them produce the reference's bytes. */
void aadk_set_encoder_schedule(int mode);

#ifdef __cplusplus
}
#endif
#endif
