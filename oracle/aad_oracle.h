/*
 * aad_oracle.h -- CPU restatement of the AAD ADPCM hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is the checker the CUDA path is compared against.  Nothing in the product
 * library (aad_b200/) may include, link or call it: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement byte-for-byte
 * against (1) the golden .aad / decoded .wav fixtures shipped by the reference
 * (test/sin300Hz*.aad, test/sin300Hz*_decoded.wav; pinned by
 * test/test_aad_decoder.c:307-316,336-337 and test/make_test_data.sh:4-7), (2) the
 * sha256 table in tests/golden/golden.json generated from the compiled reference, and
 * (3) -- whenever oracle/_ref/ is present -- the compiled reference itself on seeded
 * random, pathological and ragged inputs, for 1/2 channels (stock build) and up to 8
 * channels (build with src/aad.h:13 patched).
 *
 * All sample arithmetic is 32-bit two's-complement wrapping, which is what the
 * reference compiles to.
 */
#ifndef AAD_ORACLE_H
#define AAD_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AAD_ORACLE_MAX_CHANNELS 8

/* return codes mirror AADApiResult (src/aad.h:25-33) */
enum {
  AAD_ORACLE_OK = 0,
  AAD_ORACLE_INVALID_ARGUMENT = 1,
  AAD_ORACLE_INVALID_FORMAT = 2,
  AAD_ORACLE_INSUFFICIENT_BUFFER = 3,
  AAD_ORACLE_INSUFFICIENT_DATA = 4
};

/* What the 31-byte stream header carries (src/aad.h:43-53). */
struct aad_oracle_info {
  uint32_t format_version;
  uint32_t codec_version;
  uint32_t channels;
  uint32_t num_samples;          /* per channel */
  uint32_t sampling_rate;
  uint32_t bits;                 /* 2..4 */
  uint32_t block_size;           /* bytes */
  uint32_t samples_per_block;    /* per channel, includes the 4 header samples */
  uint32_t ms;                   /* 0 = none, 1 = mid/side on channels 0,1 */
};

/* Encoder chain state carried from block to block (and across calls on one handle):
 * src/aad_encoder.c:10-15,21. */
struct aad_oracle_chain {
  int32_t weight[4];
  int32_t stepsize_index;        /* Q4, 0..4080 */
};

/* src/aad_encoder.c:85-131 */
int aad_oracle_geometry(uint32_t max_block_size, uint32_t channels, uint32_t bits,
                        uint32_t *block_size, uint32_t *samples_per_block);

/* src/aad_encoder.c:134-221 */
int aad_oracle_write_header(const struct aad_oracle_info *info, uint8_t *out, size_t cap);

/* src/aad_decoder.c:99-170 (+ validation :173-225 when validate != 0) */
int aad_oracle_read_header(const uint8_t *data, size_t size, struct aad_oracle_info *info,
                           int validate);

/* Whole-stream encode, src/aad_encoder.c:814-891.  pcm is planar int16: channel c
 * starts at pcm + c*ch_stride.  state (nullable) holds `channels` entries: initial
 * chain state in, final state out (NULL = freshly created handle, all zero).
 * Returns bytes written (>0) or -(error code). */
int64_t aad_oracle_encode(const int16_t *pcm, size_t ch_stride, uint32_t channels,
                          uint32_t num_samples, uint32_t sampling_rate, uint32_t bits,
                          uint32_t max_block_size, uint32_t ms, uint32_t trials,
                          struct aad_oracle_chain *state, uint8_t *out, size_t cap);

/* Whole-stream decode, src/aad_decoder.c:478-538.  pcm is planar int16 with room for
 * buf_samples per channel.  Bytes past `size` read as zero (the reference would read
 * out of bounds there). */
int aad_oracle_decode(const uint8_t *data, size_t size, int16_t *pcm, size_t ch_stride,
                      uint32_t buf_channels, uint32_t buf_samples,
                      struct aad_oracle_info *info_out);

/* Batch helpers used for the CPU baseline: clip i lives at pcm + i*clip_stride (planar,
 * channel stride ch_stride) and aad + i*aad_stride.  out_sizes[i] receives the byte
 * count.  Plain loops, one thread. */
int aad_oracle_encode_batch(const int16_t *pcm, size_t clip_stride, size_t ch_stride,
                            uint32_t num_clips, uint32_t channels, uint32_t num_samples,
                            uint32_t sampling_rate, uint32_t bits, uint32_t max_block_size,
                            uint32_t ms, uint32_t trials, uint8_t *aad, size_t aad_stride,
                            uint32_t *out_sizes);
int aad_oracle_decode_batch(const uint8_t *aad, size_t aad_stride, const uint32_t *sizes,
                            uint32_t num_clips, int16_t *pcm, size_t clip_stride,
                            size_t ch_stride, uint32_t buf_channels, uint32_t buf_samples);

#ifdef __cplusplus
}
#endif
#endif
