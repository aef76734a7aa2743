/*
 * aad_encoder.h -- encoder half of the drop-in C ABI of libaad_b200.so.
 *
 * Same seven entry points, parameter struct, argument meaning and result codes as the reference's
 * src/aad_encoder.h:8-50.  Geometry, header writing and handle bookkeeping are host C
 * (aad_b200/csrc/aad_encoder.c, aad_format.h); the start-state search and the block encoder
 * (src/aad_encoder.c:343-727) run as sm_100a CUDA kernels, one GPU thread per (stream, channel) chain.  Output is
 * byte-identical to the reference's.  There is no CPU encode path: without a CUDA device EncodeWhole returns
 * AAD_APIRESULT_NG.
 *
 * Threading: one thread per handle at a time; handles are independent of each other.
 */
#ifndef AAD_B200_ENCODER_H
#define AAD_B200_ENCODER_H

#include <stdint.h>

#include "aad.h"

#ifdef __cplusplus
extern "C" {
#endif

/* layout of src/aad_encoder.h:8-15 */
struct AADEncodeParameter {
  uint16_t num_channels;
  uint32_t sampling_rate;
  uint16_t bits_per_sample;                   /* 2, 3 or 4 */
  uint16_t max_block_size;                    /* bytes; the block size used is the largest that fits (CalculateBlockSize) */
  AADChannelProcessMethod ch_process_method;  /* MS needs at least 2 channels */
  uint8_t  num_encode_trials;                 /* start-state search depth per block; 0 = none, the reference CLI uses 2 */
};

struct AADEncoder;   /* opaque */

/* Block geometry for a parameter set: *block_size <= max_block_size bytes, *num_samples_per_block per channel
 * including the 4 samples carried in the block header; num_samples_per_block may be NULL.  INVALID_FORMAT when no
 * block fits or the bit depth / channel count is out of range (replaces src/aad_encoder.c:85-131). */
AADApiResult AADEncoder_CalculateBlockSize(uint16_t max_block_size, uint16_t num_channels, uint32_t bits_per_sample,
                                           uint16_t *block_size, uint32_t *num_samples_per_block);

/* Writes the 31-byte stream header; every field is validated before the first byte is written.  The format and
 * codec versions written are this build's, whatever the struct holds (replaces src/aad_encoder.c:134-221). */
AADApiResult AADEncoder_EncodeHeader(const struct AADHeaderInfo *header_info, uint8_t *data, uint32_t data_size);

/* Bytes needed for a caller-provided handle area; -1 when max_block_size cannot hold a block
 * (replaces src/aad_encoder.c:224-245). */
int32_t AADEncoder_CalculateWorkSize(uint16_t max_block_size);

/* work == NULL and work_size == 0: the library allocates and Destroy frees.  Otherwise the handle lives, 16-byte
 * aligned, in the caller's `work` (at least CalculateWorkSize(max_block_size) bytes).  NULL on failure
 * (replaces src/aad_encoder.c:248-316, :319-327). */
struct AADEncoder *AADEncoder_Create(uint16_t max_block_size, void *work, int32_t work_size);
void AADEncoder_Destroy(struct AADEncoder *encoder);

/* Validates and stores the parameters and resets the step-size index of every channel; the predictor weights a
 * handle has adapted so far are kept, as in the reference (replaces src/aad_encoder.c:779-811). */
AADApiResult AADEncoder_SetEncodeParameter(struct AADEncoder *encoder, const struct AADEncodeParameter *parameter);

/* Encodes input[channel][sample] (int32 values in int16 range), header included, into `data`; *output_size bytes
 * written.  PARAMETER_NOT_SET before SetEncodeParameter; INSUFFICIENT_BUFFER when data_size is smaller than the
 * stream (the reference only asserts).  The chain state at the end of the stream stays in the handle
 * (replaces src/aad_encoder.c:814-891). */
AADApiResult AADEncoder_EncodeWhole(struct AADEncoder *encoder, const int32_t *const *input, uint32_t num_samples,
                                    uint8_t *data, uint32_t data_size, uint32_t *output_size);

#ifdef __cplusplus
}
#endif

#endif /* AAD_B200_ENCODER_H */
