/*
 * aad_encoder.h -- encoder half of the drop-in API (replaces src/aad_encoder.h:1-56).
 * Host code is C; the block encoder runs as sm_100a CUDA kernels behind these calls.
 */
#ifndef AAD_ENCODER_H_INCLDED
#define AAD_ENCODER_H_INCLDED

#include "aad.h"
#include <stdint.h>

/* src/aad_encoder.h:8-15 */
struct AADEncodeParameter {
  uint16_t num_channels;
  uint32_t sampling_rate;
  uint16_t bits_per_sample;
  uint16_t max_block_size;                    /* bytes */
  AADChannelProcessMethod ch_process_method;
  uint8_t  num_encode_trials;                 /* 0 = no start-state search */
};

struct AADEncoder;

#ifdef __cplusplus
extern "C" {
#endif

/* src/aad_encoder.h:25-27 / src/aad_encoder.c:85-131.  num_samples_per_block may be NULL. */
AADApiResult AADEncoder_CalculateBlockSize(
    uint16_t max_block_size, uint16_t num_channels, uint32_t bits_per_sample,
    uint16_t *block_size, uint32_t *num_samples_per_block);

/* src/aad_encoder.h:30-31 / src/aad_encoder.c:134-221.  Validates before writing. */
AADApiResult AADEncoder_EncodeHeader(
    const struct AADHeaderInfo *header_info, uint8_t *data, uint32_t data_size);

/* src/aad_encoder.h:34 / src/aad_encoder.c:224-245.  -1 when max_block_size cannot hold a block. */
int32_t AADEncoder_CalculateWorkSize(uint16_t max_block_size);

/* src/aad_encoder.h:37 / src/aad_encoder.c:248-316.  (work == NULL && work_size == 0) lets the
 * library allocate; otherwise the handle is placed in caller memory.  NULL on failure. */
struct AADEncoder *AADEncoder_Create(uint16_t max_block_size, void *work, int32_t work_size);

/* src/aad_encoder.h:40 / src/aad_encoder.c:319-327 */
void AADEncoder_Destroy(struct AADEncoder *encoder);

/* src/aad_encoder.h:43-44 / src/aad_encoder.c:779-811 */
AADApiResult AADEncoder_SetEncodeParameter(
    struct AADEncoder *encoder, const struct AADEncodeParameter *parameter);

/* src/aad_encoder.h:47-50 / src/aad_encoder.c:814-891.  input[ch][smpl], values in int16 range. */
AADApiResult AADEncoder_EncodeWhole(
    struct AADEncoder *encoder,
    const int32_t *const *input, uint32_t num_samples,
    uint8_t *data, uint32_t data_size, uint32_t *output_size);

#ifdef __cplusplus
}
#endif

#endif /* AAD_ENCODER_H_INCLDED */
