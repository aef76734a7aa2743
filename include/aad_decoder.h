/*
 * aad_decoder.h -- decoder half of the drop-in API (replaces src/aad_decoder.h:1-48).
 */
#ifndef AAD_DECODER_H_INCLDED
#define AAD_DECODER_H_INCLDED

#include "aad.h"
#include <stdint.h>

struct AADDecoder;

#ifdef __cplusplus
extern "C" {
#endif

/* src/aad_decoder.h:15-16 / src/aad_decoder.c:99-170.  Parses, does not validate. */
AADApiResult AADDecoder_DecodeHeader(
    const uint8_t *data, uint32_t data_size, struct AADHeaderInfo *header_info);

/* src/aad_decoder.h:19 / src/aad_decoder.c:35-38 */
int32_t AADDecoder_CalculateWorkSize(void);

/* src/aad_decoder.h:22 / src/aad_decoder.c:41-85 */
struct AADDecoder *AADDecoder_Create(void *work, int32_t work_size);

/* src/aad_decoder.h:25 / src/aad_decoder.c:88-96 */
void AADDecoder_Destroy(struct AADDecoder *decoder);

/* src/aad_decoder.h:28-29 / src/aad_decoder.c:228-253.  Validates (src/aad_decoder.c:173-225). */
AADApiResult AADDecoder_SetHeader(
    struct AADDecoder *decoder, const struct AADHeaderInfo *header);

/* src/aad_decoder.h:32-36 / src/aad_decoder.c:321-475.  One block; buffer[ch][smpl]. */
AADApiResult AADDecoder_DecodeBlock(
    struct AADDecoder *decoder,
    const uint8_t *data, uint32_t data_size,
    int32_t **buffer, uint32_t buffer_num_channels, uint32_t buffer_num_samples,
    uint32_t *num_decode_samples);

/* src/aad_decoder.h:39-42 / src/aad_decoder.c:478-538.  Header + every block. */
AADApiResult AADDecoder_DecodeWhole(
    struct AADDecoder *decoder,
    const uint8_t *data, uint32_t data_size,
    int32_t **buffer, uint32_t buffer_num_channels, uint32_t buffer_num_samples);

#ifdef __cplusplus
}
#endif

#endif /* AAD_DECODER_H_INCLDED */
