/*
 * aad_decoder.h -- decoder half of the drop-in C ABI of libaad_b200.so.
 *
 * Same seven entry points, argument meaning and result codes as the reference's src/aad_decoder.h:15-42, so
 * src/main.c and the reference's tests compile and link against this library unchanged.  The host side
 * (header parse and validation, handle bookkeeping: aad_b200/csrc/aad_decoder.c) is C; every block's header
 * parse, code unpacking and sample chain (src/aad_decoder.c:321-475) run as sm_100a CUDA kernels.  There is no
 * CPU decode path: without a CUDA device the two Decode* calls return AAD_APIRESULT_NG.
 *
 * Threading: one thread per handle at a time; handles are independent of each other.
 */
#ifndef AAD_B200_DECODER_H
#define AAD_B200_DECODER_H

#include <stdint.h>

#include "aad.h"

#ifdef __cplusplus
extern "C" {
#endif

struct AADDecoder;   /* opaque */

/* Reads the 31-byte stream header at `data` into *header_info without judging the values (replaces
 * src/aad_decoder.c:99-170).  INVALID_ARGUMENT for NULL pointers, INSUFFICIENT_DATA when data_size < 31,
 * INVALID_FORMAT when the signature is not "AAD\0". */
AADApiResult AADDecoder_DecodeHeader(const uint8_t *data, uint32_t data_size, struct AADHeaderInfo *header_info);

/* Bytes a caller must provide to AADDecoder_Create when it supplies the handle memory itself
 * (replaces src/aad_decoder.c:35-38). */
int32_t AADDecoder_CalculateWorkSize(void);

/* work == NULL and work_size == 0: the library allocates the handle and Destroy frees it.  Otherwise the handle
 * is placed, 16-byte aligned, inside the caller's `work` (at least CalculateWorkSize() bytes) and Destroy frees
 * nothing.  NULL when only one of the two is given or the area is too small (replaces src/aad_decoder.c:41-85). */
struct AADDecoder *AADDecoder_Create(void *work, int32_t work_size);
void AADDecoder_Destroy(struct AADDecoder *decoder);   /* replaces src/aad_decoder.c:88-96 */

/* Validates the header (format and codec version, channel count, bit depth, block geometry: the checks of
 * src/aad_decoder.c:173-225) and keeps it for DecodeBlock.  INVALID_FORMAT when a field is out of range
 * (replaces src/aad_decoder.c:228-253). */
AADApiResult AADDecoder_SetHeader(struct AADDecoder *decoder, const struct AADHeaderInfo *header);

/* Decodes ONE block (all channels) found at `data` into buffer[channel][sample]; int32 samples in int16 range.
 * *num_decode_samples = min(samples per block, buffer_num_samples).  PARAMETER_NOT_SET before SetHeader,
 * INSUFFICIENT_BUFFER when buffer_num_channels is smaller than the stream's, INSUFFICIENT_DATA when data_size does
 * not cover the block's channel headers (replaces src/aad_decoder.c:321-475). */
AADApiResult AADDecoder_DecodeBlock(struct AADDecoder *decoder, const uint8_t *data, uint32_t data_size, int32_t **buffer,
                                    uint32_t buffer_num_channels, uint32_t buffer_num_samples, uint32_t *num_decode_samples);

/* Header and every block of a whole stream; parses and sets the header itself.  INSUFFICIENT_BUFFER when the buffer
 * has fewer channels or samples than the stream, INSUFFICIENT_DATA when the data ends inside a block's channel
 * headers -- the blocks before it are decoded (replaces src/aad_decoder.c:478-538). */
AADApiResult AADDecoder_DecodeWhole(struct AADDecoder *decoder, const uint8_t *data, uint32_t data_size, int32_t **buffer,
                                    uint32_t buffer_num_channels, uint32_t buffer_num_samples);

#ifdef __cplusplus
}
#endif

#endif /* AAD_B200_DECODER_H */
