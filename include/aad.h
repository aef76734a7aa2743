/*
 * aad.h -- shared public types of the AAD codec API (B200 build).
 *
 * Drop-in for the reference's src/aad.h:1-55: same macro names, enum values and
 * struct layout, so code written against the reference headers compiles and links
 * against libaad_b200.so unchanged.
 */
#ifndef AAD_H_INCLDED
#define AAD_H_INCLDED

#include <stdint.h>

#define AAD_CODEC_VERSION        18   /* src/aad.h:7  */
#define AAD_FORMAT_VERSION       4    /* src/aad.h:10 */

/* The stock reference is built for 2 channels (src/aad.h:13).  This build accepts up to
 * 8 through the same API (BASELINE config 4: 8-channel 3-bit); the wire format is the
 * reference's own -- its struct arrays are simply sized by this macro.  Define
 * AAD_STRICT_REFERENCE_LIMITS to compile callers against the stock limit. */
#ifdef AAD_STRICT_REFERENCE_LIMITS
#define AAD_MAX_NUM_CHANNELS     2
#else
#define AAD_MAX_NUM_CHANNELS     8
#endif

#define AAD_MIN_BITS_PER_SAMPLE  2    /* src/aad.h:16 */
#define AAD_MAX_BITS_PER_SAMPLE  4    /* src/aad.h:19 */
#define AAD_HEADER_SIZE          31   /* src/aad.h:22 */

/* src/aad.h:25-33 */
typedef enum AADApiResultTag {
  AAD_APIRESULT_OK = 0,
  AAD_APIRESULT_INVALID_ARGUMENT,
  AAD_APIRESULT_INVALID_FORMAT,
  AAD_APIRESULT_INSUFFICIENT_BUFFER,
  AAD_APIRESULT_INSUFFICIENT_DATA,
  AAD_APIRESULT_PARAMETER_NOT_SET,
  AAD_APIRESULT_NG
} AADApiResult;

/* src/aad.h:36-40 */
typedef enum AADChannelProcessMethodTag {
  AAD_CH_PROCESS_METHOD_NONE = 0,
  AAD_CH_PROCESS_METHOD_MS,
  AAD_CH_PROCESS_METHOD_INVALID
} AADChannelProcessMethod;

/* src/aad.h:43-53 -- the fields of the 31-byte stream header */
struct AADHeaderInfo {
  uint32_t format_version;
  uint32_t codec_version;
  uint16_t num_channels;
  uint32_t num_samples;            /* per channel */
  uint32_t sampling_rate;
  uint16_t bits_per_sample;
  uint16_t block_size;             /* bytes */
  uint32_t num_samples_per_block;  /* per channel, including the 4 carried in the block header */
  AADChannelProcessMethod ch_process_method;
};

#endif /* AAD_H_INCLDED */
