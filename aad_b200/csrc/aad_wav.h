/*
 * aad_wav.h -- RIFF/WAVE PCM files as whole byte images (host I/O path of the CLI).
 *
 * Replaces the reference's src/wav.c, which parses and writes files bit by bit through a
 * buffered reader (WAVParser_GetBits src/wav.c:455, WAVWriter_PutBits :737) and costs as much as
 * the codec itself for decode.  Here a file is read with one fread into (pinned) memory and
 * parsed in place; samples are converted in bulk.  Same accepted inputs and same output bytes:
 *   - "RIFF" size "WAVE" "fmt " first, format tag 1 (PCM) only, fmt extension bytes skipped
 *     (src/wav.c:107-171), unknown chunks skipped until "data" (src/wav.c:176-193);
 *   - 8 / 16 / 24 / 32 bits per sample (src/wav.c:222-238), samples widened to 32 bits left
 *     justified (src/wav.c:391-415) and narrowed back by the matching shifts (src/wav.c:418-436);
 *   - writer: the canonical 44-byte header (src/wav.c:562-627), interleaved little-endian data.
 */
#ifndef AAD_WAV_H
#define AAD_WAV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AADWAV_HEADER_BYTES 44

struct aadwav_info {
  uint32_t num_channels;
  uint32_t sampling_rate;
  uint32_t bits_per_sample;   /* 8, 16, 24 or 32 */
  uint32_t num_samples;       /* per channel */
  size_t data_offset;         /* first byte of the sample data inside the image */
};

enum aadwav_result {
  AADWAV_OK = 0,
  AADWAV_INVALID_ARGUMENT,
  AADWAV_INVALID_FORMAT,      /* not RIFF/WAVE, "fmt " not first, not PCM, unsupported bit depth */
  AADWAV_TRUNCATED            /* the image ends before the header / declared data does */
};

/* Parse a whole-file image.  On AADWAV_OK the samples are info->num_samples * num_channels
 * interleaved little-endian values of bits_per_sample bits at image + info->data_offset. */
enum aadwav_result aadwav_parse(const uint8_t *image, size_t size, struct aadwav_info *info);

/* Sample i (interleaved order) widened to 32 bits, left justified: what WAVFile_PCM() holds. */
int32_t aadwav_sample32(const uint8_t *data, uint32_t bits_per_sample, size_t i);

/* The 16 most significant bits of every sample, interleaved order kept: (int16_t)(PCM >> 16) of
 * src/main.c:175-179.  For 16-bit files this is a plain copy of the data chunk. */
void aadwav_to_pcm16(const uint8_t *data, uint32_t bits_per_sample, size_t count, int16_t *out);

/* Write the 44-byte header; returns AADWAV_HEADER_BYTES. */
size_t aadwav_write_header(uint8_t *image, uint32_t num_channels, uint32_t sampling_rate, uint32_t bits_per_sample,
                           uint32_t num_samples);

/* Store a left-justified 32-bit sample as sample i of a data area of the given bit depth. */
void aadwav_store32(uint8_t *data, uint32_t bits_per_sample, size_t i, int32_t pcm32);

#ifdef __cplusplus
}
#endif
#endif
