/*
 * aad_wav.c -- see aad_wav.h.  Plain C, no I/O: callers hand in whole-file images.
 */
#include "aad_wav.h"

#include <string.h>

static uint32_t le16(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
static uint32_t le32(const uint8_t *p) { return le16(p) | (le16(p + 2) << 16); }
static void put_le16(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
static void put_le32(uint8_t *p, uint32_t v) { put_le16(p, v); put_le16(p + 2, v >> 16); }

enum aadwav_result aadwav_parse(const uint8_t *image, size_t size, struct aadwav_info *info)
{
  if (image == NULL || info == NULL) return AADWAV_INVALID_ARGUMENT;
  /* "RIFF" <size> "WAVE" "fmt " <fmt size> <16 bytes of format>: src/wav.c:120-166 */
  if (size < 12) return AADWAV_TRUNCATED;
  if (memcmp(image, "RIFF", 4) != 0 || memcmp(image + 8, "WAVE", 4) != 0) return AADWAV_INVALID_FORMAT;
  if (size < 16) return AADWAV_TRUNCATED;
  if (memcmp(image + 12, "fmt ", 4) != 0) return AADWAV_INVALID_FORMAT;
  if (size < 36) return AADWAV_TRUNCATED;
  const uint32_t fmt_size = le32(image + 16);
  if (le16(image + 20) != 1u) return AADWAV_INVALID_FORMAT;          /* linear PCM only */
  info->num_channels = le16(image + 22);
  info->sampling_rate = le32(image + 24);
  info->bits_per_sample = le16(image + 34);
  size_t pos = 36;
  if (fmt_size > 16u) pos += (size_t)fmt_size - 16u;                 /* extension: skipped, src/wav.c:168-171 */
  /* chunks up to "data" are skipped by their declared size, without padding: src/wav.c:176-193 */
  for (;;) {
    if (pos + 4 > size) return AADWAV_TRUNCATED;
    if (memcmp(image + pos, "data", 4) == 0) break;
    if (pos + 8 > size) return AADWAV_TRUNCATED;
    pos += 8 + (size_t)le32(image + pos + 4);
  }
  if (pos + 8 > size) return AADWAV_TRUNCATED;
  const uint32_t data_bytes = le32(image + pos + 4);
  pos += 8;
  const uint32_t b = info->bits_per_sample;
  if (b != 8u && b != 16u && b != 24u && b != 32u) return AADWAV_INVALID_FORMAT;   /* src/wav.c:222-238 */
  if (info->num_channels == 0u) return AADWAV_INVALID_FORMAT;
  info->num_samples = data_bytes / ((b / 8u) * info->num_channels);   /* src/wav.c:196-199 */
  info->data_offset = pos;
  const size_t need = (size_t)info->num_samples * info->num_channels * (b / 8u);
  if (pos + need > size) return AADWAV_TRUNCATED;                     /* the reference fails reading it */
  return AADWAV_OK;
}

int32_t aadwav_sample32(const uint8_t *data, uint32_t bits_per_sample, size_t i)
{
  switch (bits_per_sample) {   /* src/wav.c:391-415 */
    case 8:  return (int32_t)(((uint32_t)data[i] - 128u) << 24);
    case 16: return (int32_t)(le16(data + 2 * i) << 16);
    case 24: return (int32_t)(((uint32_t)data[3 * i] | ((uint32_t)data[3 * i + 1] << 8) | ((uint32_t)data[3 * i + 2] << 16)) << 8);
    default: return (int32_t)le32(data + 4 * i);
  }
}

void aadwav_to_pcm16(const uint8_t *data, uint32_t bits_per_sample, size_t count, int16_t *out)
{
  size_t i;
  switch (bits_per_sample) {
    case 8:
      for (i = 0; i < count; i++) out[i] = (int16_t)(((int32_t)data[i] - 128) * 256);
      break;
    case 16:
      for (i = 0; i < count; i++) out[i] = (int16_t)le16(data + 2 * i);
      break;
    case 24:
      for (i = 0; i < count; i++) out[i] = (int16_t)le16(data + 3 * i + 1);
      break;
    default:
      for (i = 0; i < count; i++) out[i] = (int16_t)le16(data + 4 * i + 2);
      break;
  }
}

size_t aadwav_write_header(uint8_t *image, uint32_t num_channels, uint32_t sampling_rate, uint32_t bits_per_sample,
                           uint32_t num_samples)
{
  /* src/wav.c:562-627 */
  const uint32_t bytes_per_sample = bits_per_sample / 8u;
  const uint32_t data_bytes = num_samples * bytes_per_sample * num_channels;
  memcpy(image, "RIFF", 4);
  put_le32(image + 4, data_bytes + AADWAV_HEADER_BYTES - 8u);
  memcpy(image + 8, "WAVE", 4);
  memcpy(image + 12, "fmt ", 4);
  put_le32(image + 16, 16u);
  put_le16(image + 20, 1u);
  put_le16(image + 22, num_channels);
  put_le32(image + 24, sampling_rate);
  put_le32(image + 28, sampling_rate * bytes_per_sample * num_channels);
  put_le16(image + 32, bytes_per_sample * num_channels);
  put_le16(image + 34, bits_per_sample);
  memcpy(image + 36, "data", 4);
  put_le32(image + 40, data_bytes);
  return AADWAV_HEADER_BYTES;
}

void aadwav_store32(uint8_t *data, uint32_t bits_per_sample, size_t i, int32_t pcm32)
{
  switch (bits_per_sample) {   /* src/wav.c:418-436: arithmetic shifts, low bytes written */
    case 8:
      data[i] = (uint8_t)((pcm32 >> 24) + 128);
      break;
    case 16:
      put_le16(data + 2 * i, (uint32_t)(pcm32 >> 16));
      break;
    case 24: {
      const uint32_t v = (uint32_t)(pcm32 >> 8);
      data[3 * i] = (uint8_t)v; data[3 * i + 1] = (uint8_t)(v >> 8); data[3 * i + 2] = (uint8_t)(v >> 16);
      break;
    }
    default:
      put_le32(data + 4 * i, (uint32_t)pcm32);
      break;
  }
}
