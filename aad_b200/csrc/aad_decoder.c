/*
 * aad_decoder.c -- drop-in decoder API (include/aad_decoder.h) on top of the CUDA path.
 *
 * Host: 31-byte header parse / validation, handle bookkeeping, the block-loop bookkeeping of
 * AADDecoder_DecodeWhole (which blocks exist, how many samples each one yields).
 * B200: every block's header parse, code unpack, sample chain and MS->LR
 * (src/aad_decoder.c:364-470), one thread per (block, channel).
 */
#include <stdlib.h>
#include <string.h>

#include "aad_decoder.h"
#include "aad_gpu_internal.h"

#define AADDEC_ALIGNMENT 16

struct AADDecoder {
  struct AADHeaderInfo header;
  uint8_t alloced_by_own;
  uint8_t set_header;
  void *work;
};

int32_t AADDecoder_CalculateWorkSize(void) { return (int32_t)(AADDEC_ALIGNMENT + sizeof(struct AADDecoder)); }

struct AADDecoder *AADDecoder_Create(void *work, int32_t work_size)
{
  uint8_t own = 0;
  if (work == NULL && work_size == 0) {
    work_size = AADDecoder_CalculateWorkSize();
    work = malloc((size_t)work_size);
    own = 1;
  }
  if (work == NULL || work_size < AADDecoder_CalculateWorkSize()) return NULL;
  uintptr_t at = ((uintptr_t)work + (AADDEC_ALIGNMENT - 1)) & ~(uintptr_t)(AADDEC_ALIGNMENT - 1);
  struct AADDecoder *dec = (struct AADDecoder *)at;
  memset(dec, 0, sizeof(*dec));
  dec->work = work;
  dec->alloced_by_own = own;
  return dec;
}

void AADDecoder_Destroy(struct AADDecoder *decoder)
{
  if (decoder != NULL && decoder->alloced_by_own == 1) free(decoder->work);
}

AADApiResult AADDecoder_DecodeHeader(const uint8_t *data, uint32_t data_size, struct AADHeaderInfo *out)
{
  if (data == NULL || out == NULL) return AAD_APIRESULT_INVALID_ARGUMENT;
  if (data_size < AAD_HEADER_SIZE) return AAD_APIRESULT_INSUFFICIENT_DATA;
  if (data[0] != 'A' || data[1] != 'A' || data[2] != 'D' || data[3] != '\0') return AAD_APIRESULT_INVALID_FORMAT;
  /* past the signature everything is read without judgement (src/aad_decoder.c:134-162) */
  struct AADHeaderInfo h;
  memset(&h, 0, sizeof(h));
  h.format_version = aadf_get_be32(data + 4);
  h.codec_version = aadf_get_be32(data + 8);
  h.num_channels = (uint16_t)aadf_get_be16(data + 12);
  h.num_samples = aadf_get_be32(data + 14);
  h.sampling_rate = aadf_get_be32(data + 18);
  h.bits_per_sample = (uint16_t)aadf_get_be16(data + 22);
  h.block_size = (uint16_t)aadf_get_be16(data + 24);
  h.num_samples_per_block = aadf_get_be32(data + 26);
  h.ch_process_method = (AADChannelProcessMethod)data[30];
  *out = h;
  return AAD_APIRESULT_OK;
}

/* src/aad_decoder.c:173-225 */
static int header_is_decodable(const struct AADHeaderInfo *h)
{
  if (h->format_version != AAD_FORMAT_VERSION || h->codec_version != AAD_CODEC_VERSION) return 0;
  if (h->num_channels == 0 || h->num_channels > aadgpu_max_channels()) return 0;
  if (h->num_samples == 0 || h->sampling_rate == 0) return 0;
  if (h->bits_per_sample < AAD_MIN_BITS_PER_SAMPLE || h->bits_per_sample > AAD_MAX_BITS_PER_SAMPLE) return 0;
  if (h->block_size <= AADF_CHANNEL_HEADER_BYTES * (uint32_t)h->num_channels) return 0;
  if (h->num_samples_per_block == 0) return 0;
  if ((uint32_t)h->ch_process_method >= (uint32_t)AAD_CH_PROCESS_METHOD_INVALID) return 0;
  if (h->ch_process_method == AAD_CH_PROCESS_METHOD_MS && h->num_channels == 1) return 0;
  return 1;
}

AADApiResult AADDecoder_SetHeader(struct AADDecoder *decoder, const struct AADHeaderInfo *header)
{
  if (decoder == NULL || header == NULL) return AAD_APIRESULT_INVALID_ARGUMENT;
  if (!header_is_decodable(header)) return AAD_APIRESULT_INVALID_FORMAT;
  decoder->header = *header;
  decoder->set_header = 1;
  return AAD_APIRESULT_OK;
}

static void geometry_of(const struct AADHeaderInfo *h, struct aadf_geometry *geo)
{
  geo->channels = h->num_channels;
  geo->bits = h->bits_per_sample;
  geo->block_size = h->block_size;
  geo->samples_per_block = h->num_samples_per_block;
  geo->ms = (h->ch_process_method == AAD_CH_PROCESS_METHOD_MS) ? 1u : 0u;
}

/* The block layout the kernels assume (groups start right after the channel headers and a
 * block holds spb samples) must be consistent with the header fields, or a hostile header
 * could index outside the block.  The reference trusts the header (src/aad_decoder.c:394-455
 * reads as many groups as num_samples_per_block asks for). */
static int geometry_is_consistent(const struct aadf_geometry *g)
{
  const uint32_t gs = aadf_group_samples(g->bits), gb = aadf_group_bytes(g->bits);
  const uint64_t groups = ((uint64_t)g->samples_per_block - AADF_TAPS + gs - 1) / gs;
  if (g->samples_per_block < AADF_TAPS) return 0;
  return (uint64_t)g->channels * (AADF_CHANNEL_HEADER_BYTES + groups * gb) <= g->block_size;
}

/* header checks for paths that do not go through a decoder handle (aad_gpu.c) */
AADApiResult aaddec_check_header(const struct AADHeaderInfo *h)
{
  struct aadf_geometry geo;
  if (h == NULL) return AAD_APIRESULT_INVALID_ARGUMENT;
  if (!header_is_decodable(h)) return AAD_APIRESULT_INVALID_FORMAT;
  geometry_of(h, &geo);
  return geometry_is_consistent(&geo) ? AAD_APIRESULT_OK : AAD_APIRESULT_INVALID_FORMAT;
}

AADApiResult AADDecoder_DecodeBlock(struct AADDecoder *decoder, const uint8_t *data, uint32_t data_size,
                                    int32_t **buffer, uint32_t buffer_num_channels, uint32_t buffer_num_samples,
                                    uint32_t *num_decode_samples)
{
  if (decoder == NULL || data == NULL || buffer == NULL || num_decode_samples == NULL)
    return AAD_APIRESULT_INVALID_ARGUMENT;
  if (decoder->set_header != 1) return AAD_APIRESULT_PARAMETER_NOT_SET;
  const struct AADHeaderInfo *h = &decoder->header;
  if (data_size < AADF_CHANNEL_HEADER_BYTES * (uint32_t)h->num_channels) return AAD_APIRESULT_INSUFFICIENT_DATA;
  const uint32_t n = (h->num_samples_per_block < buffer_num_samples) ? h->num_samples_per_block : buffer_num_samples;
  if (buffer_num_channels < h->num_channels) return AAD_APIRESULT_INSUFFICIENT_BUFFER;
  struct aadf_geometry geo;
  geometry_of(h, &geo);
  if (!geometry_is_consistent(&geo)) return AAD_APIRESULT_INVALID_FORMAT;
  if (n == 0) {
    *num_decode_samples = 0;
    return AAD_APIRESULT_OK;
  }
  struct AADGpu *gpu = aadgpu_default();
  if (gpu == NULL) return AAD_APIRESULT_NG;
  /* Present the block as block 0 of a stream: the kernel addresses blocks at +31. */
  const uint32_t take = data_size < geo.block_size ? data_size : geo.block_size;
  uint8_t *tmp = (uint8_t *)malloc((size_t)AAD_HEADER_SIZE + take);
  if (tmp == NULL) return AAD_APIRESULT_NG;
  memset(tmp, 0, AAD_HEADER_SIZE);
  memcpy(tmp + AAD_HEADER_SIZE, data, take);
  const AADApiResult r = aadgpu_decode_stream_i32(gpu, &geo, tmp, AAD_HEADER_SIZE + take, 1, n, n, buffer);
  free(tmp);
  if (r == AAD_APIRESULT_OK) *num_decode_samples = n;
  return r;
}

AADApiResult AADDecoder_DecodeWhole(struct AADDecoder *decoder, const uint8_t *data, uint32_t data_size,
                                    int32_t **buffer, uint32_t buffer_num_channels, uint32_t buffer_num_samples)
{
  if (decoder == NULL || data == NULL || buffer == NULL) return AAD_APIRESULT_INVALID_ARGUMENT;
  struct AADHeaderInfo tmp;
  AADApiResult r = AADDecoder_DecodeHeader(data, data_size, &tmp);
  if (r != AAD_APIRESULT_OK) return r;
  if ((r = AADDecoder_SetHeader(decoder, &tmp)) != AAD_APIRESULT_OK) return r;
  const struct AADHeaderInfo *h = &decoder->header;
  if (buffer_num_channels < h->num_channels || buffer_num_samples < h->num_samples)
    return AAD_APIRESULT_INSUFFICIENT_BUFFER;
  struct aadf_geometry geo;
  geometry_of(h, &geo);
  if (!geometry_is_consistent(&geo)) return AAD_APIRESULT_INVALID_FORMAT;
  for (uint32_t c = 0; c < geo.channels; c++)
    if (buffer[c] == NULL) return AAD_APIRESULT_INVALID_ARGUMENT;

  /* The reference loop (src/aad_decoder.c:514-534) visits block b while b*spb < num_samples and
   * its first byte exists; a last block too short for its channel headers stops the loop with
   * INSUFFICIENT_DATA after the earlier blocks were decoded. */
  const uint32_t by_samples = aadf_num_blocks(h->num_samples, geo.samples_per_block);
  const uint64_t payload = (uint64_t)data_size - AAD_HEADER_SIZE;
  const uint64_t by_bytes = (payload + geo.block_size - 1) / geo.block_size;
  uint32_t blocks = (uint32_t)(by_bytes < by_samples ? by_bytes : by_samples);
  AADApiResult tail = AAD_APIRESULT_OK;
  if (blocks > 0) {
    const uint64_t last_avail = payload - (uint64_t)(blocks - 1) * geo.block_size;
    if (last_avail < (uint64_t)AADF_CHANNEL_HEADER_BYTES * geo.channels) {
      blocks--;
      tail = AAD_APIRESULT_INSUFFICIENT_DATA;
    }
  }
  struct AADGpu *gpu = aadgpu_default();
  if (gpu == NULL) return AAD_APIRESULT_NG;
  r = aadgpu_decode_stream_i32(gpu, &geo, data, data_size, blocks, h->num_samples, buffer_num_samples, buffer);
  return (r != AAD_APIRESULT_OK) ? r : tail;
}
