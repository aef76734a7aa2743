/*
 * aad_encoder.c -- drop-in encoder API (include/aad_encoder.h) on top of the CUDA path.
 *
 * What stays on the host, because it is O(1): block-size arithmetic, the 31-byte stream
 * header, handle bookkeeping, parameter validation.  What runs on the B200: everything
 * AADEncoder_EncodeWhole does per block (src/aad_encoder.c:853-886) -- start-state search,
 * block header, sample chain, bit packing.  No CPU fallback: without a CUDA device
 * EncodeWhole returns AAD_APIRESULT_NG.
 */
#include <stdlib.h>
#include <string.h>

#include "aad_encoder.h"
#include "aad_gpu_internal.h"

#define AADENC_ALIGNMENT 16   /* src/aad_internal.h:7 */

/* Per-handle state.  `chain` is what the reference keeps in processor[] across blocks AND
 * across EncodeWhole calls: weights are zeroed only by Create (src/aad_encoder.c:299-301),
 * the step index only by SetEncodeParameter (src/aad_encoder.c:797-799, src/aad_tables.c:114). */
struct AADEncoder {
  struct AADHeaderInfo header;
  uint8_t set_parameter;
  uint8_t alloced_by_own;
  uint8_t num_encode_trials;
  uint16_t max_block_size;
  int32_t chain[AADF_MAX_CHANNELS][AADK_STATE_WORDS];
  void *work;
};

AADApiResult AADEncoder_CalculateBlockSize(uint16_t max_block_size, uint16_t num_channels, uint32_t bits_per_sample,
                                           uint16_t *block_size, uint32_t *num_samples_per_block)
{
  uint32_t bs = 0, spb = 0;
  if (block_size == NULL) return AAD_APIRESULT_INVALID_ARGUMENT;
  if (!aadf_block_geometry(max_block_size, num_channels, bits_per_sample, aadgpu_max_channels(), &bs, &spb))
    return AAD_APIRESULT_INVALID_FORMAT;
  *block_size = (uint16_t)bs;
  if (num_samples_per_block != NULL) *num_samples_per_block = spb;
  return AAD_APIRESULT_OK;
}

AADApiResult AADEncoder_EncodeHeader(const struct AADHeaderInfo *h, uint8_t *data, uint32_t data_size)
{
  if (h == NULL || data == NULL) return AAD_APIRESULT_INVALID_ARGUMENT;
  if (data_size < AAD_HEADER_SIZE) return AAD_APIRESULT_INSUFFICIENT_DATA;
  /* everything is validated before the first byte is written (src/aad_encoder.c:149-185) */
  if (h->num_channels == 0 || h->num_channels > aadgpu_max_channels()) return AAD_APIRESULT_INVALID_FORMAT;
  if (h->num_samples == 0 || h->sampling_rate == 0) return AAD_APIRESULT_INVALID_FORMAT;
  if (h->bits_per_sample > AAD_MAX_BITS_PER_SAMPLE || h->bits_per_sample < AAD_MIN_BITS_PER_SAMPLE)
    return AAD_APIRESULT_INVALID_FORMAT;
  if (h->block_size <= AADF_CHANNEL_HEADER_BYTES * (uint32_t)h->num_channels) return AAD_APIRESULT_INVALID_FORMAT;
  if (h->num_samples_per_block == 0) return AAD_APIRESULT_INVALID_FORMAT;
  if ((uint32_t)h->ch_process_method >= (uint32_t)AAD_CH_PROCESS_METHOD_INVALID) return AAD_APIRESULT_INVALID_FORMAT;
  if (h->ch_process_method == AAD_CH_PROCESS_METHOD_MS && h->num_channels == 1) return AAD_APIRESULT_INVALID_FORMAT;
  aadf_write_file_header(data, h->num_channels, h->num_samples, h->sampling_rate, h->bits_per_sample, h->block_size,
                         h->num_samples_per_block, (uint32_t)h->ch_process_method);
  return AAD_APIRESULT_OK;
}

int32_t AADEncoder_CalculateWorkSize(uint16_t max_block_size)
{
  uint32_t bs, spb;
  /* same acceptance test as the reference: a mono 2-bit block must fit (src/aad_encoder.c:232-236) */
  if (!aadf_block_geometry(max_block_size, 1, AAD_MIN_BITS_PER_SAMPLE, aadgpu_max_channels(), &bs, &spb)) return -1;
  /* no per-block sample buffers are needed on the host: those live in HBM */
  return (int32_t)(AADENC_ALIGNMENT + sizeof(struct AADEncoder));
}

struct AADEncoder *AADEncoder_Create(uint16_t max_block_size, void *work, int32_t work_size)
{
  uint8_t own = 0;
  const int32_t need = AADEncoder_CalculateWorkSize(max_block_size);
  if (need < 0) return NULL;
  if (work == NULL && work_size == 0) {
    work_size = need;
    work = malloc((size_t)work_size);
    own = 1;
  }
  if (work == NULL || work_size < need) return NULL;   /* also: NULL xor zero size */
  uintptr_t at = ((uintptr_t)work + (AADENC_ALIGNMENT - 1)) & ~(uintptr_t)(AADENC_ALIGNMENT - 1);
  struct AADEncoder *enc = (struct AADEncoder *)at;
  memset(enc, 0, sizeof(*enc));
  enc->work = work;
  enc->alloced_by_own = own;
  enc->max_block_size = max_block_size;
  return enc;
}

void AADEncoder_Destroy(struct AADEncoder *encoder)
{
  if (encoder != NULL && encoder->alloced_by_own == 1) free(encoder->work);
}

AADApiResult AADEncoder_SetEncodeParameter(struct AADEncoder *encoder, const struct AADEncodeParameter *prm)
{
  uint32_t bs = 0, spb = 0;
  if (encoder == NULL || prm == NULL) return AAD_APIRESULT_INVALID_ARGUMENT;
  /* src/aad_encoder.c:741-770 */
  if (prm->bits_per_sample == 0 || prm->bits_per_sample > AAD_MAX_BITS_PER_SAMPLE) return AAD_APIRESULT_INVALID_FORMAT;
  if (prm->max_block_size < AADF_CHANNEL_HEADER_BYTES * (uint32_t)prm->num_channels) return AAD_APIRESULT_INVALID_FORMAT;
  if ((uint32_t)prm->ch_process_method >= (uint32_t)AAD_CH_PROCESS_METHOD_INVALID) return AAD_APIRESULT_INVALID_FORMAT;
  if (!aadf_block_geometry(prm->max_block_size, prm->num_channels, prm->bits_per_sample, aadgpu_max_channels(), &bs, &spb))
    return AAD_APIRESULT_INVALID_FORMAT;
  memset(&encoder->header, 0, sizeof(encoder->header));
  encoder->header.num_channels = prm->num_channels;
  encoder->header.sampling_rate = prm->sampling_rate;
  encoder->header.bits_per_sample = prm->bits_per_sample;
  encoder->header.ch_process_method = prm->ch_process_method;
  encoder->header.block_size = (uint16_t)bs;
  encoder->header.num_samples_per_block = spb;
  encoder->num_encode_trials = prm->num_encode_trials;
  /* table re-initialisation resets the step index of every channel and nothing else */
  for (int c = 0; c < AADF_MAX_CHANNELS; c++) encoder->chain[c][4] = 0;
  encoder->set_parameter = 1;
  return AAD_APIRESULT_OK;
}

AADApiResult AADEncoder_EncodeWhole(struct AADEncoder *encoder, const int32_t *const *input, uint32_t num_samples,
                                    uint8_t *data, uint32_t data_size, uint32_t *output_size)
{
  if (encoder == NULL || input == NULL || data == NULL || output_size == NULL) return AAD_APIRESULT_INVALID_ARGUMENT;
  if (encoder->set_parameter == 0) return AAD_APIRESULT_PARAMETER_NOT_SET;
  encoder->header.num_samples = num_samples;
  const AADApiResult hr = AADEncoder_EncodeHeader(&encoder->header, data, data_size);
  if (hr != AAD_APIRESULT_OK) return hr;

  const struct AADHeaderInfo *h = &encoder->header;
  struct aadf_geometry geo;
  geo.channels = h->num_channels;
  geo.bits = h->bits_per_sample;
  geo.block_size = h->block_size;
  geo.samples_per_block = h->num_samples_per_block;
  geo.ms = (h->ch_process_method == AAD_CH_PROCESS_METHOD_MS) ? 1u : 0u;
  for (uint32_t c = 0; c < geo.channels; c++)
    if (input[c] == NULL) return AAD_APIRESULT_INVALID_ARGUMENT;
  /* The reference only asserts that the caller's buffer is big enough (src/aad_encoder.c:884-885);
   * writing past it is not an option here, so an undersized buffer is reported instead. */
  if (aadf_stream_bytes(num_samples, geo.channels, geo.bits, geo.block_size, geo.samples_per_block) > data_size)
    return AAD_APIRESULT_INSUFFICIENT_BUFFER;

  struct AADGpu *gpu = aadgpu_default();
  if (gpu == NULL) return AAD_APIRESULT_NG;
  return aadgpu_encode_stream_i32(gpu, &geo, h->sampling_rate, encoder->num_encode_trials, input, num_samples,
                                  &encoder->chain[0][0], data, output_size);
}
