/*
 * aad_cli.c -- `aad`: the reference's command line (src/main.c) over libaad_b200.so.
 *
 * Same modes and options as src/main.c:20-58 (-e -d -r -g -c -i, -b -s -t -m, -h -v), same
 * defaults (4 bits, block 1024, 2 trials, no MS), same output files byte for byte.  What changed
 * is where the work happens:
 *   - files are read and written as whole images in pinned memory (aad_wav.c) instead of bit by
 *     bit through src/wav.c;
 *   - the samples of a 16-bit WAV go to the GPU as they lie in the file; de-interleaving, the codec
 *     and re-interleaving run on the device (AADGpu_*Interleaved16);
 *   - -r / -g / -c keep the encoded stream in HBM between encode and decode.
 * Additions: --device N | a,b,c | all, and --batch LIST for -e / -d (one "INPUT OUTPUT" pair per line; files of
 * the same shape are encoded / decoded by ONE kernel launch through the batch API).
 */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "aad.h"
#include "aad_b200.h"
#include "aad_decoder.h"
#include "aad_encoder.h"
#include "aad_wav.h"

enum { MODE_ENCODE, MODE_DECODE, MODE_RECONSTRUCT, MODE_GAP, MODE_CALCULATE, MODE_INFORMATION, NUM_MODES };

struct options {
  int mode_set[NUM_MODES];
  int help, version, ms;
  const char *bits, *block, *trials, *device, *batch, *segment;
  const char *files[2];
  int num_files;
};

struct option_spec {
  char short_name;
  const char *long_name;
  int needs_argument;
  const char *description;
};

static const struct option_spec k_specs[] = {
  { 'e', "encode", 0, "Encode mode (wav file -> .aad file)" },
  { 'd', "decode", 0, "Decode mode (.aad file -> wav file)" },
  { 'r', "reconstruct", 0, "Reconstruction mode (wav file -> (encode -> decode) -> decoded wav file)" },
  { 'g', "gap", 0, "Gap(residual output) mode (wav file -> (encode -> decode) -> residual wav file)" },
  { 'c', "calculate", 0, "Calculate statistics(e.g. RMS error) between original and reconstructed wav" },
  { 'i', "information", 0, "Show information of encoded .aad file" },
  { 'b', "bits-per-sample", 1, "Specify bits per sample(in 2,3,4) (default: 4)" },
  { 's', "max-block-size", 1, "Specify max block size (default: 1024)" },
  { 't', "num-encode-trials", 1, "Specify number of encode Trials (default: 2)" },
  { 'm', "ms-conversion", 0, "Switch to use LR to MS conversion (default: no)" },
  { 'D', "device", 1, "CUDA device(s): index, comma separated list, or \"all\" (default: 0); with several, -d shards one file by block range and --batch shards files" },
  { 'B', "batch", 1, "With -e / -d: file listing one \"INPUT OUTPUT\" pair per line, processed as batches" },
  { 'S', "segment-blocks", 1, "Encode every run of N blocks as an independent chain (parallel encode of one long file; decodes with any AAD decoder but is NOT byte-identical to the reference encoder's output) (default: 0 = off)" },
  { 'h', "help", 0, "Show help message" },
  { 'v', "version", 0, "Show version information" },
};
#define NUM_SPECS (sizeof(k_specs) / sizeof(k_specs[0]))

static void apply_option(struct options *o, char short_name, const char *value)
{
  switch (short_name) {
    case 'e': o->mode_set[MODE_ENCODE] = 1; break;
    case 'd': o->mode_set[MODE_DECODE] = 1; break;
    case 'r': o->mode_set[MODE_RECONSTRUCT] = 1; break;
    case 'g': o->mode_set[MODE_GAP] = 1; break;
    case 'c': o->mode_set[MODE_CALCULATE] = 1; break;
    case 'i': o->mode_set[MODE_INFORMATION] = 1; break;
    case 'b': o->bits = value; break;
    case 's': o->block = value; break;
    case 't': o->trials = value; break;
    case 'm': o->ms = 1; break;
    case 'D': o->device = value; break;
    case 'B': o->batch = value; break;
    case 'S': o->segment = value; break;
    case 'h': o->help = 1; break;
    case 'v': o->version = 1; break;
    default: break;
  }
}

/* returns 0 on success; accepted spellings as in src/command_line_parser.c:173-331:
 * -x, -xyz (an option with an argument must come last and takes the next word), --long,
 * --long WORD, --long=WORD; an argument may not start with '-'; no option twice. */
static int parse_arguments(int argc, char **argv, struct options *o)
{
  int seen[NUM_SPECS];
  memset(seen, 0, sizeof(seen));
  memset(o, 0, sizeof(*o));
  for (int i = 1; i < argc; i++) {
    const char *arg = argv[i];
    if (strncmp(arg, "--", 2) == 0) {
      size_t spec = NUM_SPECS;
      const char *value = NULL;
      for (size_t k = 0; k < NUM_SPECS && spec == NUM_SPECS; k++) {
        const size_t len = strlen(k_specs[k].long_name);
        if (strncmp(arg + 2, k_specs[k].long_name, len) != 0) continue;
        if (arg[2 + len] == '\0') spec = k;
        else if (arg[2 + len] == '=' && k_specs[k].needs_argument) { spec = k; value = arg + 2 + len + 1; }
      }
      if (spec == NUM_SPECS) {
        fprintf(stderr, "%s: Unknown long option - \"%s\" \n", argv[0], arg + 2);
        return 1;
      }
      if (seen[spec]) {
        fprintf(stderr, "%s: Option \"%s\" multiply specified. \n", argv[0], k_specs[spec].long_name);
        return 1;
      }
      if (k_specs[spec].needs_argument && value == NULL) {
        if (i + 1 == argc || argv[i + 1][0] == '-') {
          fprintf(stderr, "%s: Option \"%s\" needs argument. \n", argv[0], k_specs[spec].long_name);
          return 1;
        }
        value = argv[++i];
      }
      seen[spec] = 1;
      apply_option(o, k_specs[spec].short_name, value);
    } else if (arg[0] == '-') {
      for (size_t c = 1; arg[c] != '\0'; c++) {
        size_t spec = NUM_SPECS;
        const char *value = NULL;
        for (size_t k = 0; k < NUM_SPECS; k++)
          if (k_specs[k].short_name == arg[c]) { spec = k; break; }
        if (spec == NUM_SPECS) {
          fprintf(stderr, "%s: Unknown short option - \'%c\' \n", argv[0], arg[c]);
          return 1;
        }
        if (seen[spec]) {
          fprintf(stderr, "%s: Option \'%c\' multiply specified. \n", argv[0], arg[c]);
          return 1;
        }
        if (k_specs[spec].needs_argument) {
          if (arg[c + 1] != '\0') {
            fprintf(stderr, "%s: Option \'%c\' needs argument. Please specify tail of short option sequence.\n", argv[0], arg[c]);
            return 1;
          }
          if (i + 1 == argc || argv[i + 1][0] == '-') {
            fprintf(stderr, "%s: Option \'%c\' needs argument. \n", argv[0], arg[c]);
            return 1;
          }
          value = argv[++i];
        }
        seen[spec] = 1;
        apply_option(o, k_specs[spec].short_name, value);
        if (value != NULL) break;   /* arg was the last word of this sequence */
      }
    } else {
      if (o->num_files >= 2) {
        fprintf(stderr, "%s: Too many strings specified. \n", argv[0]);
        return 1;
      }
      o->files[o->num_files++] = arg;
    }
  }
  return 0;
}

static void print_usage(const char *program)
{
  printf("Usage: %s [options] INPUT_FILE_NAME OUTPUT_FILE_NAME \n", program);
}

static void print_help(const char *program)
{
  print_usage(program);
  printf("options: \n");
  for (size_t k = 0; k < NUM_SPECS; k++) {
    char name[64];
    snprintf(name, sizeof(name), "--%s%s", k_specs[k].long_name, k_specs[k].needs_argument ? " (needs argument)" : "");
    printf("  -%c, %-40s %s \n", k_specs[k].short_name, name, k_specs[k].description);
  }
}

/* ---- whole-file I/O in pinned memory ------------------------------------------------------- */

/* Pinned when a GPU context exists (the buffers feed cudaMemcpyAsync), plain malloc otherwise (-i). */
static int g_have_gpu = 0;
static void *io_alloc(size_t bytes) { return g_have_gpu ? AADGpu_HostAlloc(bytes ? bytes : 1) : malloc(bytes ? bytes : 1); }
static void io_free(void *p)
{
  if (g_have_gpu) AADGpu_HostFree(p);
  else free(p);
}

static uint8_t *read_file(const char *name, size_t *size)
{
  FILE *fp = fopen(name, "rb");
  if (fp == NULL) return NULL;
  if (fseek(fp, 0, SEEK_END) != 0) { fclose(fp); return NULL; }
  const long len = ftell(fp);
  if (len < 0 || fseek(fp, 0, SEEK_SET) != 0) { fclose(fp); return NULL; }
  uint8_t *buf = (uint8_t *)io_alloc((size_t)len);
  if (buf != NULL && fread(buf, 1, (size_t)len, fp) != (size_t)len) { io_free(buf); buf = NULL; }
  fclose(fp);
  if (buf != NULL) *size = (size_t)len;
  return buf;
}

static int write_file(const char *name, const uint8_t *buf, size_t size)
{
  FILE *fp = fopen(name, "wb");
  if (fp == NULL) {
    fprintf(stderr, "Failed to open output file %s \n", name);
    return 1;
  }
  const int bad = fwrite(buf, 1, size, fp) != size;
  if (fclose(fp) != 0 || bad) {
    fprintf(stderr, "Warning: failed to write %s \n", name);
    return 1;
  }
  return 0;
}

/* a WAV file: image + parsed info + its samples as interleaved int16 (top 16 bits) */
struct wav_input {
  uint8_t *image;
  size_t size;
  struct aadwav_info info;
  int16_t *pcm16;        /* points into image for 16-bit files, else owned */
  int owns_pcm16;
};

static void wav_input_release(struct wav_input *w)
{
  if (w->owns_pcm16) io_free(w->pcm16);
  io_free(w->image);
  memset(w, 0, sizeof(*w));
}

/* need_pcm16: also provide the samples as interleaved int16 (host conversion for 8 / 24 / 32-bit files); paths that
 * hand the data chunk to the device as it is (AADGpu_EncodeWav, AADGpu_AnalyzeWav) do not need it */
static int wav_input_open(const char *name, struct wav_input *w, int need_pcm16)
{
  memset(w, 0, sizeof(*w));
  w->image = read_file(name, &w->size);
  if (w->image == NULL || aadwav_parse(w->image, w->size, &w->info) != AADWAV_OK) {
    fprintf(stderr, "Failed to open %s. \n", name);
    if (w->image) io_free(w->image);
    w->image = NULL;
    return 1;
  }
  const size_t count = (size_t)w->info.num_samples * w->info.num_channels;
  const uint8_t *data = w->image + w->info.data_offset;
  if (w->info.bits_per_sample == 16 && ((uintptr_t)data & 1u) == 0) {
    w->pcm16 = (int16_t *)(void *)data;            /* little-endian host: the data chunk IS the int16 array */
  } else if (need_pcm16) {
    w->pcm16 = (int16_t *)io_alloc(count * sizeof(int16_t));
    if (w->pcm16 == NULL) { wav_input_release(w); return 1; }
    w->owns_pcm16 = 1;
    aadwav_to_pcm16(data, w->info.bits_per_sample, count, w->pcm16);
  }
  return 0;
}

static void fill_parameter(struct AADEncodeParameter *p, const struct AADEncodeParameter *cli, const struct aadwav_info *info)
{
  *p = *cli;
  p->num_channels = (uint16_t)info->num_channels;
  p->sampling_rate = info->sampling_rate;
}

/* ---- modes ------------------------------------------------------------------------------------ */

static int execute_information(const char *name)
{
  static const char *const method[] = { "None", "MS-Conversion" };
  uint8_t header_bytes[AAD_HEADER_SIZE];
  struct AADHeaderInfo h;
  FILE *fp = fopen(name, "rb");
  if (fp == NULL) {
    fprintf(stderr, "Failed to open %s. \n", name);
    return 1;
  }
  const size_t got = fread(header_bytes, 1, AAD_HEADER_SIZE, fp);
  fclose(fp);
  if (got < AAD_HEADER_SIZE) {
    fprintf(stderr, "Failed to read from %s. \n", name);
    return 1;
  }
  const AADApiResult r = AADDecoder_DecodeHeader(header_bytes, AAD_HEADER_SIZE, &h);
  if (r != AAD_APIRESULT_OK) {
    fprintf(stderr, "Failed to read header. API result: %d \n", r);
    return 1;
  }
  /* src/main.c:260-269 */
  printf("%-30s %-9d   \n", "Format Version:", (int)h.format_version);
  printf("%-30s %-9d   \n", "Codec Version:", (int)h.codec_version);
  printf("%-30s %-9d   \n", "Number of Channels:", (int)h.num_channels);
  printf("%-30s %-9d   \n", "Number of Samples per Channel:", (int)h.num_samples);
  printf("%-30s %-9d   \n", "Sampling Rate:", (int)h.sampling_rate);
  printf("%-30s %-9d   \n", "Bits per Sample:", (int)h.bits_per_sample);
  printf("%-30s %-9d   \n", "Block size:", (int)h.block_size);
  printf("%-30s %-9d   \n", "Number of Samples per Block:", (int)h.num_samples_per_block);
  printf("%-30s %-9s   \n", "Channel Processing:",
         (unsigned)h.ch_process_method < 2u ? method[h.ch_process_method] : "Invalid");
  printf("%-30s %-8.1f \n", "Bits per Second(bps):",
         (8.0f * (double)h.block_size * h.sampling_rate) / h.num_samples_per_block);
  return 0;
}

static struct AADGpuGroup *g_group = NULL;   /* set when --device names more than one GPU */

static int execute_encode(struct AADGpu *gpu, const char *in_name, const char *out_name, const struct AADEncodeParameter *cli)
{
  struct wav_input w;
  struct AADEncodeParameter prm;
  /* several devices share ONE file's encode only in segment mode (--segment-blocks); otherwise one device does it */
  const uint32_t seg = AADGpu_GetEncodeSegmentBlocks(gpu);
  const int shared = (g_group != NULL && seg != 0);
  if (wav_input_open(in_name, &w, shared) != 0) return 1;
  fill_parameter(&prm, cli, &w.info);
  const uint64_t bound = AADGpu_StreamBytesBound(&prm, w.info.num_samples);
  if (bound == 0 || bound > 0xFFFFFFFFull) {
    fprintf(stderr, "Failed to set encode parameter. Please check encode parameter. \n");
    wav_input_release(&w);
    return 1;
  }
  uint8_t *data = (uint8_t *)io_alloc((size_t)bound);
  uint32_t out_size = 0;
  int rc = 1;
  if (data != NULL) {
    /* one device: the data chunk goes up as it lies in the file, whatever its bit depth (src/main.c:175-179 on the device) */
    const AADApiResult r = shared
        ? AADGpuGroup_EncodeInterleaved16(g_group, &prm, seg, w.pcm16, w.info.num_samples, data, (uint32_t)bound, &out_size)
        : AADGpu_EncodeWav(gpu, &prm, w.image + w.info.data_offset, w.info.bits_per_sample, w.info.num_samples, data,
                           (uint32_t)bound, &out_size);
    if (r != AAD_APIRESULT_OK) fprintf(stderr, "Failed to encode. API result:%d %s\n", r, AADGpu_LastError());
    else rc = write_file(out_name, data, out_size);
    io_free(data);
  }
  wav_input_release(&w);
  return rc;
}

static int execute_decode(struct AADGpu *gpu, const char *in_name, const char *out_name)
{
  size_t size = 0;
  uint8_t *data = read_file(in_name, &size);
  struct AADHeaderInfo h;
  if (data == NULL) {
    fprintf(stderr, "Failed to open %s. \n", in_name);
    return 1;
  }
  AADApiResult r = AADDecoder_DecodeHeader(data, size > 0xFFFFFFFFu ? 0xFFFFFFFFu : (uint32_t)size, &h);
  if (r != AAD_APIRESULT_OK) {
    fprintf(stderr, "Failed to read header. API result: %d \n", r);
    io_free(data);
    return 1;
  }
  /* validate the header before it sizes an allocation: AADDecoder_SetHeader applies the format checks of
   * src/aad_decoder.c:173-225 (what the reference's DecodeWhole does first, src/aad_decoder.c:499) */
  {
    struct AADDecoder *probe = AADDecoder_Create(NULL, 0);
    r = (probe != NULL) ? AADDecoder_SetHeader(probe, &h) : AAD_APIRESULT_NG;
    AADDecoder_Destroy(probe);
    if (r != AAD_APIRESULT_OK) {
      fprintf(stderr, "Failed to decode. API result: %d \n", r);
      io_free(data);
      return 1;
    }
  }
  /* the output file image: 44-byte header, then the decoder's interleaved int16 land in place */
  const size_t count = (size_t)h.num_channels * h.num_samples;
  uint8_t *image = (uint8_t *)io_alloc(AADWAV_HEADER_BYTES + count * 2 + 2);
  int rc = 1;
  if (image != NULL) {
    aadwav_write_header(image, h.num_channels, h.sampling_rate, 16, h.num_samples);
    int16_t *samples = (int16_t *)(void *)(image + AADWAV_HEADER_BYTES);
    r = g_group ? AADGpuGroup_DecodeInterleaved16(g_group, data, (uint32_t)size, samples, h.num_samples)   /* by block range */
                : AADGpu_DecodeInterleaved16(gpu, data, (uint32_t)size, samples, h.num_samples);
    if (r != AAD_APIRESULT_OK) fprintf(stderr, "Failed to decode. API result: %d %s\n", r, AADGpu_LastError());
    else rc = write_file(out_name, image, AADWAV_HEADER_BYTES + count * 2);
    io_free(image);
  }
  io_free(data);
  return rc;
}

/* -r, -g, -c: src/main.c:349-503 around execute_reconstruction_core */
static int execute_analysis(struct AADGpu *gpu, int mode, const char *in_name, const char *out_name,
                            const struct AADEncodeParameter *cli)
{
  /* src/main.c:275-503: the data chunk goes to the device as it lies in the file; narrowing to 16 bits, encode,
   * decode and the per-sample arithmetic of the mode run there (AADGpu_AnalyzeWav) */
  struct wav_input w;
  struct AADEncodeParameter prm;
  if (wav_input_open(in_name, &w, 0) != 0) return 1;
  fill_parameter(&prm, cli, &w.info);
  const uint32_t bits = w.info.bits_per_sample, C = w.info.num_channels, n = w.info.num_samples;
  const size_t count = (size_t)n * C;
  const uint8_t *in_data = w.image + w.info.data_offset;
  const size_t bytes = AADWAV_HEADER_BYTES + count * (bits / 8);
  uint8_t *image = (mode == MODE_CALCULATE) ? NULL : (uint8_t *)io_alloc(bytes);
  double stats[3] = { 0.0, 0.0, 0.0 };
  int rc = 1;
  if (mode != MODE_CALCULATE && image == NULL) { wav_input_release(&w); return 1; }
  const enum AADGpuAnalysis what = (mode == MODE_CALCULATE) ? AADGPU_ANALYSIS_STATISTICS
                                 : (mode == MODE_GAP) ? AADGPU_ANALYSIS_RESIDUAL : AADGPU_ANALYSIS_RECONSTRUCT;
  const AADApiResult r = AADGpu_AnalyzeWav(gpu, &prm, in_data, bits, n, what, image ? image + AADWAV_HEADER_BYTES : NULL, stats, NULL);
  if (r == AAD_APIRESULT_INVALID_FORMAT) {
    fprintf(stderr, "Failed to set encode parameter. Please check encode parameter. \n");
  } else if (r != AAD_APIRESULT_OK) {
    fprintf(stderr, "Failed to encode. API result:%d %s\n", r, AADGpu_LastError());
  } else if (mode == MODE_CALCULATE) {
    printf("RMSE:%f MSD:%f MaxAE:%f \n", stats[0], stats[1], stats[2]);   /* src/main.c:492-495 */
    rc = 0;
  } else {
    /* the output keeps the input's format: src/main.c:372-381, :418-428 */
    aadwav_write_header(image, C, w.info.sampling_rate, bits, n);
    rc = write_file(out_name, image, bytes);
  }
  if (image != NULL) io_free(image);
  wav_input_release(&w);
  return rc;
}

/* ---- --batch: many files, one launch per shape ------------------------------------------------ */

struct batch_item {
  char *in_name, *out_name;
  struct wav_input wav;       /* -e */
  uint8_t *aad;               /* -d */
  size_t aad_size;
  struct AADHeaderInfo header;
  int done, failed;
};

static int read_manifest(const char *name, struct batch_item **items_out, size_t *count_out)
{
  FILE *fp = fopen(name, "r");
  char line[8192];
  size_t count = 0, cap = 0;
  struct batch_item *items = NULL;
  if (fp == NULL) {
    fprintf(stderr, "Failed to open %s. \n", name);
    return 1;
  }
  while (fgets(line, sizeof(line), fp) != NULL) {
    char a[4096], b[4096];
    if (sscanf(line, "%4095s %4095s", a, b) != 2) continue;     /* blank or malformed lines are skipped */
    if (a[0] == '#') continue;
    if (count == cap) {
      cap = cap ? 2 * cap : 64;
      struct batch_item *grown = (struct batch_item *)realloc(items, cap * sizeof(*items));
      if (grown == NULL) { free(items); fclose(fp); return 1; }
      items = grown;
    }
    memset(&items[count], 0, sizeof(items[count]));
    items[count].in_name = strdup(a);
    items[count].out_name = strdup(b);
    count++;
  }
  fclose(fp);
  *items_out = items;
  *count_out = count;
  return 0;
}

static int execute_encode_batch(struct AADGpu *gpu, const char *manifest, const struct AADEncodeParameter *cli)
{
  struct batch_item *items = NULL;
  size_t count = 0;
  int failures = 0;
  if (read_manifest(manifest, &items, &count) != 0) return 1;
  for (size_t i = 0; i < count; i++)
    if (wav_input_open(items[i].in_name, &items[i].wav, 1) != 0) { items[i].done = items[i].failed = 1; failures++; }
  /* one AADGpu_EncodeBatch per (channels, sampling rate): lengths may be ragged inside a batch */
  for (size_t lead = 0; lead < count; lead++) {
    if (items[lead].done) continue;
    const struct aadwav_info *li = &items[lead].wav.info;
    size_t members = 0;
    uint32_t longest = 0;
    for (size_t i = lead; i < count; i++) {
      const struct aadwav_info *ii = &items[i].wav.info;
      if (items[i].done || ii->num_channels != li->num_channels || ii->sampling_rate != li->sampling_rate) continue;
      members++;
      if (ii->num_samples > longest) longest = ii->num_samples;
    }
    struct AADGpuBatch b;
    memset(&b, 0, sizeof(b));
    fill_parameter(&b.param, cli, li);
    b.num_streams = (uint32_t)members;
    b.num_samples = longest;
    b.pcm_channel_stride = ((uint64_t)longest + 63u) & ~(uint64_t)63u;
    b.pcm_stream_stride = b.pcm_channel_stride * li->num_channels;
    b.aad_stream_stride = AADGpu_StreamBytesBound(&b.param, longest);
    const uint32_t C = li->num_channels;
    int16_t *pcm = (b.aad_stream_stride != 0) ? (int16_t *)io_alloc((size_t)members * b.pcm_stream_stride * 2) : NULL;
    uint8_t *aad = pcm ? (uint8_t *)io_alloc((size_t)members * b.aad_stream_stride) : NULL;
    uint32_t *lens = (uint32_t *)malloc(members * sizeof(uint32_t)), *sizes = (uint32_t *)malloc(members * sizeof(uint32_t));
    AADApiResult r = AAD_APIRESULT_INVALID_FORMAT;
    if (b.aad_stream_stride == 0) {
      fprintf(stderr, "Failed to set encode parameter. Please check encode parameter. \n");
    } else if (pcm == NULL || aad == NULL || lens == NULL || sizes == NULL) {
      r = AAD_APIRESULT_NG;
    } else {
      size_t m = 0;
      for (size_t i = lead; i < count; i++) {      /* WAV order -> planar rows of the batch */
        const struct aadwav_info *ii = &items[i].wav.info;
        if (items[i].done || ii->num_channels != C || ii->sampling_rate != li->sampling_rate) continue;
        int16_t *dst = pcm + m * b.pcm_stream_stride;
        for (uint32_t c = 0; c < C; c++)
          for (uint32_t s = 0; s < ii->num_samples; s++) dst[c * b.pcm_channel_stride + s] = items[i].wav.pcm16[(size_t)s * C + c];
        lens[m++] = ii->num_samples;
      }
      r = g_group ? AADGpuGroup_EncodeBatch(g_group, &b, pcm, lens, aad, sizes) : AADGpu_EncodeBatch(gpu, &b, pcm, lens, aad, sizes);
      if (r != AAD_APIRESULT_OK) fprintf(stderr, "Failed to encode. API result:%d %s\n", r, AADGpu_LastError());
    }
    size_t m = 0;
    for (size_t i = lead; i < count; i++) {
      const struct aadwav_info *ii = &items[i].wav.info;
      if (items[i].done || ii->num_channels != C || ii->sampling_rate != li->sampling_rate) continue;
      items[i].done = 1;
      if (r != AAD_APIRESULT_OK || write_file(items[i].out_name, aad + m * b.aad_stream_stride, sizes[m]) != 0) {
        items[i].failed = 1;
        failures++;
      }
      m++;
    }
    if (pcm) io_free(pcm);
    if (aad) io_free(aad);
    free(lens);
    free(sizes);
  }
  for (size_t i = 0; i < count; i++) {
    if (items[i].wav.image) wav_input_release(&items[i].wav);
    free(items[i].in_name);
    free(items[i].out_name);
  }
  free(items);
  return failures ? 1 : 0;
}

static int same_stream_shape(const struct AADHeaderInfo *a, const struct AADHeaderInfo *b)
{
  return a->num_channels == b->num_channels && a->bits_per_sample == b->bits_per_sample && a->block_size == b->block_size &&
         a->num_samples_per_block == b->num_samples_per_block && a->ch_process_method == b->ch_process_method;
}

static int execute_decode_batch(struct AADGpu *gpu, const char *manifest)
{
  struct batch_item *items = NULL;
  size_t count = 0;
  int failures = 0;
  if (read_manifest(manifest, &items, &count) != 0) return 1;
  for (size_t i = 0; i < count; i++) {
    items[i].aad = read_file(items[i].in_name, &items[i].aad_size);
    AADApiResult r = AAD_APIRESULT_NG;
    if (items[i].aad == NULL) fprintf(stderr, "Failed to open %s. \n", items[i].in_name);
    else if ((r = AADDecoder_DecodeHeader(items[i].aad, (uint32_t)items[i].aad_size, &items[i].header)) != AAD_APIRESULT_OK)
      fprintf(stderr, "Failed to read header. API result: %d \n", r);
    if (r != AAD_APIRESULT_OK) { items[i].done = items[i].failed = 1; failures++; }
  }
  for (size_t lead = 0; lead < count; lead++) {
    if (items[lead].done) continue;
    const struct AADHeaderInfo *lh = &items[lead].header;
    size_t members = 0, widest = 0;
    uint32_t longest = 0;
    for (size_t i = lead; i < count; i++) {
      if (items[i].done || !same_stream_shape(&items[i].header, lh)) continue;
      members++;
      if (items[i].header.num_samples > longest) longest = items[i].header.num_samples;
      if (items[i].aad_size > widest) widest = items[i].aad_size;
    }
    struct AADGpuBatch b;
    memset(&b, 0, sizeof(b));
    b.param.num_channels = lh->num_channels;
    b.param.sampling_rate = lh->sampling_rate;
    b.param.bits_per_sample = lh->bits_per_sample;
    b.param.max_block_size = lh->block_size;
    b.param.ch_process_method = lh->ch_process_method;
    b.num_streams = (uint32_t)members;
    b.num_samples = longest;
    b.pcm_channel_stride = ((uint64_t)longest + 63u) & ~(uint64_t)63u;
    b.pcm_stream_stride = b.pcm_channel_stride * lh->num_channels;
    const uint64_t bound = AADGpu_StreamBytesBound(&b.param, longest);
    b.aad_stream_stride = bound > widest ? bound : widest;
    const uint32_t C = lh->num_channels;
    uint8_t *aad = (bound != 0) ? (uint8_t *)io_alloc((size_t)members * b.aad_stream_stride) : NULL;
    int16_t *pcm = aad ? (int16_t *)io_alloc((size_t)members * b.pcm_stream_stride * 2) : NULL;
    uint32_t *sizes = (uint32_t *)malloc(members * sizeof(uint32_t));
    AADApiResult r = AAD_APIRESULT_INVALID_FORMAT;
    if (aad != NULL && pcm != NULL && sizes != NULL) {
      size_t m = 0;
      for (size_t i = lead; i < count; i++) {
        if (items[i].done || !same_stream_shape(&items[i].header, lh)) continue;
        memcpy(aad + m * b.aad_stream_stride, items[i].aad, items[i].aad_size);
        sizes[m++] = (uint32_t)items[i].aad_size;
      }
      r = g_group ? AADGpuGroup_DecodeBatch(g_group, &b, aad, sizes, pcm) : AADGpu_DecodeBatch(gpu, &b, aad, sizes, pcm);
    }
    if (r != AAD_APIRESULT_OK) fprintf(stderr, "Failed to decode. API result: %d %s\n", r, AADGpu_LastError());
    size_t m = 0;
    for (size_t i = lead; i < count; i++) {
      if (items[i].done || !same_stream_shape(&items[i].header, lh)) continue;
      items[i].done = 1;
      int bad = (r != AAD_APIRESULT_OK);
      if (!bad) {
        const uint32_t n = items[i].header.num_samples;
        const size_t bytes = AADWAV_HEADER_BYTES + (size_t)n * C * 2;
        uint8_t *image = (uint8_t *)malloc(bytes);
        if (image == NULL) bad = 1;
        else {
          aadwav_write_header(image, C, items[i].header.sampling_rate, 16, n);
          const int16_t *src = pcm + m * b.pcm_stream_stride;
          for (uint32_t s = 0; s < n; s++)
            for (uint32_t c = 0; c < C; c++) aadwav_store32(image + AADWAV_HEADER_BYTES, 16, (size_t)s * C + c,
                                                            (int32_t)((uint32_t)(int32_t)src[c * b.pcm_channel_stride + s] << 16));
          bad = write_file(items[i].out_name, image, bytes);
          free(image);
        }
      }
      if (bad) { items[i].failed = 1; failures++; }
      m++;
    }
    if (aad) io_free(aad);
    if (pcm) io_free(pcm);
    free(sizes);
  }
  for (size_t i = 0; i < count; i++) {
    if (items[i].aad) io_free(items[i].aad);
    free(items[i].in_name);
    free(items[i].out_name);
  }
  free(items);
  return failures ? 1 : 0;
}

/* ---- main -------------------------------------------------------------------------------------- */

int main(int argc, char **argv)
{
  struct options o;
  struct AADEncodeParameter cli;
  if (argc == 1) {
    print_usage(argv[0]);
    printf("type `%s -h` to display usage. \n", argv[0]);
    return 1;
  }
  if (parse_arguments(argc, argv, &o) != 0) return 1;
  if (o.help) {
    print_help(argv[0]);
    return 0;
  }
  if (o.version) {
    printf("AAD(Ayashi Adaptive Differential pulse code modulation) encoder/decoder Version.%d (B200 build) \n", AAD_CODEC_VERSION);
    return 0;
  }
  int modes = 0, mode = -1;
  for (int m = 0; m < NUM_MODES; m++)
    if (o.mode_set[m]) { modes++; mode = m; }
  if (modes == 0) {
    fprintf(stderr, "%s: must specify at least one mode. \n", argv[0]);
    return 1;
  }
  if (modes >= 2) {
    fprintf(stderr, "%s: multiple modes cannot specify simultaneously. \n", argv[0]);
    return 1;
  }
  const int batch = o.batch != NULL;
  if (batch && mode != MODE_ENCODE && mode != MODE_DECODE) {
    fprintf(stderr, "%s: --batch goes with -e or -d. \n", argv[0]);
    return 1;
  }
  if (!batch && o.num_files < 1) {
    fprintf(stderr, "%s: input file must be specified. \n", argv[0]);
    return 1;
  }
  memset(&cli, 0, sizeof(cli));
  cli.bits_per_sample = (uint16_t)(uint8_t)strtol(o.bits ? o.bits : "4", NULL, 10);
  cli.max_block_size = (uint16_t)strtol(o.block ? o.block : "1024", NULL, 10);
  cli.num_encode_trials = (uint8_t)strtol(o.trials ? o.trials : "2", NULL, 10);
  cli.ch_process_method = o.ms ? AAD_CH_PROCESS_METHOD_MS : AAD_CH_PROCESS_METHOD_NONE;

  if (mode == MODE_INFORMATION) return execute_information(o.files[0]);
  if (!batch && mode != MODE_CALCULATE && o.num_files < 2) {
    fprintf(stderr, "%s: output file must be specified. \n", argv[0]);
    return 1;
  }

  /* --device: one index, a list, or "all" */
  int devices[16], num_devices = 0;
  const char *spec = o.device ? o.device : "0";
  if (strcmp(spec, "all") == 0) {
    num_devices = AADGpu_DeviceCount();
    if (num_devices > 16) num_devices = 16;
    for (int d = 0; d < num_devices; d++) devices[d] = d;
  } else {
    const char *p = spec;
    while (*p != '\0' && num_devices < 16) {
      char *end = NULL;
      const long v = strtol(p, &end, 10);
      if (end == p || v < 0) { num_devices = 0; break; }
      devices[num_devices++] = (int)v;
      p = (*end == ',') ? end + 1 : end;
      if (*end != ',' && *end != '\0') { num_devices = 0; break; }
    }
  }
  if (num_devices == 0) {
    fprintf(stderr, "%s: no usable CUDA device in \"%s\" (%d visible; this program has no CPU fallback) \n", argv[0], spec,
            AADGpu_DeviceCount());
    return 1;
  }
  struct AADGpu *gpu = NULL;
  if (num_devices > 1) {
    g_group = AADGpuGroup_Create(devices, num_devices);
    gpu = AADGpuGroup_Device(g_group, 0);
  } else {
    gpu = AADGpu_Create(devices[0]);
  }
  if (gpu == NULL) {
    fprintf(stderr, "%s: %s \n", argv[0], AADGpu_LastError());
    return 1;
  }
  g_have_gpu = 1;
  if (o.segment) {   /* extension: segment-parallel encoding on every device this run uses */
    const uint32_t seg = (uint32_t)strtoul(o.segment, NULL, 10);
    for (int d = 0; d < (g_group ? AADGpuGroup_Size(g_group) : 1); d++)
      AADGpu_SetEncodeSegmentBlocks(g_group ? AADGpuGroup_Device(g_group, d) : gpu, seg);
  }
  int rc;
  if (batch) rc = (mode == MODE_ENCODE) ? execute_encode_batch(gpu, o.batch, &cli) : execute_decode_batch(gpu, o.batch);
  else if (mode == MODE_ENCODE) rc = execute_encode(gpu, o.files[0], o.files[1], &cli);
  else if (mode == MODE_DECODE) rc = execute_decode(gpu, o.files[0], o.files[1]);
  else rc = execute_analysis(gpu, mode, o.files[0], o.files[1], &cli);
  if (g_group) AADGpuGroup_Destroy(g_group);
  else AADGpu_Destroy(gpu);
  return rc;
}
