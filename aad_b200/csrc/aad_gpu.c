/*
 * aad_gpu.c -- host side of libaad_b200.so, in C: device context, scratch memory, and the
 * pipelines that move streams host -> device -> host around the kernels in aad_kernels.cu.
 *
 * Pipelining: the whole batch stays resident on the device (12,500 ten-second clips are
 * ~14 GB of the 180 GB HBM); the COPIES are sliced by block range.  Slice k's H2D runs on
 * one stream while slice k-1's kernel runs on a second and slice k-2's D2H on a third.  The
 * encoder's chain state crosses slices through a device-resident state array, exactly the way
 * the reference carries its per-channel processor from block to block
 * (src/aad_encoder.c:21,853-886) -- so slicing never changes a byte of output.
 */
#define _GNU_SOURCE
#include "aad_gpu_internal.h"
#include "aad_decoder.h"

#include <ctype.h>
#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

static __thread char tl_error[320] = "";
static uint32_t g_max_channels = AADF_MAX_CHANNELS;

void aadgpu_set_error(const char *msg) { snprintf(tl_error, sizeof(tl_error), "%s", msg ? msg : ""); }

AADApiResult aadgpu_fail(const char *what, cudaError_t err)
{
  snprintf(tl_error, sizeof(tl_error), "%s: %s", what, cudaGetErrorString(err));
  return AAD_APIRESULT_NG;
}

const char *AADGpu_LastError(void) { return tl_error; }
uint64_t AADGpu_KernelLaunchCount(void) { return aadk_launch_count(); }
uint64_t AADGpu_TmaLaunchCount(void) { return aadk_tma_launch_count(); }
void AADGpu_SetKernelPath(int path) { aadk_force_generic(path); }
void AADGpu_SetEncoderPairing(int on) { aadk_set_encoder_schedule(on ? 1 : 0); }
void AADGpu_SetEncoderSchedule(int mode) { aadk_set_encoder_schedule(mode); }
uint32_t aadgpu_max_channels(void) { return g_max_channels; }
uint32_t AADGpu_GetMaxChannels(void) { return g_max_channels; }
void AADGpu_SetMaxChannels(uint32_t n)
{
  if (n >= 1 && n <= AADF_MAX_CHANNELS) g_max_channels = n;
}

/* A failed CUDA call ends the entry point -- but never with copies or kernels still in flight on the caller's
 * host buffers or on the context's scratch memory: the three pipeline streams are drained first, and the
 * (possibly sticky) error state is read so the next call starts clean.  Every use has `gpu` in scope. */
static AADApiResult aadgpu_fail_drained(struct AADGpu *g, const char *what, cudaError_t err)
{
  if (g) {
    if (g->s_in) (void)cudaStreamSynchronize(g->s_in);
    if (g->s_run) (void)cudaStreamSynchronize(g->s_run);
    if (g->s_out) (void)cudaStreamSynchronize(g->s_out);
    if (g->s_out2) (void)cudaStreamSynchronize(g->s_out2);
  }
  (void)cudaGetLastError();
  return aadgpu_fail(what, err);
}

#define CU(call, what)                                  \
  do {                                                  \
    cudaError_t e__ = (call);                           \
    if (e__ != cudaSuccess) return aadgpu_fail_drained(gpu, what, e__); \
  } while (0)

/* Every entry point that uses a context's streams and scratch buffers runs under that context's lock
 * (aad_gpu_internal.h): the body is the *_unlocked function of the same name. */
#define WITH_CONTEXT_LOCK(gpu, call)                               \
  do {                                                             \
    if (!(gpu)) return AAD_APIRESULT_INVALID_ARGUMENT;             \
    pthread_mutex_lock(&(gpu)->lock);                              \
    const AADApiResult locked_result_ = (call);                    \
    pthread_mutex_unlock(&(gpu)->lock);                            \
    return locked_result_;                                         \
  } while (0)

int AADGpu_DeviceCount(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return n;
}

struct AADGpu *AADGpu_Create(int device)
{
  if (device < 0 || device >= AADGpu_DeviceCount()) {
    aadgpu_set_error("AADGpu_Create: no such CUDA device (this library has no CPU fallback)");
    return NULL;
  }
  if (cudaSetDevice(device) != cudaSuccess) {
    aadgpu_fail("cudaSetDevice", cudaGetLastError());
    return NULL;
  }
  struct AADGpu *g = (struct AADGpu *)calloc(1, sizeof(*g));
  if (!g) return NULL;
  g->device = device;
  pthread_mutex_init(&g->lock, NULL);
  cudaError_t e = cudaStreamCreateWithFlags(&g->s_in, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g->s_run, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g->s_out, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g->s_out2, cudaStreamNonBlocking);
  for (int i = 0; i < AADGPU_MAX_SLICES && e == cudaSuccess; i++) {
    e = cudaEventCreateWithFlags(&g->ev_in[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->ev_run[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->ev_out[i], cudaEventDisableTiming);
  }
  if (e != cudaSuccess) {
    aadgpu_fail("AADGpu_Create", e);
    AADGpu_Destroy(g);
    return NULL;
  }
  return g;
}

void AADGpu_Destroy(struct AADGpu *g)
{
  if (!g) return;
  cudaSetDevice(g->device);
  cudaDeviceSynchronize();
  struct aadgpu_buffer *bufs[] = { &g->pcm, &g->aad, &g->state, &g->lens, &g->sizes, &g->lut, &g->wav, &g->pcm2, &g->raw, &g->stats };
  for (size_t i = 0; i < sizeof(bufs) / sizeof(bufs[0]); i++)
    if (bufs[i]->ptr) cudaFree(bufs[i]->ptr);
  for (int i = 0; i < AADGPU_MAX_SLICES; i++) {
    if (g->ev_in[i]) cudaEventDestroy(g->ev_in[i]);
    if (g->ev_run[i]) cudaEventDestroy(g->ev_run[i]);
    if (g->ev_out[i]) cudaEventDestroy(g->ev_out[i]);
  }
  for (int i = 0; i < 3; i++) {
    if (g->ring_in[i]) cudaFreeHost(g->ring_in[i]);
    if (g->ring_out[i]) cudaFreeHost(g->ring_out[i]);
  }
  if (g->s_in) cudaStreamDestroy(g->s_in);
  if (g->s_run) cudaStreamDestroy(g->s_run);
  if (g->s_out) cudaStreamDestroy(g->s_out);
  if (g->s_out2) cudaStreamDestroy(g->s_out2);
  pthread_mutex_destroy(&g->lock);
  free(g);
}

static pthread_mutex_t g_default_lock = PTHREAD_MUTEX_INITIALIZER;
static struct AADGpu *g_default = NULL;

struct AADGpu *aadgpu_default(void)
{
  pthread_mutex_lock(&g_default_lock);
  if (!g_default) {
    const char *env = getenv("AAD_B200_DEVICE");
    g_default = AADGpu_Create(env ? atoi(env) : 0);
  }
  struct AADGpu *g = g_default;
  pthread_mutex_unlock(&g_default_lock);
  return g;
}

void *AADGpu_HostAlloc(size_t bytes)
{
  void *p = NULL;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
    aadgpu_fail("cudaMallocHost", cudaGetLastError());
    return NULL;
  }
  return p;
}

void AADGpu_HostFree(void *p)
{
  if (p) cudaFreeHost(p);
}

/* The CPUs next to the device (its PCIe root's NUMA node), from sysfs.  Host buffers a thread pins
 * after binding itself there are allocated on that node, so the copies of one device do not cross the
 * socket interconnect -- which is what limits a box where eight ranks pin 25 GB each. */
int AADGpu_BindHostThread(struct AADGpu *gpu)
{
  char bus[32], path[128], list[1024];
  if (!gpu) return 0;
  if (gpu->cpus_known)   /* the device's CPU list was read before: no file I/O on the per-call worker threads */
    return gpu->cpus_count > 0 && sched_setaffinity(0, sizeof(gpu->cpus), &gpu->cpus) == 0 ? gpu->cpus_count : 0;
  gpu->cpus_known = 1;
  gpu->cpus_count = 0;
  if (cudaDeviceGetPCIBusId(bus, (int)sizeof(bus), gpu->device) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  for (char *c = bus; *c; c++) *c = (char)tolower((unsigned char)*c);
  snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/local_cpulist", bus);
  FILE *fp = fopen(path, "r");
  if (!fp) return 0;
  const int ok = fgets(list, sizeof(list), fp) != NULL;
  fclose(fp);
  if (!ok) return 0;
  cpu_set_t set;
  CPU_ZERO(&set);
  int count = 0;
  for (const char *c = list; *c;) {          /* "0-15,32-47" */
    if (!isdigit((unsigned char)*c)) { c++; continue; }
    char *end;
    long a = strtol(c, &end, 10), b = a;
    if (*end == '-') b = strtol(end + 1, &end, 10);
    for (long k = a; k <= b && k < CPU_SETSIZE; k++) { CPU_SET((int)k, &set); count++; }
    c = end;
  }
  if (count == 0) return 0;
  gpu->cpus = set;
  gpu->cpus_count = count;
  return sched_setaffinity(0, sizeof(set), &set) == 0 ? count : 0;
}

int aadgpu_reserve(struct AADGpu *gpu, struct aadgpu_buffer *b, size_t bytes)
{
  (void)gpu;
  if (b->cap >= bytes && b->ptr) return 1;
  if (b->ptr) cudaFree(b->ptr);
  b->ptr = NULL;
  b->cap = 0;
  const size_t want = (bytes + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
  if (cudaMalloc(&b->ptr, want) != cudaSuccess) {
    aadgpu_fail("cudaMalloc", cudaGetLastError());
    b->ptr = NULL;
    return 0;
  }
  b->cap = want;
  return 1;
}

/* ---- host link probe ---------------------------------------------------------------------- */

static int trace_on(void);
static long env_number(const char *name, long lo, long hi, long fallback);
static double wall_seconds(void)
{
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

/* What this device's host link delivers to plain 1-D copies between pinned host memory and HBM, through the same
 * streams the pipelines use: host -> device alone, device -> host alone, and both directions at once (GB/s,
 * aggregate for the last).  bench.py runs it on every rank at the same time, so the end-to-end numbers can be
 * read against what the box's host side actually gives N devices together (tools/microbench/pcie.cu is the
 * stand-alone version). */
static AADApiResult AADGpu_LinkProbe_unlocked(struct AADGpu *gpu, size_t bytes, int repeats, double gbs[3])
{
  if (bytes == 0 || repeats <= 0 || gbs == NULL) return AAD_APIRESULT_INVALID_ARGUMENT;
  CU(cudaSetDevice(gpu->device), "cudaSetDevice");
  if (!aadgpu_reserve(gpu, &gpu->pcm, bytes)) return AAD_APIRESULT_NG;
  if (!aadgpu_reserve(gpu, &gpu->pcm2, bytes)) return AAD_APIRESULT_NG;
  void *up = NULL, *down = NULL;
  if (cudaMallocHost(&up, bytes) != cudaSuccess || cudaMallocHost(&down, bytes) != cudaSuccess) {
    if (up) cudaFreeHost(up);
    return aadgpu_fail("cudaMallocHost (link probe)", cudaGetLastError());
  }
  memset(up, 0x5A, bytes);
  memset(down, 0, bytes);
  cudaError_t e = cudaSuccess;
  for (int mode = 0; mode < 3 && e == cudaSuccess; mode++) {
    double t0 = 0.0;
    for (int r = -1; r < repeats && e == cudaSuccess; r++) {      /* r == -1: warm-up, not timed */
      if (mode != 1) e = cudaMemcpyAsync(gpu->pcm.ptr, up, bytes, cudaMemcpyHostToDevice, gpu->s_in);
      if (e == cudaSuccess && mode != 0) e = cudaMemcpyAsync(down, gpu->pcm2.ptr, bytes, cudaMemcpyDeviceToHost, gpu->s_out);
      if (r == -1) {
        if (e == cudaSuccess) e = cudaStreamSynchronize(gpu->s_in);
        if (e == cudaSuccess) e = cudaStreamSynchronize(gpu->s_out);
        t0 = wall_seconds();
      }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(gpu->s_in);
    if (e == cudaSuccess) e = cudaStreamSynchronize(gpu->s_out);
    gbs[mode] = (double)bytes * repeats * (mode == 2 ? 2.0 : 1.0) / (wall_seconds() - t0) / 1e9;
  }
  cudaFreeHost(up);
  cudaFreeHost(down);
  if (e != cudaSuccess) return aadgpu_fail_drained(gpu, "link probe", e);
  return AAD_APIRESULT_OK;
}

AADApiResult AADGpu_LinkProbe(struct AADGpu *gpu, size_t bytes, int repeats, double gbs[3])
{
  WITH_CONTEXT_LOCK(gpu, AADGpu_LinkProbe_unlocked(gpu, bytes, repeats, gbs));
}

/* The same measurement for the copies the batch pipelines issue: `rows` rows of `width` bytes, `host_pitch` bytes
 * apart in a pinned host buffer (a block-range slice of every stream of a batch), packed on the device.  Tells a
 * slow link from a slow access pattern. */
static AADApiResult AADGpu_LinkProbeRows_unlocked(struct AADGpu *gpu, size_t rows, size_t width, size_t host_pitch, int repeats,
                                                  double gbs[3])
{
  if (rows == 0 || width == 0 || host_pitch < width || repeats <= 0 || gbs == NULL) return AAD_APIRESULT_INVALID_ARGUMENT;
  CU(cudaSetDevice(gpu->device), "cudaSetDevice");
  const size_t dev_bytes = rows * width, host_bytes = rows * host_pitch;
  if (!aadgpu_reserve(gpu, &gpu->pcm, dev_bytes)) return AAD_APIRESULT_NG;
  if (!aadgpu_reserve(gpu, &gpu->pcm2, dev_bytes)) return AAD_APIRESULT_NG;
  void *up = NULL, *down = NULL;
  if (cudaMallocHost(&up, host_bytes) != cudaSuccess || cudaMallocHost(&down, host_bytes) != cudaSuccess) {
    if (up) cudaFreeHost(up);
    return aadgpu_fail("cudaMallocHost (link probe)", cudaGetLastError());
  }
  memset(up, 0x5A, host_bytes);
  memset(down, 0, host_bytes);
  cudaError_t e = cudaSuccess;
  for (int mode = 0; mode < 3 && e == cudaSuccess; mode++) {
    double t0 = 0.0;
    for (int r = -1; r < repeats && e == cudaSuccess; r++) {
      if (mode != 1) e = cudaMemcpy2DAsync(gpu->pcm.ptr, width, up, host_pitch, width, rows, cudaMemcpyHostToDevice, gpu->s_in);
      if (e == cudaSuccess && mode != 0)
        e = cudaMemcpy2DAsync(down, host_pitch, gpu->pcm2.ptr, width, width, rows, cudaMemcpyDeviceToHost, gpu->s_out);
      if (r == -1) {
        if (e == cudaSuccess) e = cudaStreamSynchronize(gpu->s_in);
        if (e == cudaSuccess) e = cudaStreamSynchronize(gpu->s_out);
        t0 = wall_seconds();
      }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(gpu->s_in);
    if (e == cudaSuccess) e = cudaStreamSynchronize(gpu->s_out);
    gbs[mode] = (double)dev_bytes * repeats * (mode == 2 ? 2.0 : 1.0) / (wall_seconds() - t0) / 1e9;
  }
  cudaFreeHost(up);
  cudaFreeHost(down);
  if (e != cudaSuccess) return aadgpu_fail_drained(gpu, "link probe (rows)", e);
  return AAD_APIRESULT_OK;
}

AADApiResult AADGpu_LinkProbeRows(struct AADGpu *gpu, size_t rows, size_t width, size_t host_pitch, int repeats, double gbs[3])
{
  WITH_CONTEXT_LOCK(gpu, AADGpu_LinkProbeRows_unlocked(gpu, rows, width, host_pitch, repeats, gbs));
}

/* ---- parameter checks ------------------------------------------------------------------- */

/* The checks AADEncoder_SetEncodeParameter (src/aad_encoder.c:741-770) and
 * AADEncoder_EncodeHeader (src/aad_encoder.c:149-185) apply, in that order. */
/* ---- segment mode ---------------------------------------------------------------------------- */

AADApiResult AADGpu_SetEncodeSegmentBlocks(struct AADGpu *gpu, uint32_t blocks)
{
  if (!gpu) return AAD_APIRESULT_INVALID_ARGUMENT;
  gpu->segment_blocks = blocks;
  return AAD_APIRESULT_OK;
}

uint32_t AADGpu_GetEncodeSegmentBlocks(const struct AADGpu *gpu) { return gpu ? gpu->segment_blocks : 0; }

/* chains per (stream, channel) for a stream of total_blocks blocks; fills the kernel parameters */
static uint32_t apply_segments(const struct AADGpu *gpu, struct aadk_encode_params *p, uint32_t total_blocks)
{
  p->segment_blocks = gpu->segment_blocks;
  p->num_segments = gpu->segment_blocks ? (total_blocks + gpu->segment_blocks - 1) / gpu->segment_blocks : 1;
  if (p->num_segments == 0) p->num_segments = 1;
  return p->num_segments;
}

static AADApiResult check_encode_shape(const struct AADEncodeParameter *prm, uint32_t num_samples,
                                       struct aadf_geometry *geo)
{
  uint32_t bs = 0, spb = 0;
  if (prm->bits_per_sample == 0 || prm->bits_per_sample > AAD_MAX_BITS_PER_SAMPLE) return AAD_APIRESULT_INVALID_FORMAT;
  if (prm->max_block_size < AADF_CHANNEL_HEADER_BYTES * (uint32_t)prm->num_channels) return AAD_APIRESULT_INVALID_FORMAT;
  if ((uint32_t)prm->ch_process_method >= (uint32_t)AAD_CH_PROCESS_METHOD_INVALID) return AAD_APIRESULT_INVALID_FORMAT;
  if (!aadf_block_geometry(prm->max_block_size, prm->num_channels, prm->bits_per_sample, g_max_channels, &bs, &spb))
    return AAD_APIRESULT_INVALID_FORMAT;
  if (num_samples == 0 || prm->sampling_rate == 0) return AAD_APIRESULT_INVALID_FORMAT;
  if (prm->bits_per_sample < AAD_MIN_BITS_PER_SAMPLE) return AAD_APIRESULT_INVALID_FORMAT;
  if (bs <= AADF_CHANNEL_HEADER_BYTES * (uint32_t)prm->num_channels) return AAD_APIRESULT_INVALID_FORMAT;
  if (prm->ch_process_method == AAD_CH_PROCESS_METHOD_MS && prm->num_channels == 1) return AAD_APIRESULT_INVALID_FORMAT;
  geo->channels = prm->num_channels;
  geo->bits = prm->bits_per_sample;
  geo->block_size = bs;
  geo->samples_per_block = spb;
  geo->ms = (prm->ch_process_method == AAD_CH_PROCESS_METHOD_MS) ? 1u : 0u;
  return AAD_APIRESULT_OK;
}

uint64_t AADGpu_StreamBytesBound(const struct AADEncodeParameter *prm, uint32_t num_samples)
{
  struct aadf_geometry geo;
  if (!prm || check_encode_shape(prm, num_samples ? num_samples : 1, &geo) != AAD_APIRESULT_OK) return 0;
  return aadf_stream_bytes_bound(num_samples, geo.block_size, geo.samples_per_block);
}

uint64_t AADGpu_StreamBytes(const struct AADEncodeParameter *prm, uint32_t num_samples)
{
  struct aadf_geometry geo;
  if (!prm || check_encode_shape(prm, num_samples ? num_samples : 1, &geo) != AAD_APIRESULT_OK) return 0;
  return aadf_stream_bytes(num_samples, geo.channels, geo.bits, geo.block_size, geo.samples_per_block);
}

static AADApiResult check_batch(const struct AADGpuBatch *b, struct aadf_geometry *geo)
{
  const AADApiResult r = check_encode_shape(&b->param, b->num_samples, geo);
  if (r != AAD_APIRESULT_OK) return r;
  if (b->pcm_channel_stride < b->num_samples) return AAD_APIRESULT_INSUFFICIENT_BUFFER;
  if (b->pcm_stream_stride < b->pcm_channel_stride * (geo->channels - 1) + b->num_samples)
    return AAD_APIRESULT_INSUFFICIENT_BUFFER;
  if (b->aad_stream_stride < aadf_stream_bytes_bound(b->num_samples, geo->block_size, geo->samples_per_block))
    return AAD_APIRESULT_INSUFFICIENT_BUFFER;
  if (b->aad_stream_stride > 0xFFFFFFFFull) return AAD_APIRESULT_INVALID_ARGUMENT;
  return AAD_APIRESULT_OK;
}

/* ---- device-resident entry points -------------------------------------------------------- */

AADApiResult AADGpu_EncodeBatchDevice(struct AADGpu *gpu, const struct AADGpuBatch *batch, const int16_t *pcm_dev,
                                      const uint32_t *num_samples_dev, uint8_t *aad_dev, uint32_t *out_sizes_dev,
                                      void *stream)
{
  if (!gpu || !batch || !pcm_dev || !aad_dev) return AAD_APIRESULT_INVALID_ARGUMENT;
  struct aadf_geometry geo;
  const AADApiResult r = check_batch(batch, &geo);
  if (r != AAD_APIRESULT_OK) return r;
  if (batch->num_streams == 0) return AAD_APIRESULT_OK;
  CU(cudaSetDevice(gpu->device), "cudaSetDevice");
  struct aadk_encode_params p;
  memset(&p, 0, sizeof(p));
  p.pcm = pcm_dev;
  p.pcm_clip_stride = batch->pcm_stream_stride;
  p.pcm_ch_stride = batch->pcm_channel_stride;
  p.num_samples = num_samples_dev;
  p.uniform_samples = batch->num_samples;
  p.num_streams = batch->num_streams;
  p.geo = geo;
  p.sampling_rate = batch->param.sampling_rate;
  p.trials = batch->param.num_encode_trials;
  p.aad = aad_dev;
  p.aad_stride = batch->aad_stream_stride;
  p.out_sizes = out_sizes_dev;
  p.block_begin = 0;
  p.block_end = aadf_num_blocks(batch->num_samples, geo.samples_per_block);
  (void)apply_segments(gpu, &p, p.block_end);
  CU((cudaError_t)aadk_launch_encode(&p, stream), "encode kernel launch");
  return AAD_APIRESULT_OK;
}

AADApiResult AADGpu_DecodeBatchDevice(struct AADGpu *gpu, const struct AADGpuBatch *batch, const uint8_t *aad_dev,
                                      const uint32_t *sizes_dev, int16_t *pcm_dev, void *stream)
{
  if (!gpu || !batch || !pcm_dev || !aad_dev) return AAD_APIRESULT_INVALID_ARGUMENT;
  struct aadf_geometry geo;
  const AADApiResult r = check_batch(batch, &geo);
  if (r != AAD_APIRESULT_OK) return r;
  if (batch->num_streams == 0) return AAD_APIRESULT_OK;
  CU(cudaSetDevice(gpu->device), "cudaSetDevice");
  struct aadk_decode_params p;
  memset(&p, 0, sizeof(p));
  p.aad = aad_dev;
  p.aad_stride = batch->aad_stream_stride;
  p.sizes = sizes_dev;
  p.uniform_size = (uint32_t)batch->aad_stream_stride;
  p.num_streams = batch->num_streams;
  p.geo = geo;
  p.block_begin = 0;
  p.block_end = aadf_num_blocks(batch->num_samples, geo.samples_per_block);
  p.read_headers = 1;
  /* a stream's own header may claim more samples than the batch describes (corrupt or mismatched): the rows
   * of pcm_dev hold batch->num_samples, and nothing is written past that */
  p.uniform_samples = batch->num_samples;
  p.pcm = pcm_dev;
  p.pcm_clip_stride = batch->pcm_stream_stride;
  p.pcm_ch_stride = batch->pcm_channel_stride;
  CU((cudaError_t)aadk_launch_decode(&p, stream), "decode kernel launch");
  return AAD_APIRESULT_OK;
}

/* ---- sine table for the synthetic generator ---------------------------------------------- */

void AADGpu_SynthLut(int16_t lut[1024]);
void AADGpu_SynthLut(int16_t lut[1024])
{
  for (int k = 0; k < 1024; k++) lut[k] = (int16_t)lrint(32767.0 * sin(2.0 * 3.14159265358979323846 * k / 1024.0));
}

static AADApiResult AADGpu_SynthBatchDevice_unlocked(struct AADGpu *gpu, const struct AADGpuBatch *batch, uint32_t first_stream,
                                     int16_t *pcm_dev, void *stream)
{
  if (!gpu || !batch || !pcm_dev) return AAD_APIRESULT_INVALID_ARGUMENT;
  if (batch->param.num_channels == 0 || batch->param.sampling_rate == 0) return AAD_APIRESULT_INVALID_FORMAT;
  CU(cudaSetDevice(gpu->device), "cudaSetDevice");
  if (!gpu->lut_ready) {
    int16_t lut[1024];
    AADGpu_SynthLut(lut);
    if (!aadgpu_reserve(gpu, &gpu->lut, sizeof(lut))) return AAD_APIRESULT_NG;
    CU(cudaMemcpy(gpu->lut.ptr, lut, sizeof(lut), cudaMemcpyHostToDevice), "upload sine table");
    gpu->lut_ready = 1;
  }
  struct aadk_synth_params p;
  memset(&p, 0, sizeof(p));
  p.pcm = pcm_dev;
  p.pcm_clip_stride = batch->pcm_stream_stride;
  p.pcm_ch_stride = batch->pcm_channel_stride;
  p.num_streams = batch->num_streams;
  p.channels = batch->param.num_channels;
  p.num_samples = batch->num_samples;
  p.sampling_rate = batch->param.sampling_rate;
  p.first_stream = first_stream;
  p.lut = (const int16_t *)gpu->lut.ptr;
  CU((cudaError_t)aadk_launch_synth(&p, stream), "synth kernel launch");
  return AAD_APIRESULT_OK;
}

AADApiResult AADGpu_SynthBatchDevice(struct AADGpu *gpu, const struct AADGpuBatch *batch, uint32_t first_stream,
                                     int16_t *pcm_dev, void *stream)
{
  WITH_CONTEXT_LOCK(gpu, AADGpu_SynthBatchDevice_unlocked(gpu, batch, first_stream, pcm_dev, stream));
}

AADApiResult AADGpu_Deinterleave16Device(struct AADGpu *gpu, const int16_t *interleaved_dev, int16_t *planar_dev,
                                         uint64_t channel_stride, uint32_t channels, uint32_t num_samples,
                                         void *stream)
{
  if (!gpu || !interleaved_dev || !planar_dev) return AAD_APIRESULT_INVALID_ARGUMENT;
  CU(cudaSetDevice(gpu->device), "cudaSetDevice");
  CU((cudaError_t)aadk_launch_deinterleave16(interleaved_dev, planar_dev, channel_stride, channels, num_samples, stream),
     "deinterleave kernel launch");
  return AAD_APIRESULT_OK;
}

AADApiResult AADGpu_Interleave16Device(struct AADGpu *gpu, const int16_t *planar_dev, uint64_t channel_stride,
                                       int16_t *interleaved_dev, uint32_t channels, uint32_t num_samples, void *stream)
{
  if (!gpu || !interleaved_dev || !planar_dev) return AAD_APIRESULT_INVALID_ARGUMENT;
  CU(cudaSetDevice(gpu->device), "cudaSetDevice");
  CU((cudaError_t)aadk_launch_interleave16(planar_dev, channel_stride, interleaved_dev, channels, num_samples, stream),
     "interleave kernel launch");
  return AAD_APIRESULT_OK;
}

/* ---- host pipelines ----------------------------------------------------------------------- */

static uint64_t round_up64(uint64_t v, uint64_t m) { return (v + m - 1) / m * m; }


/* how many block-range slices to cut the copies into: ~32 MiB of PCM each, at most AADGPU_MAX_SLICES */
static uint32_t pick_slices(uint64_t pcm_bytes, uint32_t num_blocks)
{
  uint64_t s = pcm_bytes / ((uint64_t)32 << 20);
  const char *env = getenv("AAD_B200_SLICES");        /* measurement: force the slice count of the batch pipelines */
  if (env && atoi(env) > 0) s = (uint64_t)atoi(env);
  if (s < 1) s = 1;
  if (s > AADGPU_MAX_SLICES) s = AADGPU_MAX_SLICES;
  if (s > num_blocks) s = num_blocks ? num_blocks : 1;
  return (uint32_t)s;
}

/* First block of slice k of `slices` over `nblk` blocks (k = slices: nblk).  With `ramp` the first and the last two slices
 * are a quarter and a half of the others: the pipeline's fill (slice 0 up the link and through both kernels before
 * anything comes down) and drain (the last slice down the link after everything else is done) overlap with nothing, so
 * they should be short.  AAD_B200_SLICE_RAMP=0 / 1 overrides the default for measurement. */
static uint32_t slice_bound(uint32_t nblk, uint32_t slices, uint32_t k, int ramp)
{
  if (!ramp || slices < 16) return (uint32_t)((uint64_t)nblk * k / slices);
  const uint64_t total = 4ull * slices - 10ull;          /* in quarters of a full slice */
  uint64_t cum;
  if (k == 0) cum = 0;
  else if (k == 1) cum = 1;
  else if (k <= slices - 2) cum = 3 + 4ull * (k - 2);
  else if (k == slices - 1) cum = total - 1;
  else cum = total;
  uint64_t b = (uint64_t)nblk * cum / total;
  /* no empty slices while there are at least as many blocks as slices */
  if (nblk >= slices) {
    if (b < k) b = k;
    if (b > (uint64_t)nblk - (slices - k)) b = (uint64_t)nblk - (slices - k);
  }
  return (uint32_t)b;
}
static int slice_ramp_default(void)
{
  static int on = -1;
  if (on < 0) on = (int)env_number("AAD_B200_SLICE_RAMP", 0, 1, 1);   /* on: 301.1 -> 296.4 ms on the bench batch, profiles/r02_host_link.md */
  return on;
}

/* one stream (or a shard of one): pieces of ~16 MiB so that even an eighth of an hour-long file overlaps its
 * copies with its kernels, without turning a small shard into dozens of API calls (the calls of the threads of
 * one process queue up behind each other); at most AADGPU_MAX_SLICES (one event each), at most `units` */
static uint32_t pick_stream_slices(uint64_t bytes, uint64_t units)
{
  uint64_t s = bytes / ((uint64_t)16 << 20);
  if (s < 1) s = 1;
  if (s > AADGPU_MAX_SLICES) s = AADGPU_MAX_SLICES;
  if (s > units) s = units ? units : 1;
  return (uint32_t)s;
}

/* 2-D copy of `rows` rows of `width` bytes; collapses to one 1-D copy when both sides are dense */
static cudaError_t copy_rows(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows,
                             enum cudaMemcpyKind kind, cudaStream_t s)
{
  if (rows == 0 || width == 0) return cudaSuccess;
  if (dpitch == width && spitch == width) return cudaMemcpyAsync(dst, src, width * rows, kind, s);
  return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, kind, s);
}

/* rows of a batch's PCM: uniform pitch when streams are packed back to back */
static cudaError_t copy_pcm_slice(const struct AADGpuBatch *b, uint32_t C, int to_device, int16_t *dev, uint64_t dev_pitch,
                                  int16_t *host, uint32_t s0, uint32_t s1, cudaStream_t st)
{
  const size_t width = (size_t)(s1 - s0) * sizeof(int16_t);
  const enum cudaMemcpyKind kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
  if (b->pcm_stream_stride == b->pcm_channel_stride * C) {
    void *d = dev + s0;
    void *h = host + s0;
    return to_device ? copy_rows(d, dev_pitch * 2, h, b->pcm_channel_stride * 2, width, (size_t)b->num_streams * C, kind, st)
                     : copy_rows(h, b->pcm_channel_stride * 2, d, dev_pitch * 2, width, (size_t)b->num_streams * C, kind, st);
  }
  for (uint32_t i = 0; i < b->num_streams; i++) {
    void *d = dev + (uint64_t)i * C * dev_pitch + s0;
    void *h = host + (uint64_t)i * b->pcm_stream_stride + s0;
    const cudaError_t e = to_device ? copy_rows(d, dev_pitch * 2, h, b->pcm_channel_stride * 2, width, C, kind, st)
                                    : copy_rows(h, b->pcm_channel_stride * 2, d, dev_pitch * 2, width, C, kind, st);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

static AADApiResult AADGpu_EncodeBatch_unlocked(struct AADGpu *gpu, const struct AADGpuBatch *batch, const int16_t *pcm,
                                const uint32_t *num_samples, uint8_t *aad, uint32_t *out_sizes)
{
  if (!gpu || !batch || !pcm || !aad) return AAD_APIRESULT_INVALID_ARGUMENT;
  struct aadf_geometry geo;
  const AADApiResult r = check_batch(batch, &geo);
  if (r != AAD_APIRESULT_OK) return r;
  const uint32_t N = batch->num_streams, C = geo.channels, ns = batch->num_samples;
  if (N == 0) return AAD_APIRESULT_OK;
  if (num_samples)
    for (uint32_t i = 0; i < N; i++)
      if (num_samples[i] == 0 || num_samples[i] > ns) return AAD_APIRESULT_INVALID_FORMAT;
  CU(cudaSetDevice(gpu->device), "cudaSetDevice");

  const uint32_t spb = geo.samples_per_block, bs = geo.block_size;
  const uint32_t nblk = aadf_num_blocks(ns, spb);
  const uint64_t pitch = round_up64(ns, 64);                              /* samples, 128-byte rows */
  const uint64_t astride = round_up64(aadf_stream_bytes_bound(ns, bs, spb) + 1, 128);
  if (!aadgpu_reserve(gpu, &gpu->pcm, (size_t)N * C * pitch * 2)) return AAD_APIRESULT_NG;
  if (!aadgpu_reserve(gpu, &gpu->aad, (size_t)N * astride + 128)) return AAD_APIRESULT_NG;
  const uint32_t segs = gpu->segment_blocks ? (nblk + gpu->segment_blocks - 1) / gpu->segment_blocks : 1;
  const size_t state_bytes = (size_t)N * C * (segs ? segs : 1) * AADK_STATE_WORDS * 4;
  if (!aadgpu_reserve(gpu, &gpu->state, state_bytes)) return AAD_APIRESULT_NG;
  if (num_samples && !aadgpu_reserve(gpu, &gpu->lens, (size_t)N * 4)) return AAD_APIRESULT_NG;
  int16_t *d_pcm = (int16_t *)gpu->pcm.ptr;
  uint8_t *d_aad = (uint8_t *)gpu->aad.ptr + 1;   /* block 0 of every stream lands 32-byte aligned */

  CU(cudaMemsetAsync(gpu->state.ptr, 0, state_bytes, gpu->s_run), "memset state");
  CU(cudaMemsetAsync(gpu->aad.ptr, 0, (size_t)N * astride + 128, gpu->s_run), "memset aad");
  if (num_samples)
    CU(cudaMemcpyAsync(gpu->lens.ptr, num_samples, (size_t)N * 4, cudaMemcpyHostToDevice, gpu->s_run), "H2D lengths");

  struct aadk_encode_params p;
  memset(&p, 0, sizeof(p));
  p.pcm = d_pcm;
  p.pcm_clip_stride = (uint64_t)C * pitch;
  p.pcm_ch_stride = pitch;
  p.num_samples = num_samples ? (const uint32_t *)gpu->lens.ptr : NULL;
  p.uniform_samples = ns;
  p.num_streams = N;
  p.geo = geo;
  p.sampling_rate = batch->param.sampling_rate;
  p.trials = batch->param.num_encode_trials;
  p.aad = d_aad;
  p.aad_stride = astride;
  p.state_in = (const int32_t *)gpu->state.ptr;
  p.state_out = (int32_t *)gpu->state.ptr;
  (void)apply_segments(gpu, &p, nblk);

  const uint32_t slices = pick_slices((uint64_t)N * C * ns * 2, nblk);
  for (uint32_t k = 0; k < slices; k++) {
    const uint32_t b0 = (uint32_t)((uint64_t)nblk * k / slices), b1 = (uint32_t)((uint64_t)nblk * (k + 1) / slices);
    const uint32_t s0 = b0 * spb, s1 = (b1 * (uint64_t)spb < ns) ? b1 * spb : ns;
    CU(copy_pcm_slice(batch, C, 1, d_pcm, pitch, (int16_t *)pcm, s0, s1, gpu->s_in), "H2D pcm");
    CU(cudaEventRecord(gpu->ev_in[k], gpu->s_in), "event");
    CU(cudaStreamWaitEvent(gpu->s_run, gpu->ev_in[k], 0), "wait");
    p.block_begin = b0;
    p.block_end = b1;
    CU((cudaError_t)aadk_launch_encode(&p, gpu->s_run), "encode kernel launch");
    CU(cudaEventRecord(gpu->ev_run[k], gpu->s_run), "event");
    CU(cudaStreamWaitEvent(gpu->s_out, gpu->ev_run[k], 0), "wait");
    const size_t off = b0 ? AADF_FILE_HEADER_BYTES + (size_t)b0 * bs : 0;
    const size_t end = AADF_FILE_HEADER_BYTES + (size_t)b1 * bs;
    CU(copy_rows(aad + off, batch->aad_stream_stride, d_aad + off, astride, end - off, N, cudaMemcpyDeviceToHost,
                 gpu->s_out), "D2H aad");
  }
  CU(cudaStreamSynchronize(gpu->s_out), "sync");
  if (out_sizes)
    for (uint32_t i = 0; i < N; i++)
      out_sizes[i] = (uint32_t)aadf_stream_bytes(num_samples ? num_samples[i] : ns, C, geo.bits, bs, spb);
  return AAD_APIRESULT_OK;
}

AADApiResult AADGpu_EncodeBatch(struct AADGpu *gpu, const struct AADGpuBatch *batch, const int16_t *pcm,
                                const uint32_t *num_samples, uint8_t *aad, uint32_t *out_sizes)
{
  WITH_CONTEXT_LOCK(gpu, AADGpu_EncodeBatch_unlocked(gpu, batch, pcm, num_samples, aad, out_sizes));
}

static AADApiResult AADGpu_DecodeBatch_unlocked(struct AADGpu *gpu, const struct AADGpuBatch *batch, const uint8_t *aad,
                                const uint32_t *sizes, int16_t *pcm)
{
  if (!gpu || !batch || !pcm || !aad) return AAD_APIRESULT_INVALID_ARGUMENT;
  struct aadf_geometry geo;
  const AADApiResult r = check_batch(batch, &geo);
  if (r != AAD_APIRESULT_OK) return r;
  const uint32_t N = batch->num_streams, C = geo.channels, ns = batch->num_samples;
  if (N == 0) return AAD_APIRESULT_OK;
  /* every stream must carry the geometry the batch describes (one kernel configuration) */
  for (uint32_t i = 0; i < N; i++) {
    const uint8_t *h = aad + (uint64_t)i * batch->aad_stream_stride;
    const uint32_t sz = sizes ? sizes[i] : (uint32_t)batch->aad_stream_stride;
    if (sz < AADF_FILE_HEADER_BYTES) return AAD_APIRESULT_INSUFFICIENT_DATA;
    if (h[0] != 'A' || h[1] != 'A' || h[2] != 'D' || h[3] != 0) return AAD_APIRESULT_INVALID_FORMAT;
    if (aadf_get_be32(h + 4) != AAD_FORMAT_VERSION || aadf_get_be32(h + 8) != AAD_CODEC_VERSION ||
        aadf_get_be16(h + 12) != C || aadf_get_be16(h + 22) != geo.bits || aadf_get_be16(h + 24) != geo.block_size ||
        aadf_get_be32(h + 26) != geo.samples_per_block || h[30] != geo.ms)
      return AAD_APIRESULT_INVALID_FORMAT;
    const uint32_t n = aadf_get_be32(h + 14);
    if (n == 0 || aadf_get_be32(h + 18) == 0) return AAD_APIRESULT_INVALID_FORMAT;
    if (n > ns) return AAD_APIRESULT_INSUFFICIENT_BUFFER;
  }
  CU(cudaSetDevice(gpu->device), "cudaSetDevice");

  const uint32_t spb = geo.samples_per_block, bs = geo.block_size;
  const uint32_t nblk = aadf_num_blocks(ns, spb);
  const uint64_t pitch = round_up64(ns, 64);
  const uint64_t bound = aadf_stream_bytes_bound(ns, bs, spb);
  const uint64_t astride = round_up64(bound + 1, 128);
  if (!aadgpu_reserve(gpu, &gpu->pcm, (size_t)N * C * pitch * 2)) return AAD_APIRESULT_NG;
  if (!aadgpu_reserve(gpu, &gpu->aad, (size_t)N * astride + 128)) return AAD_APIRESULT_NG;
  if (sizes && !aadgpu_reserve(gpu, &gpu->sizes, (size_t)N * 4)) return AAD_APIRESULT_NG;
  int16_t *d_pcm = (int16_t *)gpu->pcm.ptr;
  uint8_t *d_aad = (uint8_t *)gpu->aad.ptr + 1;
  if (sizes) CU(cudaMemcpyAsync(gpu->sizes.ptr, sizes, (size_t)N * 4, cudaMemcpyHostToDevice, gpu->s_in), "H2D sizes");

  struct aadk_decode_params p;
  memset(&p, 0, sizeof(p));
  p.aad = d_aad;
  p.aad_stride = astride;
  p.sizes = sizes ? (const uint32_t *)gpu->sizes.ptr : NULL;
  p.uniform_size = (uint32_t)(batch->aad_stream_stride < bound ? batch->aad_stream_stride : bound);
  p.num_streams = N;
  p.geo = geo;
  p.read_headers = 1;
  p.uniform_samples = ns;
  p.pcm = d_pcm;
  p.pcm_clip_stride = (uint64_t)C * pitch;
  p.pcm_ch_stride = pitch;

  const uint32_t slices = pick_slices((uint64_t)N * C * ns * 2, nblk);
  for (uint32_t k = 0; k < slices; k++) {
    const uint32_t b0 = (uint32_t)((uint64_t)nblk * k / slices), b1 = (uint32_t)((uint64_t)nblk * (k + 1) / slices);
    const uint32_t s0 = b0 * spb, s1 = (b1 * (uint64_t)spb < ns) ? b1 * spb : ns;
    const size_t off = b0 ? AADF_FILE_HEADER_BYTES + (size_t)b0 * bs : 0;
    size_t end = AADF_FILE_HEADER_BYTES + (size_t)b1 * bs;
    if (end > batch->aad_stream_stride) end = batch->aad_stream_stride;
    CU(copy_rows(d_aad + off, astride, aad + off, batch->aad_stream_stride, end - off, N, cudaMemcpyHostToDevice,
                 gpu->s_in), "H2D aad");
    CU(cudaEventRecord(gpu->ev_in[k], gpu->s_in), "event");
    CU(cudaStreamWaitEvent(gpu->s_run, gpu->ev_in[k], 0), "wait");
    p.block_begin = b0;
    p.block_end = b1;
    CU((cudaError_t)aadk_launch_decode(&p, gpu->s_run), "decode kernel launch");
    CU(cudaEventRecord(gpu->ev_run[k], gpu->s_run), "event");
    CU(cudaStreamWaitEvent(gpu->s_out, gpu->ev_run[k], 0), "wait");
    CU(copy_pcm_slice(batch, C, 0, d_pcm, pitch, pcm, s0, s1, gpu->s_out), "D2H pcm");
  }
  CU(cudaStreamSynchronize(gpu->s_out), "sync");
  return AAD_APIRESULT_OK;
}

AADApiResult AADGpu_DecodeBatch(struct AADGpu *gpu, const struct AADGpuBatch *batch, const uint8_t *aad,
                                const uint32_t *sizes, int16_t *pcm)
{
  WITH_CONTEXT_LOCK(gpu, AADGpu_DecodeBatch_unlocked(gpu, batch, aad, sizes, pcm));
}

/* Encode a batch and decode it back in one pass over the data (what src/main.c:275-346 does for one
 * file, execute_reconstruction_core): per block-range slice  H2D pcm | encode | decode | D2H .aad + D2H pcm.
 * The encoded streams never make the round trip over PCIe, and the two directions of the link are busy
 * at the same time: ~25 GB per 12,500 ten-second clips instead of 27.7 GB half-duplex.  aad / out_sizes
 * may be NULL when only the reconstruction is wanted. */
static AADApiResult AADGpu_ReconstructBatch_unlocked(struct AADGpu *gpu, const struct AADGpuBatch *batch, const int16_t *pcm,
                                     const uint32_t *num_samples, uint8_t *aad, uint32_t *out_sizes,
                                     int16_t *reconstructed, int copies_only)
{
  if (!gpu || !batch || !pcm || !reconstructed) return AAD_APIRESULT_INVALID_ARGUMENT;
  struct aadf_geometry geo;
  const AADApiResult r = check_batch(batch, &geo);
  if (r != AAD_APIRESULT_OK) return r;
  const uint32_t N = batch->num_streams, C = geo.channels, ns = batch->num_samples;
  if (N == 0) return AAD_APIRESULT_OK;
  if (num_samples)
    for (uint32_t i = 0; i < N; i++)
      if (num_samples[i] == 0 || num_samples[i] > ns) return AAD_APIRESULT_INVALID_FORMAT;
  CU(cudaSetDevice(gpu->device), "cudaSetDevice");

  const uint32_t spb = geo.samples_per_block, bs = geo.block_size;
  const uint32_t nblk = aadf_num_blocks(ns, spb);
  const uint64_t pitch = round_up64(ns, 64);
  const uint64_t astride = round_up64(aadf_stream_bytes_bound(ns, bs, spb) + 1, 128);
  if (!aadgpu_reserve(gpu, &gpu->pcm, (size_t)N * C * pitch * 2)) return AAD_APIRESULT_NG;
  if (!aadgpu_reserve(gpu, &gpu->pcm2, (size_t)N * C * pitch * 2)) return AAD_APIRESULT_NG;
  if (!aadgpu_reserve(gpu, &gpu->aad, (size_t)N * astride + 128)) return AAD_APIRESULT_NG;
  const uint32_t segs = gpu->segment_blocks ? (nblk + gpu->segment_blocks - 1) / gpu->segment_blocks : 1;
  const size_t state_bytes = (size_t)N * C * (segs ? segs : 1) * AADK_STATE_WORDS * 4;
  if (!aadgpu_reserve(gpu, &gpu->state, state_bytes)) return AAD_APIRESULT_NG;
  if (!aadgpu_reserve(gpu, &gpu->lens, (size_t)N * 4)) return AAD_APIRESULT_NG;
  if (!aadgpu_reserve(gpu, &gpu->sizes, (size_t)N * 4)) return AAD_APIRESULT_NG;
  int16_t *d_pcm = (int16_t *)gpu->pcm.ptr, *d_out = (int16_t *)gpu->pcm2.ptr;
  uint8_t *d_aad = (uint8_t *)gpu->aad.ptr + 1;

  if (!copies_only) {
    CU(cudaMemsetAsync(gpu->state.ptr, 0, state_bytes, gpu->s_run), "memset state");
    CU(cudaMemsetAsync(gpu->aad.ptr, 0, (size_t)N * astride + 128, gpu->s_run), "memset aad");
    if (num_samples)
      CU(cudaMemcpyAsync(gpu->lens.ptr, num_samples, (size_t)N * 4, cudaMemcpyHostToDevice, gpu->s_run), "H2D lengths");
  }

  struct aadk_encode_params e;
  memset(&e, 0, sizeof(e));
  e.pcm = d_pcm;
  e.pcm_clip_stride = (uint64_t)C * pitch;
  e.pcm_ch_stride = pitch;
  e.num_samples = num_samples ? (const uint32_t *)gpu->lens.ptr : NULL;
  e.uniform_samples = ns;
  e.num_streams = N;
  e.geo = geo;
  e.sampling_rate = batch->param.sampling_rate;
  e.trials = batch->param.num_encode_trials;
  e.aad = d_aad;
  e.aad_stride = astride;
  e.out_sizes = (uint32_t *)gpu->sizes.ptr;          /* written with the first slice: the decoder's byte bounds */
  e.state_in = (const int32_t *)gpu->state.ptr;
  e.state_out = (int32_t *)gpu->state.ptr;
  (void)apply_segments(gpu, &e, nblk);

  struct aadk_decode_params d;
  memset(&d, 0, sizeof(d));
  d.aad = d_aad;
  d.aad_stride = astride;
  d.sizes = (const uint32_t *)gpu->sizes.ptr;
  d.num_streams = N;
  d.geo = geo;
  d.read_headers = 1;                                 /* each stream's own length, written by the encoder */
  d.uniform_samples = ns;                             /* ... and never more than a row holds */
  d.pcm = d_out;
  d.pcm_clip_stride = (uint64_t)C * pitch;
  d.pcm_ch_stride = pitch;

  const uint32_t slices = pick_slices((uint64_t)N * C * ns * 2, nblk);
  /* the .aad rows travel on a device -> host queue of their own, beside the PCM rows (the bench batch's round trip:
   * 307 -> 296 ms; AAD_B200_D2H_QUEUES=1 puts them back behind the PCM rows, for measurement) */
  const char *q2 = getenv("AAD_B200_D2H_QUEUES");
  cudaStream_t s_aad = (q2 && atoi(q2) == 1) ? gpu->s_out : gpu->s_out2;
  const int ramp = slice_ramp_default();
  for (uint32_t k = 0; k < slices; k++) {
    const uint32_t b0 = slice_bound(nblk, slices, k, ramp), b1 = slice_bound(nblk, slices, k + 1, ramp);
    const uint32_t s0 = b0 * spb, s1 = (b1 * (uint64_t)spb < ns) ? b1 * spb : ns;
    CU(copy_pcm_slice(batch, C, 1, d_pcm, pitch, (int16_t *)pcm, s0, s1, gpu->s_in), "H2D pcm");
    if (!copies_only) {   /* copies_only: the same copies with nothing between them (AADGpu_CopyProbeBatch) */
      CU(cudaEventRecord(gpu->ev_in[k], gpu->s_in), "event");
      CU(cudaStreamWaitEvent(gpu->s_run, gpu->ev_in[k], 0), "wait");
      e.block_begin = d.block_begin = b0;
      e.block_end = d.block_end = b1;
      CU((cudaError_t)aadk_launch_encode(&e, gpu->s_run), "encode kernel launch");
      CU((cudaError_t)aadk_launch_decode(&d, gpu->s_run), "decode kernel launch");
      CU(cudaEventRecord(gpu->ev_run[k], gpu->s_run), "event");
      CU(cudaStreamWaitEvent(gpu->s_out, gpu->ev_run[k], 0), "wait");
      if (s_aad != gpu->s_out) CU(cudaStreamWaitEvent(s_aad, gpu->ev_run[k], 0), "wait");
    }
    if (aad) {
      const size_t off = b0 ? AADF_FILE_HEADER_BYTES + (size_t)b0 * bs : 0;
      const size_t end = AADF_FILE_HEADER_BYTES + (size_t)b1 * bs;
      CU(copy_rows(aad + off, batch->aad_stream_stride, d_aad + off, astride, end - off, N, cudaMemcpyDeviceToHost,
                   s_aad), "D2H aad");
    }
    CU(copy_pcm_slice(batch, C, 0, d_out, pitch, reconstructed, s0, s1, gpu->s_out), "D2H pcm");
  }
  if (copies_only) CU(cudaStreamSynchronize(gpu->s_in), "sync");
  if (s_aad != gpu->s_out) CU(cudaStreamSynchronize(s_aad), "sync");
  CU(cudaStreamSynchronize(gpu->s_out), "sync");
  if (out_sizes)
    for (uint32_t i = 0; i < N; i++)
      out_sizes[i] = (uint32_t)aadf_stream_bytes(num_samples ? num_samples[i] : ns, C, geo.bits, bs, spb);
  return AAD_APIRESULT_OK;
}

AADApiResult AADGpu_ReconstructBatch(struct AADGpu *gpu, const struct AADGpuBatch *batch, const int16_t *pcm,
                                     const uint32_t *num_samples, uint8_t *aad, uint32_t *out_sizes,
                                     int16_t *reconstructed)
{
  WITH_CONTEXT_LOCK(gpu, AADGpu_ReconstructBatch_unlocked(gpu, batch, pcm, num_samples, aad, out_sizes, reconstructed, 0));
}

/* Exactly the host <-> device copies AADGpu_ReconstructBatch issues for this batch and these (pinned) buffers -- same
 * slices, same row shapes, same two streams -- with no kernel and no dependency between them: how long the copies
 * alone take, i.e. what the end-to-end call could at best reach.  Right after a AADGpu_ReconstructBatch of the same
 * batch the device buffers still hold its results, so the host buffers receive the same bytes again. */
AADApiResult AADGpu_CopyProbeBatch(struct AADGpu *gpu, const struct AADGpuBatch *batch, const int16_t *pcm, uint8_t *aad,
                                   int16_t *reconstructed)
{
  WITH_CONTEXT_LOCK(gpu, AADGpu_ReconstructBatch_unlocked(gpu, batch, pcm, NULL, aad, NULL, reconstructed, 1));
}

/* ---- single-stream paths behind the drop-in API ------------------------------------------- */
/*
 * AADEncoder_EncodeWhole / AADDecoder_DecodeWhole hand over what src/main.c:94-103,169-172 allocates: plain malloc'd
 * memory, one int32 per 16-bit sample.  Copying that as it is would move 4 bytes per sample through the driver's own
 * staging, one synchronous piece at a time.  Instead the stream is cut into slices that travel through a small ring of
 * pinned buffers as int16 (2 bytes per sample on the link), and the conversion int32 <-> int16 is done by a few host
 * threads while they move the samples between the caller's rows and the ring -- the only place where every sample has
 * to be touched by the host anyway.  While slice k is converted, slice k+1 is on the link and slice k+2 in the kernel.
 */
#define AADGPU_RING_SLOTS 3
#define AADGPU_RING_BYTES ((size_t)16 << 20)      /* per slot and direction (default; AAD_B200_RING_MIB) */
#define AADGPU_MAX_HOST_THREADS 32
#define AADGPU_DEFAULT_HOST_THREADS 16

/* measurement knobs, read once: AAD_B200_HOST_THREADS (conversion threads, caller included; default: the CPUs, at most
 * AADGPU_DEFAULT_HOST_THREADS) and AAD_B200_RING_MIB (size of a ring slot = of a slice; default 16, at most 64)
 * -- tools/dropin_sweep.py */
static long env_number(const char *name, long lo, long hi, long fallback)
{
  const char *v = getenv(name);
  if (!v || !*v) return fallback;
  char *end;
  const long x = strtol(v, &end, 10);
  return (*end == 0 && x >= lo && x <= hi) ? x : fallback;
}
static size_t ring_slice_bytes(void)
{
  static size_t bytes = 0;     /* benign race: every thread computes the same value */
  if (!bytes) bytes = (size_t)env_number("AAD_B200_RING_MIB", 1, 64, (long)(AADGPU_RING_BYTES >> 20)) << 20;
  return bytes;
}

/* a minimal fork-join pool: run fn(arg, i) for i in [0, n) on the calling thread and its helpers (AAD_B200_HOST_THREADS - 1
 * of them, 15 by default where the machine has the CPUs) */
struct host_pool {
  pthread_mutex_t lock;
  pthread_cond_t wake, done;
  pthread_t threads[AADGPU_MAX_HOST_THREADS];
  int num_threads, started;
  void (*fn)(void *, uint32_t);
  void *arg;
  uint32_t next, total, finished;
  uint64_t generation;
};
static struct host_pool g_pool = { PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER,
                                   {0}, 0, 0, NULL, NULL, 0, 0, 0, 0 };
static pthread_mutex_t g_pool_user = PTHREAD_MUTEX_INITIALIZER;   /* one parallel_for at a time */

static void *host_pool_worker(void *unused)
{
  (void)unused;
  uint64_t seen = 0;
  pthread_mutex_lock(&g_pool.lock);
  for (;;) {
    while (g_pool.generation == seen || g_pool.next >= g_pool.total) {
      seen = g_pool.generation;
      pthread_cond_wait(&g_pool.wake, &g_pool.lock);
    }
    while (g_pool.next < g_pool.total) {
      const uint32_t i = g_pool.next++;
      void (*fn)(void *, uint32_t) = g_pool.fn;
      void *arg = g_pool.arg;
      pthread_mutex_unlock(&g_pool.lock);
      fn(arg, i);
      pthread_mutex_lock(&g_pool.lock);
      if (++g_pool.finished == g_pool.total) pthread_cond_signal(&g_pool.done);
    }
    seen = g_pool.generation;
  }
  return NULL;
}

static void host_parallel_for(uint32_t n, void (*fn)(void *, uint32_t), void *arg)
{
  if (n == 0) return;
  pthread_mutex_lock(&g_pool_user);
  pthread_mutex_lock(&g_pool.lock);
  if (!g_pool.started) {
    long cpus = sysconf(_SC_NPROCESSORS_ONLN);
    if (cpus > AADGPU_DEFAULT_HOST_THREADS) cpus = AADGPU_DEFAULT_HOST_THREADS;
    cpus = env_number("AAD_B200_HOST_THREADS", 1, AADGPU_MAX_HOST_THREADS, cpus);
    int want = (int)(cpus > 1 ? cpus - 1 : 0);
    if (want > AADGPU_MAX_HOST_THREADS - 1) want = AADGPU_MAX_HOST_THREADS - 1;
    for (int t = 0; t < want; t++)
      if (pthread_create(&g_pool.threads[g_pool.num_threads], NULL, host_pool_worker, NULL) == 0) {
        pthread_detach(g_pool.threads[g_pool.num_threads]);
        g_pool.num_threads++;
      }
    g_pool.started = 1;
  }
  g_pool.fn = fn;
  g_pool.arg = arg;
  g_pool.next = 0;
  g_pool.total = n;
  g_pool.finished = 0;
  g_pool.generation++;
  pthread_cond_broadcast(&g_pool.wake);
  while (g_pool.next < g_pool.total) {          /* the caller works too (and alone when no helper could be started) */
    const uint32_t i = g_pool.next++;
    pthread_mutex_unlock(&g_pool.lock);
    fn(arg, i);
    pthread_mutex_lock(&g_pool.lock);
    g_pool.finished++;
  }
  while (g_pool.finished < g_pool.total) pthread_cond_wait(&g_pool.done, &g_pool.lock);
  g_pool.total = 0;
  pthread_mutex_unlock(&g_pool.lock);
  pthread_mutex_unlock(&g_pool_user);
}

/* one conversion job: `rows` rows, samples [0, n) each, cut into pieces of kConvPiece samples */
/* 2^16 samples = 128 KiB of int16 per piece: a 16 MiB slice is 128 pieces, every thread of the pool gets several
 * (2^18, round 2's first choice, left threads idle at the end of every slice; AAD_B200_CONV_PIECE_LOG2 for the A/B) */
static uint64_t conv_piece_samples(void)
{
  static uint64_t n = 0;
  if (!n) n = (uint64_t)1 << env_number("AAD_B200_CONV_PIECE_LOG2", 10, 24, 16);
  return n;
}
#define kConvPiece conv_piece_samples()
struct conv_job {
  int widen;                       /* 1: ring int16 -> caller int32, 0: caller int32 -> ring int16 */
  uint32_t rows, pieces_per_row;
  uint64_t n;                      /* samples per row in this slice */
  int16_t *ring;                   /* [rows][n] */
  int32_t *const *wide;            /* caller rows */
  uint64_t first;                  /* first sample of the slice in the caller's rows */
};

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
/* 16 samples per turn: two sign-extending 256-bit loads' worth, two streaming 32-byte stores (a quarter of the SSE2
 * loop's instructions; chosen at run time where the CPU has AVX2).  wide + t is 32-byte aligned on entry. */
__attribute__((target("avx2"))) static uint64_t widen_avx2(const int16_t *narrow, int32_t *wide, uint64_t t, uint64_t b)
{
  for (; t + 16 <= b; t += 16) {
    const __m256i lo = _mm256_cvtepi16_epi32(_mm_loadu_si128((const __m128i *)(narrow + t)));
    const __m256i hi = _mm256_cvtepi16_epi32(_mm_loadu_si128((const __m128i *)(narrow + t + 8)));
    _mm256_stream_si256((__m256i *)(wide + t), lo);
    _mm256_stream_si256((__m256i *)(wide + t + 8), hi);
  }
  return t;
}
static int have_avx2(void)
{
  static int have = -1;
  if (have < 0) {
    const char *off = getenv("AAD_B200_NO_AVX2");     /* measurement: the SSE2 loop */
    have = (__builtin_cpu_supports("avx2") && !(off && atoi(off) > 0)) ? 1 : 0;
  }
  return have;
}
#define AADGPU_HAVE_AVX2_PATH 1
#endif

static void conv_piece(void *arg, uint32_t i)
{
  const struct conv_job *j = (const struct conv_job *)arg;
  const uint32_t r = i / j->pieces_per_row, k = i % j->pieces_per_row;
  const uint64_t a = (uint64_t)k * kConvPiece, b = (a + kConvPiece < j->n) ? a + kConvPiece : j->n;
  int16_t *narrow = j->ring + (uint64_t)r * j->n;
  int32_t *wide = j->wide[r] + j->first;
  uint64_t t = a;
  if (j->widen) {
#if defined(__SSE2__)
    /* the caller's rows are written once and not read back here: streaming stores skip the read-for-ownership,
     * a third of the memory traffic of this loop */
#ifdef AADGPU_HAVE_AVX2_PATH
    if (have_avx2()) {
      for (; t < b && (((uintptr_t)(wide + t)) & 31u); t++) wide[t] = narrow[t];
      t = widen_avx2(narrow, wide, t, b);
    }
#endif
    for (; t < b && (((uintptr_t)(wide + t)) & 15u); t++) wide[t] = narrow[t];
    const __m128i zero = _mm_setzero_si128();
    for (; t + 8 <= b; t += 8) {
      const __m128i v = _mm_loadu_si128((const __m128i *)(narrow + t));
      _mm_stream_si128((__m128i *)(wide + t), _mm_srai_epi32(_mm_unpacklo_epi16(zero, v), 16));
      _mm_stream_si128((__m128i *)(wide + t + 4), _mm_srai_epi32(_mm_unpackhi_epi16(zero, v), 16));
    }
    _mm_sfence();
#endif
    for (; t < b; t++) wide[t] = narrow[t];
  } else {
    for (; t < b; t++) narrow[t] = (int16_t)wide[t];   /* int16-range values: src/aad_encoder.c:451,612 */
  }
}

static void convert_rows(int widen, int16_t *ring, int32_t *const *wide, uint32_t rows, uint64_t first, uint64_t n)
{
  struct conv_job j;
  j.widen = widen;
  j.rows = rows;
  j.n = n;
  j.pieces_per_row = (uint32_t)((n + kConvPiece - 1) / kConvPiece);
  j.ring = ring;
  j.wide = wide;
  j.first = first;
  host_parallel_for(rows * j.pieces_per_row, conv_piece, &j);
}

struct copy_job { uint8_t *dst; const uint8_t *src; size_t bytes; };
/* the .aad bytes of a slice are a quarter of its samples: 128 KiB pieces, or only a few threads would copy (1 MiB pieces
 * = 4 threads per 16 MiB slice; AAD_B200_COPY_PIECE_LOG2 for the A/B) */
static size_t copy_piece_bytes(void)
{
  static size_t n = 0;
  if (!n) n = (size_t)1 << env_number("AAD_B200_COPY_PIECE_LOG2", 10, 24, 17);
  return n;
}
#define kCopyPiece copy_piece_bytes()
static void copy_piece(void *arg, uint32_t i)
{
  const struct copy_job *j = (const struct copy_job *)arg;
  const size_t a = (size_t)i * kCopyPiece, b = (a + kCopyPiece < j->bytes) ? a + kCopyPiece : j->bytes;
  memcpy(j->dst + a, j->src + a, b - a);
}
static void copy_bytes(uint8_t *dst, const uint8_t *src, size_t bytes)
{
  struct copy_job j = { dst, src, bytes };
  host_parallel_for((uint32_t)((bytes + kCopyPiece - 1) / kCopyPiece), copy_piece, &j);
}

static int ring_ready(struct AADGpu *gpu)
{
  for (int i = 0; i < AADGPU_RING_SLOTS; i++) {
    if (!gpu->ring_in[i] && cudaMallocHost(&gpu->ring_in[i], ring_slice_bytes()) != cudaSuccess) goto fail;
    if (!gpu->ring_out[i] && cudaMallocHost(&gpu->ring_out[i], ring_slice_bytes()) != cudaSuccess) goto fail;
  }
  return 1;
fail:
  aadgpu_fail("cudaMallocHost (bounce ring)", cudaGetLastError());
  return 0;
}

/* blocks per slice so that a slice's int16 samples (all channels) and its encoded bytes both fit a ring slot */
static uint32_t ring_slice_blocks(const struct aadf_geometry *geo)
{
  const uint64_t per_block = (uint64_t)geo->channels * geo->samples_per_block * 2;
  uint64_t n = ring_slice_bytes() / (per_block > geo->block_size ? per_block : geo->block_size);
  return (uint32_t)(n ? n : 1);
}

static AADApiResult aadgpu_encode_stream_i32_unlocked(struct AADGpu *gpu, const struct aadf_geometry *geo, uint32_t sampling_rate,
                                      uint32_t trials, const int32_t *const *input, uint32_t num_samples,
                                      int32_t *state, uint8_t *data, uint32_t *output_size)
{
  const uint32_t C = geo->channels, spb = geo->samples_per_block, bs = geo->block_size;
  CU(cudaSetDevice(gpu->device), "cudaSetDevice");
  const uint64_t pitch = round_up64(num_samples, 64);
  const uint64_t bytes = aadf_stream_bytes(num_samples, C, geo->bits, bs, spb);
  const uint64_t bound = aadf_stream_bytes_bound(num_samples, bs, spb);
  const uint32_t nblk = aadf_num_blocks(num_samples, spb);
  if (!aadgpu_reserve(gpu, &gpu->pcm, (size_t)C * pitch * 2)) return AAD_APIRESULT_NG;    /* int16 rows for the kernel */
  if (!aadgpu_reserve(gpu, &gpu->aad, (size_t)bound + 128)) return AAD_APIRESULT_NG;
  if (!aadgpu_reserve(gpu, &gpu->state, (size_t)C * AADK_STATE_WORDS * 4)) return AAD_APIRESULT_NG;
  if (!ring_ready(gpu)) return AAD_APIRESULT_NG;
  int16_t *d_pcm = (int16_t *)gpu->pcm.ptr;
  uint8_t *d_aad = (uint8_t *)gpu->aad.ptr + 1;      /* block 0 lands 32-byte aligned */
  CU(cudaMemsetAsync(gpu->aad.ptr, 0, (size_t)bound + 128, gpu->s_run), "memset aad");
  CU(cudaMemcpyAsync(gpu->state.ptr, state, (size_t)C * AADK_STATE_WORDS * 4, cudaMemcpyHostToDevice, gpu->s_run), "H2D state");
  struct aadk_encode_params p;
  memset(&p, 0, sizeof(p));
  p.pcm = d_pcm;
  p.pcm_clip_stride = (uint64_t)C * pitch;
  p.pcm_ch_stride = pitch;
  p.uniform_samples = num_samples;
  p.num_streams = 1;
  p.geo = *geo;
  p.sampling_rate = sampling_rate;
  p.trials = trials;
  p.aad = d_aad;
  p.aad_stride = bound;
  p.state_in = (const int32_t *)gpu->state.ptr;
  p.state_out = (int32_t *)gpu->state.ptr;

  /* slice k: narrow the caller's int32 rows into ring slot k % R | H2D | encode its blocks (the chain state crosses
   * the launches in the device state array, as in the batch pipelines) | D2H its bytes | copy them to the caller */
  const uint32_t per = ring_slice_blocks(geo);
  const uint32_t slices = (nblk + per - 1) / per;
  const uint32_t lag = AADGPU_RING_SLOTS - 1;
  for (uint32_t k = 0; k < slices + lag; k++) {
    if (k < slices) {
      const uint32_t slot = k % AADGPU_RING_SLOTS;
      const uint32_t b0 = k * per, b1 = (b0 + per < nblk) ? b0 + per : nblk;
      const uint64_t s0 = (uint64_t)b0 * spb, s1 = ((uint64_t)b1 * spb < num_samples) ? (uint64_t)b1 * spb : num_samples;
      if (k >= AADGPU_RING_SLOTS) CU(cudaEventSynchronize(gpu->ev_in[slot]), "ring wait");   /* slot's last H2D has left it */
      convert_rows(0, (int16_t *)gpu->ring_in[slot], (int32_t *const *)input, C, s0, s1 - s0);
      for (uint32_t c = 0; c < C; c++)
        CU(cudaMemcpyAsync(d_pcm + c * pitch + s0, (int16_t *)gpu->ring_in[slot] + (uint64_t)c * (s1 - s0), (size_t)(s1 - s0) * 2,
                           cudaMemcpyHostToDevice, gpu->s_in), "H2D pcm");
      CU(cudaEventRecord(gpu->ev_in[slot], gpu->s_in), "event");
      CU(cudaStreamWaitEvent(gpu->s_run, gpu->ev_in[slot], 0), "wait");
      p.block_begin = b0;
      p.block_end = b1;
      CU((cudaError_t)aadk_launch_encode(&p, gpu->s_run), "encode kernel launch");
      CU(cudaEventRecord(gpu->ev_run[slot], gpu->s_run), "event");
      CU(cudaStreamWaitEvent(gpu->s_out, gpu->ev_run[slot], 0), "wait");
      const size_t off = b0 ? AADF_FILE_HEADER_BYTES + (size_t)b0 * bs : 0;
      size_t end = AADF_FILE_HEADER_BYTES + (size_t)b1 * bs;
      if (end > bytes) end = (size_t)bytes;
      CU(cudaMemcpyAsync(gpu->ring_out[slot], d_aad + off, end - off, cudaMemcpyDeviceToHost, gpu->s_out), "D2H aad");
      CU(cudaEventRecord(gpu->ev_out[slot], gpu->s_out), "event");
    }
    if (k >= lag) {
      const uint32_t j = k - lag, slot = j % AADGPU_RING_SLOTS;
      const uint32_t b0 = j * per, b1 = (b0 + per < nblk) ? b0 + per : nblk;
      const size_t off = b0 ? AADF_FILE_HEADER_BYTES + (size_t)b0 * bs : 0;
      size_t end = AADF_FILE_HEADER_BYTES + (size_t)b1 * bs;
      if (end > bytes) end = (size_t)bytes;
      CU(cudaEventSynchronize(gpu->ev_out[slot]), "ring wait");
      copy_bytes(data + off, (const uint8_t *)gpu->ring_out[slot], end - off);
    }
  }
  CU(cudaMemcpyAsync(state, gpu->state.ptr, (size_t)C * AADK_STATE_WORDS * 4, cudaMemcpyDeviceToHost, gpu->s_run), "D2H state");
  CU(cudaStreamSynchronize(gpu->s_run), "sync");
  *output_size = (uint32_t)bytes;
  return AAD_APIRESULT_OK;
}

AADApiResult aadgpu_encode_stream_i32(struct AADGpu *gpu, const struct aadf_geometry *geo, uint32_t sampling_rate,
                                      uint32_t trials, const int32_t *const *input, uint32_t num_samples,
                                      int32_t *state, uint8_t *data, uint32_t *output_size)
{
  WITH_CONTEXT_LOCK(gpu, aadgpu_encode_stream_i32_unlocked(gpu, geo, sampling_rate, trials, input, num_samples, state, data, output_size));
}

static AADApiResult aadgpu_decode_stream_i32_unlocked(struct AADGpu *gpu, const struct aadf_geometry *geo, const uint8_t *data,
                                      uint32_t data_size, uint32_t num_blocks, uint32_t num_samples,
                                      uint32_t buf_samples, int32_t *const *buffer)
{
  const uint32_t C = geo->channels, spb = geo->samples_per_block, bs = geo->block_size;
  if (num_blocks == 0) return AAD_APIRESULT_OK;
  CU(cudaSetDevice(gpu->device), "cudaSetDevice");
  /* samples the block loop writes: every block min(spb, room) -- src/aad_decoder.c:356,524 */
  uint64_t total = (uint64_t)num_blocks * spb;
  if (total > buf_samples) total = buf_samples;
  const uint64_t pitch = round_up64(total, 64);
  uint64_t span = AADF_FILE_HEADER_BYTES + (uint64_t)num_blocks * bs;
  if (span > data_size) span = data_size;
  if (!aadgpu_reserve(gpu, &gpu->pcm, (size_t)C * pitch * 2)) return AAD_APIRESULT_NG;    /* int16 rows from the kernel */
  if (!aadgpu_reserve(gpu, &gpu->aad, (size_t)span + 128)) return AAD_APIRESULT_NG;
  if (!ring_ready(gpu)) return AAD_APIRESULT_NG;
  int16_t *d_pcm = (int16_t *)gpu->pcm.ptr;
  uint8_t *d_aad = (uint8_t *)gpu->aad.ptr + 1;

  struct aadk_decode_params p;
  memset(&p, 0, sizeof(p));
  p.aad = d_aad;
  p.aad_stride = span;
  p.uniform_size = (uint32_t)span;
  p.num_streams = 1;
  p.geo = *geo;
  p.uniform_samples = num_samples;
  p.buf_samples = buf_samples;
  p.pcm = d_pcm;
  p.pcm_clip_stride = (uint64_t)C * pitch;
  p.pcm_ch_stride = pitch;

  /* slice k: the caller's bytes into ring slot k % R | H2D | decode its blocks | D2H its int16 rows into the ring |
   * widen them into the caller's int32 rows (host threads) */
  const uint32_t per = ring_slice_blocks(geo);
  const uint32_t slices = (num_blocks + per - 1) / per;
  const uint32_t lag = AADGPU_RING_SLOTS - 1;
  /* AAD_B200_TRACE: where the host thread's time goes (ring waits | staging copy | enqueue | result wait | widening) */
  const int trace = trace_on();
  double t_phase[5] = {0, 0, 0, 0, 0}, t_mark = trace ? wall_seconds() : 0.0;
#define PHASE(i) do { if (trace) { const double now_ = wall_seconds(); t_phase[i] += now_ - t_mark; t_mark = now_; } } while (0)
  for (uint32_t k = 0; k < slices + lag; k++) {
    if (k < slices) {
      const uint32_t slot = k % AADGPU_RING_SLOTS;
      const uint32_t b0 = k * per, b1 = (b0 + per < num_blocks) ? b0 + per : num_blocks;
      const uint64_t s0 = (uint64_t)b0 * spb, s1 = ((uint64_t)b1 * spb < total) ? (uint64_t)b1 * spb : total;
      const size_t off = b0 ? AADF_FILE_HEADER_BYTES + (size_t)b0 * bs : 0;
      size_t end = AADF_FILE_HEADER_BYTES + (size_t)b1 * bs;
      if (end > span) end = (size_t)span;
      if (k >= AADGPU_RING_SLOTS) CU(cudaEventSynchronize(gpu->ev_in[slot]), "ring wait");
      PHASE(0);
      if (end > off) {
        copy_bytes((uint8_t *)gpu->ring_in[slot], data + off, end - off);
        PHASE(1);
        CU(cudaMemcpyAsync(d_aad + off, gpu->ring_in[slot], end - off, cudaMemcpyHostToDevice, gpu->s_in), "H2D aad");
      }
      CU(cudaEventRecord(gpu->ev_in[slot], gpu->s_in), "event");
      CU(cudaStreamWaitEvent(gpu->s_run, gpu->ev_in[slot], 0), "wait");
      p.block_begin = b0;
      p.block_end = b1;
      CU((cudaError_t)aadk_launch_decode(&p, gpu->s_run), "decode kernel launch");
      CU(cudaEventRecord(gpu->ev_run[slot], gpu->s_run), "event");
      CU(cudaStreamWaitEvent(gpu->s_out, gpu->ev_run[slot], 0), "wait");
      for (uint32_t c = 0; c < C && s1 > s0; c++)
        CU(cudaMemcpyAsync((int16_t *)gpu->ring_out[slot] + (uint64_t)c * (s1 - s0), d_pcm + c * pitch + s0, (size_t)(s1 - s0) * 2,
                           cudaMemcpyDeviceToHost, gpu->s_out), "D2H pcm");
      CU(cudaEventRecord(gpu->ev_out[slot], gpu->s_out), "event");
      PHASE(2);
    }
    if (k >= lag) {
      const uint32_t j = k - lag, slot = j % AADGPU_RING_SLOTS;
      const uint32_t b0 = j * per, b1 = (b0 + per < num_blocks) ? b0 + per : num_blocks;
      const uint64_t s0 = (uint64_t)b0 * spb, s1 = ((uint64_t)b1 * spb < total) ? (uint64_t)b1 * spb : total;
      CU(cudaEventSynchronize(gpu->ev_out[slot]), "ring wait");
      PHASE(3);
      if (s1 > s0) convert_rows(1, (int16_t *)gpu->ring_out[slot], buffer, C, s0, s1 - s0);
      PHASE(4);
    }
  }
#undef PHASE
  if (trace)
    fprintf(stderr, "[aad_b200] drop-in decode: %u slices of %u blocks: ring wait %.2f ms, staging copy %.2f, enqueue %.2f, result wait %.2f, "
            "widening %.2f\n", slices, per, 1e3 * t_phase[0], 1e3 * t_phase[1], 1e3 * t_phase[2], 1e3 * t_phase[3], 1e3 * t_phase[4]);
  return AAD_APIRESULT_OK;
}

AADApiResult aadgpu_decode_stream_i32(struct AADGpu *gpu, const struct aadf_geometry *geo, const uint8_t *data,
                                      uint32_t data_size, uint32_t num_blocks, uint32_t num_samples,
                                      uint32_t buf_samples, int32_t *const *buffer)
{
  WITH_CONTEXT_LOCK(gpu, aadgpu_decode_stream_i32_unlocked(gpu, geo, data, data_size, num_blocks, num_samples, buf_samples, buffer));
}

/* ---- WAV-order (interleaved int16) single-stream paths: what `aad -e / -d / -r` do ---------- */

/* src/main.c:175-179 + AADEncoder_EncodeWhole, with the de-interleave done on the device: the
 * samples of a 16-bit WAV data chunk go to HBM as they are.  Leaves the stream at gpu->aad + 1. */
static AADApiResult encode_wav_device(struct AADGpu *gpu, const struct AADEncodeParameter *prm,
                                      const struct aadf_geometry *geo, const void *interleaved, uint32_t wav_bits,
                                      uint32_t num_samples, uint64_t *pitch_out, uint64_t *bytes_out);

static AADApiResult encode_interleaved_device(struct AADGpu *gpu, const struct AADEncodeParameter *prm,
                                              const struct aadf_geometry *geo, const int16_t *interleaved,
                                              uint32_t num_samples, uint64_t *pitch_out, uint64_t *bytes_out)
{
  return encode_wav_device(gpu, prm, geo, interleaved, 16, num_samples, pitch_out, bytes_out);
}

/* interleaved: the samples as they lie in a WAV data chunk of wav_bits bits per sample (16: plain int16), or NULL
 * when the caller has put the planar int16 samples into gpu->pcm already (same pitch) */
static AADApiResult encode_wav_device(struct AADGpu *gpu, const struct AADEncodeParameter *prm,
                                      const struct aadf_geometry *geo, const void *interleaved, uint32_t wav_bits,
                                      uint32_t num_samples, uint64_t *pitch_out, uint64_t *bytes_out)
{
  const uint32_t C = geo->channels;
  const uint64_t pitch = round_up64(num_samples, 64);
  const uint64_t bytes = aadf_stream_bytes(num_samples, C, geo->bits, geo->block_size, geo->samples_per_block);
  const uint64_t bound = aadf_stream_bytes_bound(num_samples, geo->block_size, geo->samples_per_block);
  if (!aadgpu_reserve(gpu, &gpu->pcm, (size_t)C * pitch * 2)) return AAD_APIRESULT_NG;
  if (!aadgpu_reserve(gpu, &gpu->aad, (size_t)bound + 128)) return AAD_APIRESULT_NG;
  cudaStream_t s = gpu->s_run;
  if (interleaved != NULL) {
    /* the data chunk goes up in pieces on the copy stream; each piece is narrowed to 16 bits (8 / 24 / 32-bit
     * chunks: top 16 bits) and de-interleaved on the kernel stream while the next one is on the link */
    const uint32_t sample_bytes = wav_bits / 8;
    const size_t raw_bytes = (size_t)C * num_samples * sample_bytes;
    struct aadgpu_buffer *up = (wav_bits == 16) ? &gpu->wav : &gpu->raw;
    if (!aadgpu_reserve(gpu, up, raw_bytes)) return AAD_APIRESULT_NG;
    const uint32_t pieces = pick_stream_slices(raw_bytes, num_samples);
    for (uint32_t k = 0; k < pieces; k++) {
      const uint64_t a = (uint64_t)num_samples * k / pieces, b = (uint64_t)num_samples * (k + 1) / pieces;
      if (b == a) continue;
      const size_t off = (size_t)a * C * sample_bytes;
      CU(cudaMemcpyAsync((uint8_t *)up->ptr + off, (const uint8_t *)interleaved + off, (size_t)(b - a) * C * sample_bytes,
                         cudaMemcpyHostToDevice, gpu->s_in), "H2D wav data");
      CU(cudaEventRecord(gpu->ev_in[k], gpu->s_in), "event");
      CU(cudaStreamWaitEvent(s, gpu->ev_in[k], 0), "wait");
      if (wav_bits == 16)
        CU((cudaError_t)aadk_launch_deinterleave16((const int16_t *)up->ptr + a * C, (int16_t *)gpu->pcm.ptr + a, pitch, C,
                                                   (uint32_t)(b - a), s), "deinterleave kernel launch");
      else
        CU((cudaError_t)aadk_launch_wav_to_planar16((const uint8_t *)up->ptr + off, wav_bits, (int16_t *)gpu->pcm.ptr + a, pitch, C,
                                                    (uint32_t)(b - a), s), "wav_to_planar16 kernel launch");
    }
  }
  CU(cudaMemsetAsync(gpu->aad.ptr, 0, (size_t)bound + 128, s), "memset aad");
  struct aadk_encode_params p;
  memset(&p, 0, sizeof(p));
  p.pcm = gpu->pcm.ptr;
  p.pcm_clip_stride = (uint64_t)C * pitch;
  p.pcm_ch_stride = pitch;
  p.uniform_samples = num_samples;
  p.num_streams = 1;
  p.geo = *geo;
  p.sampling_rate = prm->sampling_rate;
  p.trials = prm->num_encode_trials;
  p.aad = (uint8_t *)gpu->aad.ptr + 1;   /* block 0 lands 32-byte aligned */
  p.aad_stride = bound;
  p.block_begin = 0;
  p.block_end = aadf_num_blocks(num_samples, geo->samples_per_block);
  (void)apply_segments(gpu, &p, p.block_end);
  CU((cudaError_t)aadk_launch_encode(&p, s), "encode kernel launch");
  *pitch_out = pitch;
  *bytes_out = bytes;
  return AAD_APIRESULT_OK;
}

static AADApiResult encode_stream_range_unlocked(struct AADGpu *gpu, const struct AADEncodeParameter *prm, uint32_t segment_blocks,
                                        const int16_t *interleaved, uint32_t num_samples, uint32_t b0, uint32_t b1,
                                        uint8_t *data);

static AADApiResult AADGpu_EncodeInterleaved16_unlocked(struct AADGpu *gpu, const struct AADEncodeParameter *prm,
                                        const int16_t *interleaved, uint32_t num_samples, uint8_t *data,
                                        uint32_t data_size, uint32_t *output_size)
{
  if (!gpu || !prm || !interleaved || !data || !output_size) return AAD_APIRESULT_INVALID_ARGUMENT;
  struct aadf_geometry geo;
  const AADApiResult r = check_encode_shape(prm, num_samples, &geo);
  if (r != AAD_APIRESULT_OK) return r;
  if (data_size < AADF_FILE_HEADER_BYTES) return AAD_APIRESULT_INSUFFICIENT_DATA;
  CU(cudaSetDevice(gpu->device), "cudaSetDevice");
  uint64_t pitch = 0, bytes = 0;
  if (aadf_stream_bytes(num_samples, geo.channels, geo.bits, geo.block_size, geo.samples_per_block) > data_size)
    return AAD_APIRESULT_INSUFFICIENT_BUFFER;
  if (gpu->segment_blocks != 0) {
    /* segment mode: the pipeline that is cut across the segments (also what a device group runs per shard) */
    const AADApiResult es = encode_stream_range_unlocked(gpu, prm, gpu->segment_blocks, interleaved, num_samples, 0,
                                                         aadf_num_blocks(num_samples, geo.samples_per_block), data);
    if (es == AAD_APIRESULT_OK)
      *output_size = (uint32_t)aadf_stream_bytes(num_samples, geo.channels, geo.bits, geo.block_size, geo.samples_per_block);
    return es;
  }
  const AADApiResult e = encode_interleaved_device(gpu, prm, &geo, interleaved, num_samples, &pitch, &bytes);
  if (e != AAD_APIRESULT_OK) return e;
  CU(cudaMemcpyAsync(data, (uint8_t *)gpu->aad.ptr + 1, (size_t)bytes, cudaMemcpyDeviceToHost, gpu->s_run), "D2H aad");
  CU(cudaStreamSynchronize(gpu->s_run), "sync");
  *output_size = (uint32_t)bytes;
  return AAD_APIRESULT_OK;
}

AADApiResult AADGpu_EncodeInterleaved16(struct AADGpu *gpu, const struct AADEncodeParameter *prm,
                                        const int16_t *interleaved, uint32_t num_samples, uint8_t *data,
                                        uint32_t data_size, uint32_t *output_size)
{
  WITH_CONTEXT_LOCK(gpu, AADGpu_EncodeInterleaved16_unlocked(gpu, prm, interleaved, num_samples, data, data_size, output_size));
}

static AADApiResult AADGpu_EncodeWav_unlocked(struct AADGpu *gpu, const struct AADEncodeParameter *prm, const uint8_t *wav_data,
                                              uint32_t wav_bits_per_sample, uint32_t num_samples, uint8_t *data,
                                              uint32_t data_size, uint32_t *output_size)
{
  if (!gpu || !prm || !wav_data || !data || !output_size) return AAD_APIRESULT_INVALID_ARGUMENT;
  const uint32_t wb = wav_bits_per_sample;
  if (wb != 8 && wb != 16 && wb != 24 && wb != 32) return AAD_APIRESULT_INVALID_FORMAT;   /* src/wav.c:222-238 */
  struct aadf_geometry geo;
  const AADApiResult r = check_encode_shape(prm, num_samples, &geo);
  if (r != AAD_APIRESULT_OK) return r;
  if (data_size < AADF_FILE_HEADER_BYTES) return AAD_APIRESULT_INSUFFICIENT_DATA;
  if (aadf_stream_bytes(num_samples, geo.channels, geo.bits, geo.block_size, geo.samples_per_block) > data_size)
    return AAD_APIRESULT_INSUFFICIENT_BUFFER;
  CU(cudaSetDevice(gpu->device), "cudaSetDevice");
  uint64_t pitch = 0, bytes = 0;
  const AADApiResult e = encode_wav_device(gpu, prm, &geo, wav_data, wb, num_samples, &pitch, &bytes);
  if (e != AAD_APIRESULT_OK) return e;
  CU(cudaMemcpyAsync(data, (uint8_t *)gpu->aad.ptr + 1, (size_t)bytes, cudaMemcpyDeviceToHost, gpu->s_run), "D2H aad");
  CU(cudaStreamSynchronize(gpu->s_run), "sync");
  *output_size = (uint32_t)bytes;
  return AAD_APIRESULT_OK;
}

AADApiResult AADGpu_EncodeWav(struct AADGpu *gpu, const struct AADEncodeParameter *prm, const uint8_t *wav_data,
                              uint32_t wav_bits_per_sample, uint32_t num_samples, uint8_t *data, uint32_t data_size,
                              uint32_t *output_size)
{
  WITH_CONTEXT_LOCK(gpu, AADGpu_EncodeWav_unlocked(gpu, prm, wav_data, wav_bits_per_sample, num_samples, data, data_size, output_size));
}

static void geometry_of_header(const struct AADHeaderInfo *h, struct aadf_geometry *geo)
{
  geo->channels = h->num_channels;
  geo->bits = h->bits_per_sample;
  geo->block_size = h->block_size;
  geo->samples_per_block = h->num_samples_per_block;
  geo->ms = (h->ch_process_method == AAD_CH_PROCESS_METHOD_MS) ? 1u : 0u;
}

/* Blocks [b0, b1) of the stream whose byte `byte_base` lies at d_aad, into WAV order at d_wav (frame
 * `sample_base` first).  The decoder's flush writes the frames itself where the shape allows (mono, 2 / 4 / 8
 * channels on the staged kernels); otherwise the planes go to gpu->pcm and one more pass interleaves them. */
static AADApiResult launch_decode_wav_order(struct AADGpu *gpu, const struct aadf_geometry *geo, const uint8_t *d_aad,
                                            uint64_t byte_base, uint64_t valid_size, uint32_t num_samples, uint32_t b0,
                                            uint32_t b1, int16_t *d_wav, uint64_t sample_base, uint64_t shard_samples,
                                            cudaStream_t s)
{
  const uint32_t C = geo->channels, spb = geo->samples_per_block;
  struct aadk_decode_params p;
  memset(&p, 0, sizeof(p));
  p.aad = d_aad;
  p.aad_stride = 0;
  p.uniform_size = (uint32_t)valid_size;
  p.num_streams = 1;
  p.geo = *geo;
  p.block_begin = b0;
  p.block_end = b1;
  p.uniform_samples = num_samples;
  p.byte_base = byte_base;
  p.sample_base = sample_base;
  p.pcm = d_wav;
  p.interleaved = 1;
  if (aadk_decode_interleaved_ok(&p)) {
    CU((cudaError_t)aadk_launch_decode(&p, s), "decode kernel launch");
    return AAD_APIRESULT_OK;
  }
  const uint64_t pitch = round_up64(shard_samples, 64);
  if (!aadgpu_reserve(gpu, &gpu->pcm, (size_t)C * pitch * 2)) return AAD_APIRESULT_NG;
  p.interleaved = 0;
  p.pcm = gpu->pcm.ptr;
  p.pcm_ch_stride = pitch;
  CU((cudaError_t)aadk_launch_decode(&p, s), "decode kernel launch");
  const uint64_t f0 = (uint64_t)b0 * spb - sample_base;
  const uint64_t f1 = ((uint64_t)b1 * spb < num_samples ? (uint64_t)b1 * spb : num_samples) - sample_base;
  if (f1 > f0)
    CU((cudaError_t)aadk_launch_interleave16((const int16_t *)gpu->pcm.ptr + f0, pitch, d_wav + f0 * C, C, (uint32_t)(f1 - f0), s),
       "interleave kernel launch");
  return AAD_APIRESULT_OK;
}

/* decode the complete stream at d_aad (device, just encoded there) into gpu->wav in WAV order */
static AADApiResult decode_to_interleaved_device(struct AADGpu *gpu, const struct AADHeaderInfo *h, const uint8_t *d_aad,
                                                 uint64_t aad_bytes)
{
  struct aadf_geometry geo;
  geometry_of_header(h, &geo);
  const uint32_t C = geo.channels, ns = h->num_samples;
  if (!aadgpu_reserve(gpu, &gpu->wav, (size_t)C * ns * 2)) return AAD_APIRESULT_NG;
  return launch_decode_wav_order(gpu, &geo, d_aad, 0, aad_bytes, ns, 0, aadf_num_blocks(ns, geo.samples_per_block),
                                 (int16_t *)gpu->wav.ptr, 0, ns, gpu->s_run);
}

/* the checks of AADDecoder_DecodeHeader + CheckHeaderFormat (src/aad_decoder.c:99-225), via the drop-in entry */
static AADApiResult parse_stream_header(const uint8_t *data, uint32_t data_size, struct AADHeaderInfo *h)
{
  const AADApiResult r = AADDecoder_DecodeHeader(data, data_size, h);
  if (r != AAD_APIRESULT_OK) return r;
  return aaddec_check_header(h);
}

/* The block loop of AADDecoder_DecodeWhole (src/aad_decoder.c:514-534): block b is visited while b * spb <
 * num_samples and its first byte exists; a last block too short for its channel headers ends the loop with
 * AAD_APIRESULT_INSUFFICIENT_DATA (*tail) after the earlier blocks were decoded.  Returns the blocks to decode. */
static uint32_t stream_block_span(const struct AADHeaderInfo *h, uint32_t data_size, AADApiResult *tail)
{
  const uint32_t by_samples = aadf_num_blocks(h->num_samples, h->num_samples_per_block);
  const uint64_t payload = (uint64_t)data_size - AADF_FILE_HEADER_BYTES;
  const uint64_t by_bytes = (payload + h->block_size - 1) / h->block_size;
  uint32_t blocks = (uint32_t)(by_bytes < by_samples ? by_bytes : by_samples);
  *tail = AAD_APIRESULT_OK;
  if (blocks > 0) {
    const uint64_t last_avail = payload - (uint64_t)(blocks - 1) * h->block_size;
    if (last_avail < (uint64_t)AADF_CHANNEL_HEADER_BYTES * h->num_channels) {
      blocks--;
      *tail = AAD_APIRESULT_INSUFFICIENT_DATA;
    }
  }
  return blocks;
}

/* samples (per channel) that `blocks` leading blocks deliver */
static uint64_t stream_decoded_samples(const struct AADHeaderInfo *h, uint32_t blocks)
{
  const uint64_t n = (uint64_t)blocks * h->num_samples_per_block;
  return n < h->num_samples ? n : h->num_samples;
}

static AADApiResult decode_stream_range_unlocked(struct AADGpu *gpu, const struct AADHeaderInfo *h, const uint8_t *data,
                                        uint32_t data_size, uint32_t b0, uint32_t b1, int16_t *interleaved);

static AADApiResult AADGpu_DecodeInterleaved16_unlocked(struct AADGpu *gpu, const uint8_t *data, uint32_t data_size,
                                        int16_t *interleaved, uint32_t capacity_samples)
{
  if (!gpu || !data || !interleaved) return AAD_APIRESULT_INVALID_ARGUMENT;
  struct AADHeaderInfo h;
  const AADApiResult r = parse_stream_header(data, data_size, &h);
  if (r != AAD_APIRESULT_OK) return r;
  if (capacity_samples < h.num_samples) return AAD_APIRESULT_INSUFFICIENT_BUFFER;
  AADApiResult tail;
  const uint32_t blocks = stream_block_span(&h, data_size, &tail);
  /* samples of blocks the data does not reach are zero (the reference leaves them untouched) */
  const uint64_t decoded = stream_decoded_samples(&h, blocks);
  if (decoded < h.num_samples)
    memset(interleaved + decoded * h.num_channels, 0, (size_t)(h.num_samples - decoded) * h.num_channels * 2);
  const AADApiResult e = decode_stream_range_unlocked(gpu, &h, data, data_size, 0, blocks, interleaved);
  return (e != AAD_APIRESULT_OK) ? e : tail;
}

AADApiResult AADGpu_DecodeInterleaved16(struct AADGpu *gpu, const uint8_t *data, uint32_t data_size,
                                        int16_t *interleaved, uint32_t capacity_samples)
{
  WITH_CONTEXT_LOCK(gpu, AADGpu_DecodeInterleaved16_unlocked(gpu, data, data_size, interleaved, capacity_samples));
}

/* src/main.c:275-346 (execute_reconstruction_core): encode, then decode what was encoded; the
 * stream never leaves the device. */
static AADApiResult AADGpu_ReconstructInterleaved16_unlocked(struct AADGpu *gpu, const struct AADEncodeParameter *prm,
                                             const int16_t *interleaved, uint32_t num_samples,
                                             int16_t *reconstructed, uint32_t *encoded_size)
{
  if (!gpu || !prm || !interleaved || !reconstructed) return AAD_APIRESULT_INVALID_ARGUMENT;
  struct aadf_geometry geo;
  const AADApiResult r = check_encode_shape(prm, num_samples, &geo);
  if (r != AAD_APIRESULT_OK) return r;
  CU(cudaSetDevice(gpu->device), "cudaSetDevice");
  uint64_t pitch = 0, bytes = 0;
  AADApiResult e = encode_interleaved_device(gpu, prm, &geo, interleaved, num_samples, &pitch, &bytes);
  if (e != AAD_APIRESULT_OK) return e;
  struct AADHeaderInfo h;
  memset(&h, 0, sizeof(h));
  h.num_channels = (uint16_t)geo.channels;
  h.num_samples = num_samples;
  h.sampling_rate = prm->sampling_rate;
  h.bits_per_sample = (uint16_t)geo.bits;
  h.block_size = (uint16_t)geo.block_size;
  h.num_samples_per_block = geo.samples_per_block;
  h.ch_process_method = prm->ch_process_method;
  e = decode_to_interleaved_device(gpu, &h, (const uint8_t *)gpu->aad.ptr + 1, bytes);
  if (e != AAD_APIRESULT_OK) return e;
  CU(cudaMemcpyAsync(reconstructed, gpu->wav.ptr, (size_t)geo.channels * num_samples * 2, cudaMemcpyDeviceToHost, gpu->s_run),
     "D2H wav");
  CU(cudaStreamSynchronize(gpu->s_run), "sync");
  if (encoded_size) *encoded_size = (uint32_t)bytes;
  return AAD_APIRESULT_OK;
}

AADApiResult AADGpu_ReconstructInterleaved16(struct AADGpu *gpu, const struct AADEncodeParameter *prm,
                                             const int16_t *interleaved, uint32_t num_samples,
                                             int16_t *reconstructed, uint32_t *encoded_size)
{
  WITH_CONTEXT_LOCK(gpu, AADGpu_ReconstructInterleaved16_unlocked(gpu, prm, interleaved, num_samples, reconstructed, encoded_size));
}

/* ---- the analysis modes of the command line, src/main.c:275-503, with the samples staying on the device -------- */

static void header_of(const struct AADEncodeParameter *prm, const struct aadf_geometry *geo, uint32_t num_samples,
                      struct AADHeaderInfo *h)
{
  memset(h, 0, sizeof(*h));
  h->num_channels = (uint16_t)geo->channels;
  h->num_samples = num_samples;
  h->sampling_rate = prm->sampling_rate;
  h->bits_per_sample = (uint16_t)geo->bits;
  h->block_size = (uint16_t)geo->block_size;
  h->num_samples_per_block = geo->samples_per_block;
  h->ch_process_method = prm->ch_process_method;
}

static AADApiResult AADGpu_AnalyzeWav_unlocked(struct AADGpu *gpu, const struct AADEncodeParameter *prm, const uint8_t *wav_data,
                               uint32_t wav_bits_per_sample, uint32_t num_samples, enum AADGpuAnalysis what,
                               uint8_t *out_data, double stats[3], uint32_t *encoded_size)
{
  if (!gpu || !prm || !wav_data) return AAD_APIRESULT_INVALID_ARGUMENT;
  if (what == AADGPU_ANALYSIS_STATISTICS ? stats == NULL : out_data == NULL) return AAD_APIRESULT_INVALID_ARGUMENT;
  if (what != AADGPU_ANALYSIS_RECONSTRUCT && what != AADGPU_ANALYSIS_RESIDUAL && what != AADGPU_ANALYSIS_STATISTICS)
    return AAD_APIRESULT_INVALID_ARGUMENT;
  const uint32_t wb = wav_bits_per_sample;
  if (wb != 8 && wb != 16 && wb != 24 && wb != 32) return AAD_APIRESULT_INVALID_FORMAT;   /* src/wav.c:222-238 */
  struct aadf_geometry geo;
  const AADApiResult r = check_encode_shape(prm, num_samples, &geo);
  if (r != AAD_APIRESULT_OK) return r;
  CU(cudaSetDevice(gpu->device), "cudaSetDevice");
  const uint32_t C = geo.channels;
  const uint64_t count = (uint64_t)C * num_samples;
  const size_t raw_bytes = (size_t)count * (wb / 8);
  const uint64_t in_pitch = round_up64(num_samples, 64);
  if (!aadgpu_reserve(gpu, &gpu->raw, raw_bytes)) return AAD_APIRESULT_NG;
  if (!aadgpu_reserve(gpu, &gpu->pcm, (size_t)C * in_pitch * 2)) return AAD_APIRESULT_NG;
  cudaStream_t s = gpu->s_run;
  /* the data chunk goes up as it lies in the file and stays in gpu->raw for the mode's arithmetic; narrowing to
   * 16 bits and de-interleaving in one kernel */
  CU(cudaMemcpyAsync(gpu->raw.ptr, wav_data, raw_bytes, cudaMemcpyHostToDevice, s), "H2D wav data");
  CU((cudaError_t)aadk_launch_wav_to_planar16((const uint8_t *)gpu->raw.ptr, wb, (int16_t *)gpu->pcm.ptr, in_pitch, C, num_samples, s),
     "wav_to_planar16 kernel launch");
  uint64_t pitch = 0, bytes = 0;
  AADApiResult e = encode_wav_device(gpu, prm, &geo, NULL, 16, num_samples, &pitch, &bytes);
  if (e != AAD_APIRESULT_OK) return e;
  struct AADHeaderInfo h;
  header_of(prm, &geo, num_samples, &h);
  e = decode_to_interleaved_device(gpu, &h, (const uint8_t *)gpu->aad.ptr + 1, bytes);   /* -> gpu->wav, interleaved int16 */
  if (e != AAD_APIRESULT_OK) return e;
  if (what == AADGPU_ANALYSIS_STATISTICS) {
    double partials[3 * AADK_STATS_BLOCKS];
    if (!aadgpu_reserve(gpu, &gpu->stats, sizeof(partials))) return AAD_APIRESULT_NG;
    CU((cudaError_t)aadk_launch_analysis_stats((const uint8_t *)gpu->raw.ptr, wb, (const int16_t *)gpu->wav.ptr, count,
                                               (double *)gpu->stats.ptr, s), "analysis_stats kernel launch");
    CU(cudaMemcpyAsync(partials, gpu->stats.ptr, sizeof(partials), cudaMemcpyDeviceToHost, s), "D2H partial sums");
    CU(cudaStreamSynchronize(s), "sync");
    double sq = 0.0, ab = 0.0, mx = 0.0;
    for (int b = 0; b < AADK_STATS_BLOCKS; b++) {
      sq += partials[3 * b];
      ab += partials[3 * b + 1];
      if (mx < partials[3 * b + 2]) mx = partials[3 * b + 2];
    }
    stats[0] = sqrt(sq / (double)count);    /* RMSE, src/main.c:492-494 */
    stats[1] = ab / (double)count;          /* MSD */
    stats[2] = mx;                          /* MaxAE */
  } else {
    CU((cudaError_t)aadk_launch_analysis_image((uint8_t *)gpu->raw.ptr, wb, (const int16_t *)gpu->wav.ptr, count,
                                               what == AADGPU_ANALYSIS_RESIDUAL, s), "analysis_image kernel launch");
    CU(cudaMemcpyAsync(out_data, gpu->raw.ptr, raw_bytes, cudaMemcpyDeviceToHost, s), "D2H wav data");
    CU(cudaStreamSynchronize(s), "sync");
  }
  if (encoded_size) *encoded_size = (uint32_t)bytes;
  return AAD_APIRESULT_OK;
}

AADApiResult AADGpu_AnalyzeWav(struct AADGpu *gpu, const struct AADEncodeParameter *prm, const uint8_t *wav_data,
                               uint32_t wav_bits_per_sample, uint32_t num_samples, enum AADGpuAnalysis what,
                               uint8_t *out_data, double stats[3], uint32_t *encoded_size)
{
  WITH_CONTEXT_LOCK(gpu, AADGpu_AnalyzeWav_unlocked(gpu, prm, wav_data, wav_bits_per_sample, num_samples, what, out_data, stats, encoded_size));
}

/* ---- several devices of one box: shard, run one host thread per device, done ---------------- */
/*
 * The codec has no exchange step: batches shard by stream, one long stream
 * shards its DECODE by block range (every block header reloads the chain state,
 * src/aad_decoder.c:364-380).  Each device writes its own disjoint slice of the caller's
 * buffers, so reassembly is the layout itself; no NCCL, no peer copies.  One stream cannot be
 * ENCODED in shards bit-exactly (src/aad_encoder.c:853-886 carries state), so there is no such call.
 */
struct AADGpuGroup {
  int size;
  struct AADGpu *gpu[AADGPU_MAX_GROUP];
};

struct AADGpuGroup *AADGpuGroup_Create(const int *devices, int num_devices)
{
  const int visible = AADGpu_DeviceCount();
  if (devices == NULL) num_devices = (num_devices <= 0 || num_devices > visible) ? visible : num_devices;
  if (num_devices <= 0 || num_devices > AADGPU_MAX_GROUP) {
    aadgpu_set_error("AADGpuGroup_Create: no usable CUDA devices (this library has no CPU fallback)");
    return NULL;
  }
  struct AADGpuGroup *g = (struct AADGpuGroup *)calloc(1, sizeof(*g));
  if (!g) return NULL;
  for (int i = 0; i < num_devices; i++) {
    g->gpu[i] = AADGpu_Create(devices ? devices[i] : i);
    if (g->gpu[i] == NULL) {
      AADGpuGroup_Destroy(g);
      return NULL;
    }
    g->size++;
  }
  return g;
}

void AADGpuGroup_Destroy(struct AADGpuGroup *g)
{
  if (!g) return;
  for (int i = 0; i < g->size; i++) AADGpu_Destroy(g->gpu[i]);
  free(g);
}

int AADGpuGroup_Size(const struct AADGpuGroup *g) { return g ? g->size : 0; }
struct AADGpu *AADGpuGroup_Device(const struct AADGpuGroup *g, int index) { return (g && index >= 0 && index < g->size) ? g->gpu[index] : NULL; }

/* contiguous, balanced: the first (n % parts) shards get one more (aad_b200/shard.py: split_range) */
static void split_range(uint64_t n, int parts, int index, uint64_t *begin, uint64_t *end)
{
  const uint64_t base = n / (uint64_t)parts, extra = n % (uint64_t)parts;
  const uint64_t i = (uint64_t)index;
  *begin = i * base + (i < extra ? i : extra);
  *end = *begin + base + (i < extra ? 1u : 0u);
}

enum group_op { GROUP_ENCODE_BATCH, GROUP_DECODE_BATCH, GROUP_DECODE_STREAM, GROUP_ENCODE_STREAM };

struct group_task {
  enum group_op op;
  struct AADGpu *gpu;
  struct AADGpuBatch batch;
  const int16_t *pcm_in;
  int16_t *pcm_out;
  const uint32_t *lens;
  const uint8_t *aad_in;
  uint8_t *aad_out;
  uint32_t *sizes_out;
  /* GROUP_DECODE_STREAM */
  struct AADHeaderInfo header;
  uint32_t data_size, block_begin, block_end;
  /* GROUP_ENCODE_STREAM */
  struct AADEncodeParameter param;
  uint32_t num_samples, segment_blocks;
  int bind;
  AADApiResult result;
  char error[320];
};

static AADApiResult decode_stream_range(struct AADGpu *gpu, const struct AADHeaderInfo *h, const uint8_t *data,
                                        uint32_t data_size, uint32_t b0, uint32_t b1, int16_t *interleaved);

static AADApiResult encode_stream_range(struct AADGpu *gpu, const struct AADEncodeParameter *prm, uint32_t segment_blocks,
                                        const int16_t *interleaved, uint32_t num_samples, uint32_t b0, uint32_t b1,
                                        uint8_t *data);

/* AAD_B200_TRACE=1: per-shard wall times of the group calls on stderr (diagnosis of multi-device runs) */
static int trace_on(void)
{
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("AAD_B200_TRACE");
    on = (e && atoi(e) > 0) ? 1 : 0;
  }
  return on;
}

static void *group_worker(void *arg)
{
  struct group_task *t = (struct group_task *)arg;
  const double t_start = trace_on() ? wall_seconds() : 0.0;
  if (t->bind) (void)AADGpu_BindHostThread(t->gpu);   /* a thread of its own: stay next to its device */
  const double t_bound = trace_on() ? wall_seconds() : 0.0;
  switch (t->op) {
    case GROUP_ENCODE_BATCH:
      t->result = AADGpu_EncodeBatch(t->gpu, &t->batch, t->pcm_in, t->lens, t->aad_out, t->sizes_out);
      break;
    case GROUP_DECODE_BATCH:
      t->result = AADGpu_DecodeBatch(t->gpu, &t->batch, t->aad_in, t->lens, t->pcm_out);
      break;
    case GROUP_ENCODE_STREAM:
      t->result = encode_stream_range(t->gpu, &t->param, t->segment_blocks, t->pcm_in, t->num_samples, t->block_begin,
                                      t->block_end, t->aad_out);
      break;
    default:
      t->result = decode_stream_range(t->gpu, &t->header, t->aad_in, t->data_size, t->block_begin, t->block_end, t->pcm_out);
      break;
  }
  snprintf(t->error, sizeof(t->error), "%s", AADGpu_LastError());   /* thread-local: carry it to the caller */
  if (trace_on())
    fprintf(stderr, "[aad_b200] group shard on device %d: op %d blocks [%u, %u): bind %.3f ms, work %.3f ms\n", t->gpu->device,
            (int)t->op, t->block_begin, t->block_end, 1e3 * (t_bound - t_start), 1e3 * (wall_seconds() - t_bound));
  return NULL;
}

static AADApiResult group_run(struct group_task *tasks, int n)
{
  pthread_t th[AADGPU_MAX_GROUP];
  int started[AADGPU_MAX_GROUP];
  const double t_start = trace_on() ? wall_seconds() : 0.0;
  for (int i = 0; i < n; i++) {
    tasks[i].bind = 1;
    started[i] = pthread_create(&th[i], NULL, group_worker, &tasks[i]) == 0;
    if (!started[i]) {                           /* no thread: do it here, still correct */
      tasks[i].bind = 0;
      group_worker(&tasks[i]);
    }
  }
  AADApiResult r = AAD_APIRESULT_OK;
  for (int i = 0; i < n; i++) {
    if (started[i]) pthread_join(th[i], NULL);
    if (tasks[i].result != AAD_APIRESULT_OK && r == AAD_APIRESULT_OK) {
      r = tasks[i].result;
      aadgpu_set_error(tasks[i].error);
    }
  }
  if (trace_on()) fprintf(stderr, "[aad_b200] group call over %d shard(s): %.3f ms\n", n, 1e3 * (wall_seconds() - t_start));
  return r;
}

AADApiResult AADGpuGroup_EncodeBatch(struct AADGpuGroup *g, const struct AADGpuBatch *batch, const int16_t *pcm,
                                     const uint32_t *num_samples, uint8_t *aad, uint32_t *out_sizes)
{
  if (!g || !batch || !pcm || !aad) return AAD_APIRESULT_INVALID_ARGUMENT;
  struct group_task tasks[AADGPU_MAX_GROUP];
  int n = 0;
  for (int d = 0; d < g->size; d++) {
    uint64_t i0, i1;
    split_range(batch->num_streams, g->size, d, &i0, &i1);
    if (i1 == i0) continue;
    struct group_task *t = &tasks[n++];
    memset(t, 0, sizeof(*t));
    t->op = GROUP_ENCODE_BATCH;
    t->gpu = g->gpu[d];
    t->batch = *batch;
    t->batch.num_streams = (uint32_t)(i1 - i0);
    t->pcm_in = pcm + i0 * batch->pcm_stream_stride;
    t->lens = num_samples ? num_samples + i0 : NULL;
    t->aad_out = aad + i0 * batch->aad_stream_stride;
    t->sizes_out = out_sizes ? out_sizes + i0 : NULL;
  }
  /* an empty batch still gets the argument checks of the single-device call */
  if (n == 0) return AADGpu_EncodeBatch(g->gpu[0], batch, pcm, num_samples, aad, out_sizes);
  return group_run(tasks, n);
}

AADApiResult AADGpuGroup_DecodeBatch(struct AADGpuGroup *g, const struct AADGpuBatch *batch, const uint8_t *aad,
                                     const uint32_t *sizes, int16_t *pcm)
{
  if (!g || !batch || !pcm || !aad) return AAD_APIRESULT_INVALID_ARGUMENT;
  struct group_task tasks[AADGPU_MAX_GROUP];
  int n = 0;
  for (int d = 0; d < g->size; d++) {
    uint64_t i0, i1;
    split_range(batch->num_streams, g->size, d, &i0, &i1);
    if (i1 == i0) continue;
    struct group_task *t = &tasks[n++];
    memset(t, 0, sizeof(*t));
    t->op = GROUP_DECODE_BATCH;
    t->gpu = g->gpu[d];
    t->batch = *batch;
    t->batch.num_streams = (uint32_t)(i1 - i0);
    t->aad_in = aad + i0 * batch->aad_stream_stride;
    t->lens = sizes ? sizes + i0 : NULL;
    t->pcm_out = pcm + i0 * batch->pcm_stream_stride;
  }
  if (n == 0) return AADGpu_DecodeBatch(g->gpu[0], batch, aad, sizes, pcm);
  return group_run(tasks, n);
}

/* Blocks [b0, b1) of one stream -> interleaved samples [b0*spb, min(b1*spb, ns)) of the caller's buffer; every
 * block of the range has its channel headers (stream_block_span).  The device holds only the range's own bytes and
 * samples.  Sliced by block range over three streams: while slice k decodes, slice k+1's bytes come up and slice
 * k-1's frames go down; the decoder writes WAV order itself, so the samples cross HBM once. */
static AADApiResult decode_stream_range_unlocked(struct AADGpu *gpu, const struct AADHeaderInfo *h, const uint8_t *data,
                                        uint32_t data_size, uint32_t b0, uint32_t b1, int16_t *interleaved)
{
  struct aadf_geometry geo;
  geometry_of_header(h, &geo);
  const uint32_t C = geo.channels, spb = geo.samples_per_block, bs = geo.block_size, ns = h->num_samples;
  const uint64_t s0 = (uint64_t)b0 * spb, s1 = ((uint64_t)b1 * spb < ns) ? (uint64_t)b1 * spb : ns;
  if (b1 <= b0 || s1 <= s0) return AAD_APIRESULT_OK;
  const double t_enter = trace_on() ? wall_seconds() : 0.0;
  CU(cudaSetDevice(gpu->device), "cudaSetDevice");
  const uint64_t count = s1 - s0;
  const uint64_t byte0 = AADF_FILE_HEADER_BYTES + (uint64_t)b0 * bs;
  uint64_t byte1 = AADF_FILE_HEADER_BYTES + (uint64_t)b1 * bs;
  if (byte1 > data_size) byte1 = data_size;
  const uint64_t span = byte1 > byte0 ? byte1 - byte0 : 0;
  if (!aadgpu_reserve(gpu, &gpu->aad, (size_t)span + 256)) return AAD_APIRESULT_NG;
  if (!aadgpu_reserve(gpu, &gpu->wav, (size_t)C * count * 2)) return AAD_APIRESULT_NG;
  uint8_t *d_shard = (uint8_t *)gpu->aad.ptr;       /* block b0 at the (256-byte aligned) start */
  int16_t *d_wav = (int16_t *)gpu->wav.ptr;
  const uint32_t slices = pick_stream_slices(count * C * 2, b1 - b0);
  for (uint32_t k = 0; k < slices; k++) {
    const uint32_t k0 = b0 + (uint32_t)((uint64_t)(b1 - b0) * k / slices), k1 = b0 + (uint32_t)((uint64_t)(b1 - b0) * (k + 1) / slices);
    const uint64_t f0 = (uint64_t)k0 * spb, f1 = ((uint64_t)k1 * spb < s1) ? (uint64_t)k1 * spb : s1;
    const uint64_t o0 = AADF_FILE_HEADER_BYTES + (uint64_t)k0 * bs;
    uint64_t o1 = AADF_FILE_HEADER_BYTES + (uint64_t)k1 * bs;
    if (o1 > byte1) o1 = byte1;
    if (o1 > o0)
      CU(cudaMemcpyAsync(d_shard + (o0 - byte0), data + o0, (size_t)(o1 - o0), cudaMemcpyHostToDevice, gpu->s_in), "H2D aad slice");
    CU(cudaEventRecord(gpu->ev_in[k], gpu->s_in), "event");
    CU(cudaStreamWaitEvent(gpu->s_run, gpu->ev_in[k], 0), "wait");
    const AADApiResult e = launch_decode_wav_order(gpu, &geo, d_shard, byte0, byte1, ns, k0, k1, d_wav, s0, count, gpu->s_run);
    if (e != AAD_APIRESULT_OK) return e;
    CU(cudaEventRecord(gpu->ev_run[k], gpu->s_run), "event");
    CU(cudaStreamWaitEvent(gpu->s_out, gpu->ev_run[k], 0), "wait");
    if (f1 > f0)
      CU(cudaMemcpyAsync(interleaved + f0 * C, d_wav + (f0 - s0) * C, (size_t)(f1 - f0) * C * 2, cudaMemcpyDeviceToHost, gpu->s_out),
         "D2H wav slice");
  }
  const double t_queued = trace_on() ? wall_seconds() : 0.0;
  CU(cudaStreamSynchronize(gpu->s_out), "sync");
  if (trace_on())
    fprintf(stderr, "[aad_b200] decode range on device %d: %u slice(s), %.1f MB down: queued after %.3f ms, done after %.3f ms\n",
            gpu->device, slices, (double)count * C * 2 / 1e6, 1e3 * (t_queued - t_enter), 1e3 * (wall_seconds() - t_enter));
  return AAD_APIRESULT_OK;
}

static AADApiResult decode_stream_range(struct AADGpu *gpu, const struct AADHeaderInfo *h, const uint8_t *data,
                                        uint32_t data_size, uint32_t b0, uint32_t b1, int16_t *interleaved)
{
  WITH_CONTEXT_LOCK(gpu, decode_stream_range_unlocked(gpu, h, data, data_size, b0, b1, interleaved));
}

AADApiResult AADGpuGroup_DecodeInterleaved16(struct AADGpuGroup *g, const uint8_t *data, uint32_t data_size,
                                             int16_t *interleaved, uint32_t capacity_samples)
{
  if (!g || !data || !interleaved) return AAD_APIRESULT_INVALID_ARGUMENT;
  struct AADHeaderInfo h;
  const AADApiResult r = parse_stream_header(data, data_size, &h);
  if (r != AAD_APIRESULT_OK) return r;
  if (capacity_samples < h.num_samples) return AAD_APIRESULT_INSUFFICIENT_BUFFER;
  /* blocks the data reaches (src/aad_decoder.c:514-534); samples of later blocks are zero */
  AADApiResult tail;
  const uint32_t blocks = stream_block_span(&h, data_size, &tail);
  const uint64_t decoded = stream_decoded_samples(&h, blocks);
  if (decoded < h.num_samples)
    memset(interleaved + decoded * h.num_channels, 0, (size_t)(h.num_samples - decoded) * h.num_channels * 2);
  struct group_task tasks[AADGPU_MAX_GROUP];
  int n = 0;
  for (int d = 0; d < g->size; d++) {
    uint64_t b0, b1;
    split_range(blocks, g->size, d, &b0, &b1);
    if (b1 == b0) continue;
    struct group_task *t = &tasks[n++];
    memset(t, 0, sizeof(*t));
    t->op = GROUP_DECODE_STREAM;
    t->gpu = g->gpu[d];
    t->header = h;
    t->aad_in = data;
    t->data_size = data_size;
    t->block_begin = (uint32_t)b0;
    t->block_end = (uint32_t)b1;
    t->pcm_out = interleaved;
  }
  const AADApiResult e = n ? group_run(tasks, n) : AAD_APIRESULT_OK;
  return (e != AAD_APIRESULT_OK) ? e : tail;
}

/* ---- one stream ENCODED by a group: segment mode only ------------------------------------------- */

/* rows x width bytes between a host and a device buffer that share one layout (same pitch); rows that would run past
 * `limit` bytes of the layout are cut there */
static cudaError_t copy_slice_rows(uint8_t *dst, const uint8_t *src, uint64_t first, uint64_t pitch, uint64_t width, uint32_t rows,
                                   uint64_t limit, enum cudaMemcpyKind kind, cudaStream_t st)
{
  uint32_t whole = 0;          /* leading rows that lie completely below the limit */
  while (whole < rows && first + (uint64_t)whole * pitch + width <= limit) whole++;
  if (whole) {
    const cudaError_t e = copy_rows(dst + first, pitch, src + first, pitch, width, whole, kind, st);
    if (e != cudaSuccess) return e;
  }
  if (whole < rows) {
    const uint64_t at = first + (uint64_t)whole * pitch;
    if (at < limit) return cudaMemcpyAsync(dst + at, src + at, (size_t)(limit - at), kind, st);
  }
  return cudaSuccess;
}

/* Blocks [b0, b1) of one stream in SEGMENT mode (b0 on a segment boundary) from interleaved samples [b0*spb,
 * min(b1*spb, ns)): the shard copies only its own samples, encodes its segments as independent chains and writes its
 * own byte range of the caller's stream (the 31-byte file header with block 0).  Every segment is a chain of its own,
 * so the pipeline is cut ACROSS the segments: slice j = blocks [j*m, (j+1)*m) of EVERY segment of the shard, which
 * keeps every chain busy in every launch while slice j+1 is on its way up and slice j-1 on its way down (rows of the
 * 2-D copies = segments; the chain state crosses the launches in the device state array). */
static AADApiResult encode_stream_range_unlocked(struct AADGpu *gpu, const struct AADEncodeParameter *prm, uint32_t segment_blocks,
                                        const int16_t *interleaved, uint32_t num_samples, uint32_t b0, uint32_t b1,
                                        uint8_t *data)
{
  struct aadf_geometry geo;
  const AADApiResult r = check_encode_shape(prm, num_samples, &geo);
  if (r != AAD_APIRESULT_OK) return r;
  const uint32_t C = geo.channels, spb = geo.samples_per_block, bs = geo.block_size, ns = num_samples, SB = segment_blocks;
  const uint32_t nblk = aadf_num_blocks(ns, spb);
  const uint64_t s0 = (uint64_t)b0 * spb, s1 = ((uint64_t)b1 * spb < ns) ? (uint64_t)b1 * spb : ns;
  if (b1 <= b0 || s1 <= s0 || SB == 0 || b0 % SB != 0) return (b1 <= b0 || s1 <= s0) ? AAD_APIRESULT_OK : AAD_APIRESULT_INVALID_ARGUMENT;
  CU(cudaSetDevice(gpu->device), "cudaSetDevice");
  const uint64_t count = s1 - s0;
  const uint64_t pitch = round_up64(count, 64);
  const uint64_t byte0 = AADF_FILE_HEADER_BYTES + (uint64_t)b0 * bs;
  const uint64_t total = aadf_stream_bytes(ns, C, geo.bits, bs, spb);
  uint64_t byte1 = AADF_FILE_HEADER_BYTES + (uint64_t)b1 * bs;
  if (byte1 > total) byte1 = total;
  const uint64_t span = byte1 - byte0;
  const uint32_t nseg_all = (nblk + SB - 1) / SB;
  const uint32_t g0 = b0 / SB, g1 = (b1 + SB - 1) / SB, rows = g1 - g0;
  const size_t state_bytes = (size_t)nseg_all * C * AADK_STATE_WORDS * 4;
  if (!aadgpu_reserve(gpu, &gpu->wav, (size_t)C * count * 2)) return AAD_APIRESULT_NG;
  if (!aadgpu_reserve(gpu, &gpu->pcm, (size_t)C * pitch * 2)) return AAD_APIRESULT_NG;
  if (!aadgpu_reserve(gpu, &gpu->aad, (size_t)span + (size_t)bs + 256)) return AAD_APIRESULT_NG;
  if (!aadgpu_reserve(gpu, &gpu->state, state_bytes)) return AAD_APIRESULT_NG;
  cudaStream_t s = gpu->s_run;
  CU(cudaMemsetAsync(gpu->aad.ptr, 0, (size_t)span + (size_t)bs + 256, s), "memset aad");
  CU(cudaMemsetAsync(gpu->state.ptr, 0, state_bytes, s), "memset state");
  /* the shard's first block 32-byte aligned at ptr + 32, the file header (shard 0 only) right in front of it; the
   * kernel is told which byte of the stream and which sample of the rows its buffers start at */
  uint8_t *d_block0 = (uint8_t *)gpu->aad.ptr + 32;
  const uint64_t lead = (b0 == 0) ? AADF_FILE_HEADER_BYTES : 0;
  struct aadk_encode_params p;
  memset(&p, 0, sizeof(p));
  p.pcm = (const int16_t *)gpu->pcm.ptr;
  p.sample_base = s0;
  p.pcm_clip_stride = 0;
  p.pcm_ch_stride = pitch;
  p.uniform_samples = ns;
  p.num_streams = 1;
  p.geo = geo;
  p.sampling_rate = prm->sampling_rate;
  p.trials = prm->num_encode_trials;
  p.aad = d_block0 - lead;
  p.byte_base = byte0 - lead;
  p.aad_stride = 0;
  p.state_in = (const int32_t *)gpu->state.ptr;
  p.state_out = (int32_t *)gpu->state.ptr;
  p.segment_blocks = SB;
  p.num_segments = nseg_all;
  p.segment_relative = 1;
  p.segment_begin = g0;
  p.segment_end = g1;

  /* slices across the segments: about 8 MiB of samples each, at most one per block of a segment */
  uint32_t slices = pick_stream_slices(count * C * 2, SB);
  if (slices > 8) slices = 8;                                      /* rows of the 2-D copies stay tens of KiB long */
  const uint32_t m = (SB + slices - 1) / slices;
  slices = (SB + m - 1) / m;
  const uint64_t frame = (uint64_t)C * 2;                          /* bytes per interleaved frame */
  const uint8_t *h_wav = (const uint8_t *)(interleaved + s0 * C);  /* the shard's samples: frame 0 = sample s0 */
  uint8_t *d_wav = (uint8_t *)gpu->wav.ptr;
  uint8_t *h_aad = data + byte0;                                   /* the shard's bytes: offset 0 = block b0 */
  for (uint32_t j = 0; j < slices; j++) {
    const uint32_t r0 = j * m, r1 = (r0 + m < SB) ? r0 + m : SB;   /* blocks [r0, r1) of every segment */
    CU(copy_slice_rows(d_wav, h_wav, (uint64_t)r0 * spb * frame, (uint64_t)SB * spb * frame, (uint64_t)(r1 - r0) * spb * frame, rows,
                       count * frame, cudaMemcpyHostToDevice, gpu->s_in), "H2D wav slice");
    CU(cudaEventRecord(gpu->ev_in[j], gpu->s_in), "event");
    CU(cudaStreamWaitEvent(s, gpu->ev_in[j], 0), "wait");
    CU((cudaError_t)aadk_launch_deinterleave16_rows((const int16_t *)d_wav, (int16_t *)gpu->pcm.ptr, pitch, C, (uint64_t)SB * spb,
                                                    (uint64_t)r0 * spb, (r1 - r0) * spb, rows, count, s), "deinterleave kernel launch");
    p.block_begin = r0;
    p.block_end = r1;
    CU((cudaError_t)aadk_launch_encode(&p, s), "encode kernel launch");
    CU(cudaEventRecord(gpu->ev_run[j], s), "event");
    CU(cudaStreamWaitEvent(gpu->s_out, gpu->ev_run[j], 0), "wait");
    if (j == 0 && lead)
      CU(cudaMemcpyAsync(data, d_block0 - lead, (size_t)lead, cudaMemcpyDeviceToHost, gpu->s_out), "D2H file header");
    CU(copy_slice_rows(h_aad, d_block0, (uint64_t)r0 * bs, (uint64_t)SB * bs, (uint64_t)(r1 - r0) * bs, rows, span,
                       cudaMemcpyDeviceToHost, gpu->s_out), "D2H aad slice");
  }
  CU(cudaStreamSynchronize(gpu->s_out), "sync");
  return AAD_APIRESULT_OK;
}

static AADApiResult encode_stream_range(struct AADGpu *gpu, const struct AADEncodeParameter *prm, uint32_t segment_blocks,
                                        const int16_t *interleaved, uint32_t num_samples, uint32_t b0, uint32_t b1,
                                        uint8_t *data)
{
  WITH_CONTEXT_LOCK(gpu, encode_stream_range_unlocked(gpu, prm, segment_blocks, interleaved, num_samples, b0, b1, data));
}

AADApiResult AADGpuGroup_EncodeInterleaved16(struct AADGpuGroup *g, const struct AADEncodeParameter *prm,
                                             uint32_t segment_blocks, const int16_t *interleaved, uint32_t num_samples,
                                             uint8_t *data, uint32_t data_size, uint32_t *output_size)
{
  if (!g || !prm || !interleaved || !data || !output_size) return AAD_APIRESULT_INVALID_ARGUMENT;
  if (segment_blocks == 0) {
    aadgpu_set_error("one stream cannot be encoded in shards bit-exactly (the reference carries the chain state through the "
                     "whole stream, src/aad_encoder.c:853-886): pass segment_blocks > 0 (AADGpu_SetEncodeSegmentBlocks)");
    return AAD_APIRESULT_INVALID_ARGUMENT;
  }
  struct aadf_geometry geo;
  const AADApiResult r = check_encode_shape(prm, num_samples, &geo);
  if (r != AAD_APIRESULT_OK) return r;
  if (data_size < AADF_FILE_HEADER_BYTES) return AAD_APIRESULT_INSUFFICIENT_DATA;
  const uint64_t bytes = aadf_stream_bytes(num_samples, geo.channels, geo.bits, geo.block_size, geo.samples_per_block);
  if (bytes > data_size) return AAD_APIRESULT_INSUFFICIENT_BUFFER;
  const uint32_t nblk = aadf_num_blocks(num_samples, geo.samples_per_block);
  const uint64_t nseg = ((uint64_t)nblk + segment_blocks - 1) / segment_blocks;
  struct group_task tasks[AADGPU_MAX_GROUP];
  int n = 0;
  for (int d = 0; d < g->size; d++) {
    uint64_t g0, g1;
    split_range(nseg, g->size, d, &g0, &g1);
    if (g1 == g0) continue;
    struct group_task *t = &tasks[n++];
    memset(t, 0, sizeof(*t));
    t->op = GROUP_ENCODE_STREAM;
    t->gpu = g->gpu[d];
    t->param = *prm;
    t->segment_blocks = segment_blocks;
    t->pcm_in = interleaved;
    t->num_samples = num_samples;
    t->block_begin = (uint32_t)(g0 * segment_blocks);
    t->block_end = (uint32_t)((g1 * segment_blocks < nblk) ? g1 * segment_blocks : nblk);
    t->aad_out = data;
  }
  const AADApiResult e = n ? group_run(tasks, n) : AAD_APIRESULT_OK;
  if (e == AAD_APIRESULT_OK) *output_size = (uint32_t)bytes;
  return e;
}
