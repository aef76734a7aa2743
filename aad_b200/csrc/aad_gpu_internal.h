/*
 * aad_gpu_internal.h -- private to the host C sources of libaad_b200.so.
 */
#ifndef AAD_GPU_INTERNAL_H
#define AAD_GPU_INTERNAL_H

#include <stddef.h>
#include <stdint.h>
#include <pthread.h>
#include <sched.h>
#include <cuda_runtime_api.h>

#include "aad_b200.h"
#include "aad_format.h"
#include "aad_kernels.h"

#define AADGPU_PIPE_STREAMS 3   /* H2D, kernels, D2H */
#define AADGPU_MAX_GROUP 16     /* devices in one AADGpuGroup */
#define AADGPU_MAX_SLICES 64    /* block-range slices a host pipeline cuts its copies into */

struct aadgpu_buffer {
  void *ptr;
  size_t cap;
};

/* One context = one device + its streams and grow-only scratch buffers.  Every entry point that uses the scratch
 * holds `lock` for its duration, so a context may be shared by threads (the drop-in AADEncoder_* / AADDecoder_*
 * handles of all threads share the process-wide default context: distinct handles stay independent, as in the
 * reference, their GPU work is serialised). */
struct AADGpu {
  int device;
  pthread_mutex_t lock;
  cudaStream_t s_in, s_run, s_out;
  cudaStream_t s_out2;         /* second device -> host queue (the batch round trip sends .aad and PCM side by side) */
  cudaEvent_t ev_in[AADGPU_MAX_SLICES], ev_run[AADGPU_MAX_SLICES], ev_out[AADGPU_MAX_SLICES];
  void *ring_in[3], *ring_out[3];   /* pinned bounce buffers of the drop-in paths (caller memory is pageable), lazily allocated */
  struct aadgpu_buffer pcm, aad, state, lens, sizes, lut, wav, pcm2, raw, stats;
  int lut_ready;
  int cpus_known, cpus_count;       /* AADGpu_BindHostThread: the device's local CPU list, read from sysfs once */
  cpu_set_t cpus;
  uint32_t segment_blocks;   /* AADGpu_SetEncodeSegmentBlocks; 0 = the reference's whole-stream state carry */
};

/* error plumbing: records a message for AADGpu_LastError() and returns AAD_APIRESULT_NG */
AADApiResult aadgpu_fail(const char *what, cudaError_t err);
void aadgpu_set_error(const char *msg);

/* grow-only device scratch */
int aadgpu_reserve(struct AADGpu *gpu, struct aadgpu_buffer *b, size_t bytes);

/* process-wide default context used by the drop-in handles (created on first use on the
 * device named by $AAD_B200_DEVICE, default 0); NULL when CUDA is unusable */
struct AADGpu *aadgpu_default(void);

uint32_t aadgpu_max_channels(void);

/* AADDecoder_DecodeHeader's result must also pass this before a kernel sees it (aad_decoder.c) */
AADApiResult aaddec_check_header(const struct AADHeaderInfo *h);

/* Single-stream paths behind AADEncoder_EncodeWhole / AADDecoder_DecodeWhole / _DecodeBlock.
 * Host pointers, int32 PCM (the reference API type). */
AADApiResult aadgpu_encode_stream_i32(struct AADGpu *gpu, const struct aadf_geometry *geo, uint32_t sampling_rate,
                                      uint32_t trials, const int32_t *const *input, uint32_t num_samples,
                                      int32_t *state /* [channels][AADK_STATE_WORDS] in/out */, uint8_t *data,
                                      uint32_t *output_size);
/* Decodes blocks [0, num_blocks) found at data + 31 (data_size counts from data[0]);
 * num_samples / buf_samples as in src/aad_decoder.c:478-538. */
AADApiResult aadgpu_decode_stream_i32(struct AADGpu *gpu, const struct aadf_geometry *geo, const uint8_t *data,
                                      uint32_t data_size, uint32_t num_blocks, uint32_t num_samples,
                                      uint32_t buf_samples, int32_t *const *buffer);

#endif
