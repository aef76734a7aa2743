/*
 * aad_b200.h -- extended C ABI of libaad_b200.so: the same ADPCM hot path as the drop-in
 * API (aad_encoder.h / aad_decoder.h), exposed in the shapes a GPU wants:
 *
 *   - batches of independent streams ("clips") in one call   (BASELINE config 5)
 *   - int16 PCM at the boundary (what a 16-bit WAV holds) instead of int32
 *   - device-resident entry points (no host copies) for callers that already live on the GPU
 *   - host entry points that pipeline H2D / kernels / D2H in chunks over pinned staging
 *
 * There is no CPU fallback: every entry point returns AAD_APIRESULT_NG (and
 * AADGpu_LastError() says why) when no CUDA device is usable.
 *
 * Reference interfaces this extends: AADEncoder_EncodeWhole (src/aad_encoder.h:47-50),
 * AADDecoder_DecodeWhole (src/aad_decoder.h:39-42); parameter meaning follows
 * struct AADEncodeParameter (src/aad_encoder.h:8-15) and struct AADHeaderInfo
 * (src/aad.h:43-53).
 */
#ifndef AAD_B200_H
#define AAD_B200_H

#include <stddef.h>
#include <stdint.h>
#include "aad.h"
#include "aad_encoder.h"

#ifdef __cplusplus
extern "C" {
#endif

struct AADGpu;   /* one CUDA device: streams, staging buffers, scratch */

/* Shape of a batch of streams.  Stream i's planar int16 PCM starts at
 * pcm + i*pcm_stream_stride; its channel c at + c*pcm_channel_stride (both in samples).
 * Stream i's encoded bytes (31-byte file header included) start at aad + i*aad_stream_stride. */
struct AADGpuBatch {
  uint32_t num_streams;
  uint32_t num_samples;                       /* per channel; the maximum when lengths are ragged */
  struct AADEncodeParameter param;            /* channels, rate, bits, max block size, MS, trials */
  uint64_t pcm_stream_stride;                 /* samples */
  uint64_t pcm_channel_stride;                /* samples */
  uint64_t aad_stream_stride;                 /* bytes, >= AADGpu_StreamBytesBound() */
};

/* ---- device / context ------------------------------------------------------------------ */
int  AADGpu_DeviceCount(void);                           /* 0 when CUDA is unusable */
struct AADGpu *AADGpu_Create(int device);                /* NULL on failure */
void AADGpu_Destroy(struct AADGpu *gpu);
const char *AADGpu_LastError(void);                      /* thread-local, never NULL */
uint64_t AADGpu_KernelLaunchCount(void);                 /* kernels launched by this library so far */
uint64_t AADGpu_TmaLaunchCount(void);                    /* of those, launches of the tensor-map staged decoder (kernel path 7) */
void AADGpu_SetMaxChannels(uint32_t max_channels);       /* 2 = stock reference limit, 8 = default */
uint32_t AADGpu_GetMaxChannels(void);
/* 0 (default): fast kernels wherever the shape allows (mono / stereo: aad_decode_fast, 3..8 channels:
 * aad_decode_wide), generic kernels otherwise; 1: always the generic (any channel count / alignment)
 * kernels; 2: like 0, but mono / stereo streams are decoded by aad_decode_wide too; 4: like 0, but mono 4-bit
 * streams leave shared memory through the TMA unit (cp.async.bulk); 5 / 6: like 0, but the warp tasks of aad_decode_fast
 * always / never run on from one stream into the next (default: where per-stream tasks would idle 1 lane in 16 or
 * more); 7: like 0, but mono 4-bit / 2-bit streams whose blocks are 16-byte aligned in device memory (stream stride and
 * block size multiples of 16, no per-stream size array) are staged by the TMA unit through a tensor map
 * (aad_decode_tma, cp.async.bulk.tensor + mbarrier); 8: like 7 with 24 instead of 16 warps per SM (output rows of 64
 * samples, flushed twice per window).  All bit-exact; this exists for testing and measurement. */
void AADGpu_SetKernelPath(int path);
/* 1 (default): with few chains the encoder runs the two independent dry passes of a block interleaved
 * in one thread; 0: never.  Same bytes out; this exists for testing and measurement. */
void AADGpu_SetEncoderPairing(int on);
/* The fast encoder's pass schedule, for tests and measurement: 1 (default) = chosen by the launch's shape; 0 = one pass
 * at a time in every thread; 2 = the two independent dry passes of a block interleaved in one thread (few chains);
 * 3 = helper lanes of every warp run the baseline passes (chains <= 25 per warp scheduler: 5 pass slots per block
 * instead of 6); 4 = helper lanes plus a second warp running the emitting passes ahead of the decision (a handful of
 * chains, e.g. ONE long stream: 4 slots).  A forced schedule applies only where the launch qualifies.  Same bytes. */
void AADGpu_SetEncoderSchedule(int mode);

/* Segment-parallel encoding -- an EXTENSION, NOT byte-identical to the reference encoder.
 * The reference carries the predictor weights and the step index from block to block through the whole stream
 * (src/aad_encoder.c:21,853-886), which makes one stream a serial chain: a 1-hour stereo file is 2 GPU threads.
 * With blocks > 0 every run of `blocks` consecutive blocks is encoded as if it were a stream of its own (zero
 * weights and step index at its first block, no previous-block trial pass there: exactly what the reference does
 * at the start of a stream), so a stream becomes ceil(num_blocks / blocks) independent chains per channel.
 * The result is a valid .aad stream: every block header carries the chain state the decoder needs, so the stock
 * decoder (and this library's) decodes it; each segment's bytes equal the reference encoder's output for that
 * run of samples on a fresh handle.  Applies to AADGpu_EncodeBatch[Device], AADGpu_ReconstructBatch,
 * AADGpu_EncodeInterleaved16 and AADGpu_ReconstructInterleaved16 on this context; never to the drop-in
 * AADEncoder_EncodeWhole, which stays bit-exact with the reference.  0 (default) = off. */
AADApiResult AADGpu_SetEncodeSegmentBlocks(struct AADGpu *gpu, uint32_t blocks);
uint32_t AADGpu_GetEncodeSegmentBlocks(const struct AADGpu *gpu);

/* Bind the calling host thread to the CPUs next to the device (the NUMA node of its PCIe root, from
 * sysfs).  Buffers the thread pins afterwards live on that node, so one device's copies do not cross
 * the socket interconnect.  Returns the number of CPUs in the set, 0 when the topology is unknown
 * (nothing changed).  The device-group calls do this for their own worker threads. */
int AADGpu_BindHostThread(struct AADGpu *gpu);

/* Measure this device's host link with plain 1-D copies of `bytes` between pinned host memory and HBM, `repeats`
 * times each: gbs[0] = host -> device alone, gbs[1] = device -> host alone, gbs[2] = both directions at once
 * (aggregate), in GB/s.  What the pipelined entry points below can at best reach; bench.py reports it beside
 * their end-to-end figures. */
AADApiResult AADGpu_LinkProbe(struct AADGpu *gpu, size_t bytes, int repeats, double gbs[3]);
/* the same for `rows` rows of `width` bytes that lie `host_pitch` bytes apart in pinned host memory (packed on the
 * device): the shape of the copies the batch pipelines issue -- a block-range slice of every stream of a batch */
AADApiResult AADGpu_LinkProbeRows(struct AADGpu *gpu, size_t rows, size_t width, size_t host_pitch, int repeats, double gbs[3]);

/* pinned host memory for the host entry points (plain malloc'd memory works too, slower) */
void *AADGpu_HostAlloc(size_t bytes);
void  AADGpu_HostFree(void *p);

/* ---- sizing helpers (pure arithmetic) -------------------------------------------------- */
/* worst-case encoded bytes of one stream (file header + every block full) */
uint64_t AADGpu_StreamBytesBound(const struct AADEncodeParameter *param, uint32_t num_samples);
/* exact encoded bytes of one stream */
uint64_t AADGpu_StreamBytes(const struct AADEncodeParameter *param, uint32_t num_samples);

/* ---- device-resident hot path (all pointers are device pointers on gpu's device) -------- */
/* num_samples_dev / sizes_dev: per-stream lengths, NULL = uniform (batch->num_samples / full
 * streams).  out_sizes_dev nullable.  `stream` is a cudaStream_t (NULL = default stream); the
 * calls only enqueue work. */
AADApiResult AADGpu_EncodeBatchDevice(struct AADGpu *gpu, const struct AADGpuBatch *batch,
                                      const int16_t *pcm_dev, const uint32_t *num_samples_dev,
                                      uint8_t *aad_dev, uint32_t *out_sizes_dev, void *stream);
/* Each stream's length is taken from its own file header; sizes_dev (nullable) bounds the
 * valid bytes per stream.  param.num_encode_trials is ignored. */
AADApiResult AADGpu_DecodeBatchDevice(struct AADGpu *gpu, const struct AADGpuBatch *batch,
                                      const uint8_t *aad_dev, const uint32_t *sizes_dev,
                                      int16_t *pcm_dev, void *stream);

/* ---- host entry points: H2D -> kernels -> D2H, chunked and overlapped ------------------- */
/* num_samples / sizes: host arrays or NULL (uniform).  out_sizes: host array or NULL. */
AADApiResult AADGpu_EncodeBatch(struct AADGpu *gpu, const struct AADGpuBatch *batch,
                                const int16_t *pcm, const uint32_t *num_samples,
                                uint8_t *aad, uint32_t *out_sizes);
AADApiResult AADGpu_DecodeBatch(struct AADGpu *gpu, const struct AADGpuBatch *batch,
                                const uint8_t *aad, const uint32_t *sizes, int16_t *pcm);

/* Encode the batch and decode it back in ONE pass over the data -- src/main.c:275-346
 * (execute_reconstruction_core) for a whole batch: per slice H2D pcm | encode | decode | D2H .aad and
 * D2H reconstruction, both directions of the link busy at once; the encoded streams stay in HBM between
 * the two kernels.  aad / out_sizes may be NULL when only the reconstruction is wanted. */
AADApiResult AADGpu_ReconstructBatch(struct AADGpu *gpu, const struct AADGpuBatch *batch,
                                     const int16_t *pcm, const uint32_t *num_samples,
                                     uint8_t *aad, uint32_t *out_sizes, int16_t *reconstructed);

/* Measurement: exactly the host <-> device copies AADGpu_ReconstructBatch issues for this batch and these buffers
 * (same slices, row shapes and streams), without the kernels and without any dependency between the copies -- the
 * time the end-to-end call cannot beat on this box.  Call it right after AADGpu_ReconstructBatch on the same
 * arguments: the device buffers still hold that call's results, so `aad` and `reconstructed` receive the same bytes. */
AADApiResult AADGpu_CopyProbeBatch(struct AADGpu *gpu, const struct AADGpuBatch *batch, const int16_t *pcm, uint8_t *aad,
                                   int16_t *reconstructed);

/* ---- one stream in WAV order (interleaved int16), host pointers: what `aad -e / -d / -r` call ---- */
/* src/main.c:141-227 (execute_encode) without the host-side int32 shuffle: the 16-bit samples of a
 * WAV data chunk are copied as they are and de-interleaved on the device. */
AADApiResult AADGpu_EncodeInterleaved16(struct AADGpu *gpu, const struct AADEncodeParameter *param,
                                        const int16_t *interleaved, uint32_t num_samples,
                                        uint8_t *data, uint32_t data_size, uint32_t *output_size);
/* src/main.c:61-138 (execute_decode): the stream's own header says how many channels / samples come
 * out; capacity_samples (per channel) must cover it. */
AADApiResult AADGpu_DecodeInterleaved16(struct AADGpu *gpu, const uint8_t *data, uint32_t data_size,
                                        int16_t *interleaved, uint32_t capacity_samples);
/* src/main.c:275-346 (execute_reconstruction_core): encode then decode, the stream stays in HBM.
 * encoded_size (nullable) receives the size the .aad file would have had. */
AADApiResult AADGpu_ReconstructInterleaved16(struct AADGpu *gpu, const struct AADEncodeParameter *param,
                                             const int16_t *interleaved, uint32_t num_samples,
                                             int16_t *reconstructed, uint32_t *encoded_size);

/* ---- several GPUs of one box ------------------------------------------------------------------- */
/* One context and one host thread per device.  Batches shard by stream; one long stream shards its
 * DECODE by block range (every block header reloads the chain state, src/aad_decoder.c:364-380).
 * Every device writes its own disjoint slice of the caller's buffers: no collective, no peer copy.
 * There is no sharded encode of ONE stream: the reference carries state from block to block
 * (src/aad_encoder.c:853-886), so its bytes cannot be produced in independent pieces. */
struct AADGpuGroup;
/* devices == NULL: the first num_devices visible devices (all of them when num_devices <= 0) */
struct AADGpuGroup *AADGpuGroup_Create(const int *devices, int num_devices);
void AADGpuGroup_Destroy(struct AADGpuGroup *group);
int AADGpuGroup_Size(const struct AADGpuGroup *group);
struct AADGpu *AADGpuGroup_Device(const struct AADGpuGroup *group, int index);
/* same arguments and results as AADGpu_EncodeBatch / AADGpu_DecodeBatch / AADGpu_DecodeInterleaved16 */
AADApiResult AADGpuGroup_EncodeBatch(struct AADGpuGroup *group, const struct AADGpuBatch *batch,
                                     const int16_t *pcm, const uint32_t *num_samples,
                                     uint8_t *aad, uint32_t *out_sizes);
AADApiResult AADGpuGroup_DecodeBatch(struct AADGpuGroup *group, const struct AADGpuBatch *batch,
                                     const uint8_t *aad, const uint32_t *sizes, int16_t *pcm);
AADApiResult AADGpuGroup_DecodeInterleaved16(struct AADGpuGroup *group, const uint8_t *data, uint32_t data_size,
                                             int16_t *interleaved, uint32_t capacity_samples);
/* ONE stream encoded by the whole group -- segment mode only (see AADGpu_SetEncodeSegmentBlocks: an extension,
 * not byte-identical to the reference encoder): the segments are shared out in contiguous ranges, each device
 * copies only its own samples and writes its own byte range of `data`.  Same bytes as AADGpu_EncodeInterleaved16
 * on one device with the same segment_blocks.  segment_blocks == 0 is refused with AAD_APIRESULT_INVALID_ARGUMENT:
 * without segments a stream is a serial chain (src/aad_encoder.c:853-886) and does not shard. */
AADApiResult AADGpuGroup_EncodeInterleaved16(struct AADGpuGroup *group, const struct AADEncodeParameter *param,
                                             uint32_t segment_blocks, const int16_t *interleaved, uint32_t num_samples,
                                             uint8_t *data, uint32_t data_size, uint32_t *output_size);

/* AADGpu_EncodeInterleaved16 for a WAV data chunk of any supported depth (8 / 16 / 24 / 32 bits per sample, as it
 * lies in the file): the codec sees the top 16 bits of every sample, (int16_t)(PCM >> 16) of src/main.c:175-179 with
 * the widening of src/wav.c:391-415, computed on the device while de-interleaving. */
AADApiResult AADGpu_EncodeWav(struct AADGpu *gpu, const struct AADEncodeParameter *param, const uint8_t *wav_data,
                              uint32_t wav_bits_per_sample, uint32_t num_samples, uint8_t *data, uint32_t data_size,
                              uint32_t *output_size);

/* ---- the command line's analysis modes on the device, src/main.c:275-503 ------------------- */
enum AADGpuAnalysis {
  AADGPU_ANALYSIS_RECONSTRUCT = 0,   /* -r: encode -> decode, written in the input's sample format (src/main.c:372-381) */
  AADGPU_ANALYSIS_RESIDUAL = 1,      /* -g: input minus reconstruction, 32-bit wrapping (src/main.c:418-428) */
  AADGPU_ANALYSIS_STATISTICS = 2     /* -c: RMSE, MSD, MaxAE of src/main.c:470-497 */
};
/* wav_data: the data chunk of a PCM WAV file as it lies in the file (wav_bits_per_sample 8 / 16 / 24 / 32,
 * num_samples per channel, interleaved).  The chunk is uploaded once; narrowing to the codec's 16 bits, the round
 * trip, and the per-sample arithmetic of the mode all run on the device.  RECONSTRUCT / RESIDUAL write out_data
 * (same size and format as wav_data); STATISTICS writes stats[0..2] = RMSE, MSD, MaxAE.  The two sums of STATISTICS
 * are reduced in a fixed parallel order, not sample by sample as the reference does: they agree with the
 * reference's to ~1e-15 relative (invisible at the command line's %f), MaxAE exactly.  encoded_size is optional. */
AADApiResult AADGpu_AnalyzeWav(struct AADGpu *gpu, const struct AADEncodeParameter *param, const uint8_t *wav_data,
                               uint32_t wav_bits_per_sample, uint32_t num_samples, enum AADGpuAnalysis what,
                               uint8_t *out_data, double stats[3], uint32_t *encoded_size);

/* ---- deterministic synthetic PCM (bench / tests), SURVEY.md 8(d) ------------------------ */
AADApiResult AADGpu_SynthBatchDevice(struct AADGpu *gpu, const struct AADGpuBatch *batch,
                                     uint32_t first_stream, int16_t *pcm_dev, void *stream);
/* the generator's 1024-entry sine table (host arithmetic), so a host mirror can reproduce the batch */
void AADGpu_SynthLut(int16_t lut[1024]);

/* ---- WAV-order helpers on the device (src/main.c:122-126,175-179) ----------------------- */
AADApiResult AADGpu_Deinterleave16Device(struct AADGpu *gpu, const int16_t *interleaved_dev, int16_t *planar_dev,
                                         uint64_t channel_stride, uint32_t channels, uint32_t num_samples,
                                         void *stream);
AADApiResult AADGpu_Interleave16Device(struct AADGpu *gpu, const int16_t *planar_dev, uint64_t channel_stride,
                                       int16_t *interleaved_dev, uint32_t channels, uint32_t num_samples,
                                       void *stream);

#ifdef __cplusplus
}
#endif
#endif /* AAD_B200_H */
