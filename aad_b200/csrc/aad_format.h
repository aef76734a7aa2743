/*
 * aad_format.h -- the AAD wire format as plain inline arithmetic, shared by the C host
 * code and the CUDA kernels (no state, no I/O).
 *
 * Format facts and where the reference defines them:
 *   file header   31 bytes, big-endian fields            src/aad_encoder.c:190-214
 *   block         C channel headers of 18 bytes each     src/aad_encoder.c:619-655
 *                   u16  (stepsize_index << 4) | shift
 *                   4 x { u16 weight >> shift, u16 history }
 *                 then code groups, channel-interleaved  src/aad_encoder.c:661-722
 *                   4-bit: 2 codes / 1 byte, 3-bit: 8 codes / 3 bytes, 2-bit: 4 codes / 1 byte
 *   a block holds 4 samples in its header + the samples of its groups
 */
#ifndef AAD_FORMAT_H
#define AAD_FORMAT_H

#include <stdint.h>

#if defined(__CUDACC__)
#define AADF_INLINE __host__ __device__ __forceinline__
#else
#define AADF_INLINE static inline
#endif

#define AADF_TAPS                4     /* src/aad_internal.h:10 */
#define AADF_FILE_HEADER_BYTES   31    /* src/aad.h:22 */
#define AADF_CHANNEL_HEADER_BYTES 18   /* src/aad_internal.h:37 */
#define AADF_INDEX_MAX           4080  /* (256 - 1) << 4, src/aad_tables.h:38-39 */
#define AADF_MAX_CHANNELS        8

/* bytes per interleave group per channel = lcm(8, bits) / 8   (src/aad_encoder.c:111) */
AADF_INLINE uint32_t aadf_group_bytes(uint32_t bits) { return (bits == 3u) ? 3u : 1u; }
/* samples per group per channel                               (src/aad_encoder.c:112) */
AADF_INLINE uint32_t aadf_group_samples(uint32_t bits)
{
  return (bits == 4u) ? 2u : ((bits == 3u) ? 8u : ((bits == 2u) ? 4u : 8u));
}

/* Block geometry for a stream: everything the kernels need besides pointers. */
struct aadf_geometry {
  uint32_t channels;
  uint32_t bits;
  uint32_t block_size;          /* bytes per full block */
  uint32_t samples_per_block;   /* per channel */
  uint32_t ms;                  /* 1 = mid/side on channels 0,1 */
};

/* src/aad_encoder.c:85-131; returns 0 when the parameters are rejected there. */
AADF_INLINE int aadf_block_geometry(uint32_t max_block_size, uint32_t channels, uint32_t bits, uint32_t max_channels,
                                    uint32_t *block_size, uint32_t *samples_per_block)
{
  if (channels == 0u || channels > max_channels || bits == 0u || bits > 4u) return 0;
  const uint32_t hdr = AADF_CHANNEL_HEADER_BYTES * channels;
  if (max_block_size < hdr) return 0;
  const uint32_t unit = channels * aadf_group_bytes(bits);
  const uint32_t units = (max_block_size - hdr) / unit;
  *block_size = (hdr + units * unit) & 0xFFFFu;
  *samples_per_block = units * aadf_group_samples(bits) + AADF_TAPS;
  return 1;
}

/* number of blocks the encoder emits for n samples per channel (src/aad_encoder.c:853-886) */
AADF_INLINE uint32_t aadf_num_blocks(uint32_t num_samples, uint32_t samples_per_block)
{
  return (uint32_t)(((uint64_t)num_samples + samples_per_block - 1u) / samples_per_block);
}

/* bytes of one block holding n (<= samples_per_block) samples per channel: header plus whole
 * groups, the last one zero padded (src/aad_encoder.c:592-593,663,678,704) */
AADF_INLINE uint32_t aadf_block_bytes(uint32_t n, uint32_t channels, uint32_t bits)
{
  const uint32_t gs = aadf_group_samples(bits);
  const uint32_t groups = (n > AADF_TAPS) ? (n - AADF_TAPS + gs - 1u) / gs : 0u;
  return channels * (AADF_CHANNEL_HEADER_BYTES + groups * aadf_group_bytes(bits));
}

/* exact size of a whole encoded stream */
AADF_INLINE uint64_t aadf_stream_bytes(uint32_t num_samples, uint32_t channels, uint32_t bits, uint32_t block_size,
                                       uint32_t samples_per_block)
{
  const uint64_t full = num_samples / samples_per_block;
  const uint32_t tail = num_samples % samples_per_block;
  return AADF_FILE_HEADER_BYTES + full * block_size + (tail ? aadf_block_bytes(tail, channels, bits) : 0u);
}

/* worst case (every block full) -- what callers should allocate per stream */
AADF_INLINE uint64_t aadf_stream_bytes_bound(uint32_t num_samples, uint32_t block_size, uint32_t samples_per_block)
{
  return AADF_FILE_HEADER_BYTES + (uint64_t)aadf_num_blocks(num_samples, samples_per_block) * block_size;
}

AADF_INLINE void aadf_put_be16(uint8_t *p, uint32_t v) { p[0] = (uint8_t)(v >> 8); p[1] = (uint8_t)v; }
AADF_INLINE void aadf_put_be32(uint8_t *p, uint32_t v)
{
  p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v;
}
AADF_INLINE uint32_t aadf_get_be16(const uint8_t *p) { return ((uint32_t)p[0] << 8) | p[1]; }
AADF_INLINE uint32_t aadf_get_be32(const uint8_t *p)
{
  return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}

/* 31-byte stream header, src/aad_encoder.c:190-214.  Versions are written from the build
 * constants, never from caller data (src/aad_encoder.c:195-200). */
AADF_INLINE void aadf_write_file_header(uint8_t *p, uint32_t channels, uint32_t num_samples, uint32_t sampling_rate,
                                        uint32_t bits, uint32_t block_size, uint32_t samples_per_block, uint32_t ms)
{
  p[0] = 'A'; p[1] = 'A'; p[2] = 'D'; p[3] = 0;
  aadf_put_be32(p + 4, 4u);    /* AAD_FORMAT_VERSION */
  aadf_put_be32(p + 8, 18u);   /* AAD_CODEC_VERSION  */
  aadf_put_be16(p + 12, channels);
  aadf_put_be32(p + 14, num_samples);
  aadf_put_be32(p + 18, sampling_rate);
  aadf_put_be16(p + 22, bits);
  aadf_put_be16(p + 24, block_size);
  aadf_put_be32(p + 26, samples_per_block);
  p[30] = (uint8_t)ms;
}

#endif /* AAD_FORMAT_H */
