"""numpy mirror of the device synthetic-PCM generator (aad_kernels.cu: aad_synth), bit for bit.

x[n] = 0.4*sin(2*pi*f1*n/fs) + 0.2*sin(2*pi*3001*n/fs) + uniform noise in [-1000, 1000], all in
integer arithmetic on a 1024-entry int16 sine table, so host and device
agree exactly and the CPU baseline runs on the very same samples as the GPU.
f1 = 440*(1+channel) + 7*(stream % 97) Hz.
"""
import numpy as np


def _mix32(x):
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x85EBCA6B)
    x ^= x >> np.uint32(13)
    x *= np.uint32(0xC2B2AE35)
    x ^= x >> np.uint32(16)
    return x


def synth_pcm16(lut, first_stream, num_streams, channels, num_samples, sampling_rate):
    """Returns int16 [num_streams, channels, num_samples]; lut from GpuApi.synth_lut()."""
    lut = np.asarray(lut, dtype=np.int32)
    n = np.arange(num_samples, dtype=np.uint32)
    out = np.empty((num_streams, channels, num_samples), dtype=np.int16)
    p2 = np.uint32(((3001 << 32) // sampling_rate) & 0xFFFFFFFF)
    s2 = lut[(n * p2) >> np.uint32(22)]
    with np.errstate(over="ignore"):
        for i in range(num_streams):
            gi = (first_stream + i) & 0xFFFFFFFF
            for c in range(channels):
                f1 = 440 * (1 + c) + 7 * (gi % 97)
                p1 = np.uint32(((f1 << 32) // sampling_rate) & 0xFFFFFFFF)
                seed = np.uint32(0x9E3779B9 ^ ((gi * 2654435761 + c * 40503 + 1) & 0xFFFFFFFF))
                s1 = lut[(n * p1) >> np.uint32(22)]
                noise = (_mix32(seed ^ (n * np.uint32(0x9E3779B1))) % np.uint32(2001)).astype(np.int32) - 1000
                x = ((s1 * 13107) >> 15) + ((s2 * 6553) >> 15) + noise
                out[i, c] = np.clip(x, -32768, 32767).astype(np.int16)
    return out
