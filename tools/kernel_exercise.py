"""Small end-to-end exercise of every kernel (GPU box only): ragged batches through the fast, wide and generic
kernels, truncated streams, segment mode, sliced host pipelines, the WAV analysis kernels, every result that has
an oracle compared with it.  Written to run under `compute-sanitizer --tool memcheck`; that tool is closed on
this pool (gpurun answers "closed"), so it runs plain: python tools/kernel_exercise.py"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import aad_b200
import aadtest
from aad_b200.capi import make_param, OK

api, gpu = aad_b200.load()
oracle = aadtest.Oracle(ROOT / "oracle" / "liboracle.so")
ctx = gpu.create(0)
rng = np.random.default_rng(1)
checked = 0
for channels in (1, 2, 3, 5, 8):
    for bits in (2, 3, 4):
        n_streams, n_max = 9, 2600
        lens = rng.integers(1, n_max + 1, size=n_streams).astype(np.uint32)
        lens[0] = n_max
        pcm = np.zeros((n_streams, channels, n_max), dtype=np.int16)
        for i in range(n_streams):
            pcm[i, :, :lens[i]] = aadtest.signal(aadtest.SIGNALS[i % len(aadtest.SIGNALS)], channels, int(lens[i]), i)
        for block, seg in ((256 * channels // (2 if channels >= 5 else 1), 0), (1024, 2)):
            ms = channels >= 2 and bits == 3
            for path in (0, 1, 2):
                gpu.lib.AADGpu_SetKernelPath(path)
                gpu.lib.AADGpu_SetEncodeSegmentBlocks(ctx, seg)
                aad, sizes = gpu.encode_batch(ctx, pcm, 44100, bits, block, ms, 1, num_samples=lens)
                gpu.lib.AADGpu_SetEncodeSegmentBlocks(ctx, 0)
                cut = sizes.copy()
                cut[1::3] = np.maximum(31, cut[1::3] - 37)
                dec = gpu.decode_batch(ctx, aad, n_max, 44100, channels, bits, block, ms, sizes=sizes)
                gpu.decode_batch(ctx, aad, n_max, 44100, channels, bits, block, ms, sizes=cut)
                for i in range(0, n_streams, 4):
                    rc, want, _ = oracle.decode(aad[i, :sizes[i]].tobytes())
                    assert rc == 0 and np.array_equal(dec[i, :, :lens[i]], want), (channels, bits, block, path, i)
                    if seg == 0:
                        rc, enc = oracle.encode(pcm[i, :, :lens[i]], 44100, bits, block, ms, 1)
                        assert rc == 0 and aad[i, :sizes[i]].tobytes() == enc, (channels, bits, block, path, i)
                    checked += 1
            gpu.lib.AADGpu_SetKernelPath(0)
# sliced host pipeline + analysis kernels
prm = make_param(2, 32000, 4, 1024, False, 2)
n = 50_001
for wav_bits in (8, 16, 24, 32):
    raw = rng.integers(0, 256, size=n * 2 * wav_bits // 8, dtype=np.uint8)
    out = np.zeros_like(raw)
    stats = (C.c_double * 3)()
    for what in (0, 1):
        assert gpu.lib.AADGpu_AnalyzeWav(ctx, C.byref(prm), raw.ctypes.data, wav_bits, n, what, out.ctypes.data, None, None) == OK
    assert gpu.lib.AADGpu_AnalyzeWav(ctx, C.byref(prm), raw.ctypes.data, wav_bits, n, 2, None, stats, None) == OK
pcm = np.stack([aadtest.signal("music", 2, 40000, i) for i in range(40)])
b = gpu.batch(40, 40000, prm)
aad = np.zeros((40, b.aad_stream_stride), dtype=np.uint8)
sizes = np.zeros(40, dtype=np.uint32)
out = np.zeros_like(pcm)
assert gpu.lib.AADGpu_ReconstructBatch(ctx, C.byref(b), pcm.ctypes.data, None, aad.ctypes.data, sizes.ctypes.data, out.ctypes.data) == OK
rc, want = oracle.encode(pcm[7], 32000, 4, 1024, False, 2)
assert rc == 0 and aad[7, :sizes[7]].tobytes() == want
gpu.destroy(ctx)
print(f"kernel exercise ok: {checked} streams compared with the oracle")
