"""A/B of the decoder's input staging on the bench shapes (GPU box): default kernel (16-byte loads prefetched in
registers + shared stores) against kernel path 7 (aad_decode_tma: cp.async.bulk.tensor through a tensor map + mbarrier).
The TMA unit needs 16-byte aligned blocks, so BOTH arms run on the same aligned layout: streams at base + 1 (block 0
at byte 32 of the allocation), stream stride rounded up to a multiple of 16.  Prints min / all times and checksums.
python tools/dec_tma_ab.py [shape ...]   (shape = c1b4 | c1b2)"""
import ctypes as C, sys, json, os
import torch
sys.path.insert(0, '.')
import aad_b200
from aad_b200.capi import OK, make_param
api, gpu = aad_b200.load()
ctx = gpu.create(0)
dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
N, n = int(os.environ.get("AB_CLIPS", "12500")), 441000   # AB_CLIPS=1500 for ncu captures
res = {}
for shape in (sys.argv[1:] or ["c1b4", "c1b2"]):
    ch, bits = int(shape[1]), int(shape[3])
    prm = make_param(ch, 44100, bits, 1024, False, 0)
    stride = (gpu.stream_bytes_bound(prm, n) + 15) // 16 * 16
    b = gpu.batch(N, n, prm, aad_stream_stride=stride)
    pcm = torch.zeros((N, ch, n), dtype=torch.int16, device=dev)
    raw = torch.zeros(N * stride + 16, dtype=torch.uint8, device=dev)
    aad_ptr = raw.data_ptr() + 1
    assert (aad_ptr + 31) % 16 == 0
    assert gpu.lib.AADGpu_SynthBatchDevice(ctx, C.byref(b), 0, pcm.data_ptr(), s) == OK
    assert gpu.lib.AADGpu_EncodeBatchDevice(ctx, C.byref(b), pcm.data_ptr(), None, aad_ptr, None, s) == OK
    out = torch.zeros_like(pcm)
    for path in (0, 7, 8, 6, 0, 7, 8, 6):   # 6 = default kernel with per-stream tasks (what aad_decode_tma's tasks are)
        gpu.lib.AADGpu_SetKernelPath(path)
        before = int(gpu.lib.AADGpu_TmaLaunchCount())
        out.zero_()
        times = []
        for rep in range(3):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            for it in range(8):
                if it == 3: ev[0].record()
                assert gpu.lib.AADGpu_DecodeBatchDevice(ctx, C.byref(b), aad_ptr, None, out.data_ptr(), s) == OK
            ev[1].record(); torch.cuda.synchronize()
            times.append(round(ev[0].elapsed_time(ev[1]) / 5, 4))
        flat = out.view(-1)
        key = f"{shape}_path{path}"
        res.setdefault(key, []).append({"ms": min(times), "all": times, "tma_launches": int(gpu.lib.AADGpu_TmaLaunchCount()) - before,
                                        "sum": [int(torch.sum(flat, dtype=torch.int64).item()), int(torch.sum(flat[3::1013], dtype=torch.int64).item())]})
    gpu.lib.AADGpu_SetKernelPath(0)
    del pcm, raw, out
    torch.cuda.empty_cache()
print(json.dumps(res))
