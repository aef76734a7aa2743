"""Time the encoder of one build of the library on shapes that take the different kernels / schedules (GPU box):
AAD_B200_LIBRARY=aad_b200/exp/libaad_X.so python tools/enc_variants.py
shapes: (clips, samples per clip, trials): 12,500 x 441,000 (helper-lane schedule at 2 trials, plain kernel at 0),
50,000 x 110,250 (plain kernel, 2.6 warps per scheduler), 18,000 x 220,500 (pairing).  Prints ms and a checksum of the bytes."""
import ctypes as C, sys, json, os
import torch
sys.path.insert(0, '.')
import aad_b200
from aad_b200.capi import OK, make_param
api, gpu = aad_b200.load()
ctx = gpu.create(0)
dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
res = {"lib": os.environ.get("AAD_B200_LIBRARY", "in-tree")}
for N, n, trials in ((12500, 441000, 2), (12500, 441000, 0), (50000, 110250, 2), (50000, 110250, 0), (18000, 220500, 2)):
    prm = make_param(1, 44100, 4, 1024, False, trials)
    b = gpu.batch(N, n, prm)
    pcm = torch.zeros((N, 1, n), dtype=torch.int16, device=dev)
    aad = torch.zeros((N, b.aad_stream_stride), dtype=torch.uint8, device=dev)
    assert gpu.lib.AADGpu_SynthBatchDevice(ctx, C.byref(b), 0, pcm.data_ptr(), s) == OK
    times = []
    for rep in range(2):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for it in range(3):
            if it == 1: ev[0].record()
            assert gpu.lib.AADGpu_EncodeBatchDevice(ctx, C.byref(b), pcm.data_ptr(), None, aad.data_ptr(), None, s) == OK
        ev[1].record(); torch.cuda.synchronize()
        times.append(round(ev[0].elapsed_time(ev[1]) / 2, 3))
    key = f"{N}x{n}_t{trials}"
    res[key] = min(times)
    flat = aad.view(-1)
    res[key + "_sum"] = [int(torch.sum(flat, dtype=torch.int64).item()), int(torch.sum(flat[::7], dtype=torch.int64).item())]
    del pcm, aad
    torch.cuda.empty_cache()
print(json.dumps(res))
