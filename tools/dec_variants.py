"""Time the decoder of one build of the library on the bench shapes and print a checksum of what it decoded (GPU box):
AAD_B200_LIBRARY=aad_b200/exp/libaad_X.so python tools/dec_variants.py [shape ...]   (shape = c<channels>b<bits>)
Used for A/B runs of kernel variants built with tools/build_variant.sh; the checksums of all variants must agree."""
import ctypes as C, sys, json, os
import torch
sys.path.insert(0, '.')
import aad_b200
from aad_b200.capi import OK, make_param
api, gpu = aad_b200.load()
ctx = gpu.create(0)
gpu.lib.AADGpu_SetKernelPath(int(os.environ.get("AAD_KERNEL_PATH", "0")))   # 5 / 6: tasks always / never span streams
dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
N, n = 12500, 441000
shapes = sys.argv[1:] or ["c1b4", "c2b4", "c1b2", "c1b3"]
res = {"lib": os.environ.get("AAD_B200_LIBRARY", "in-tree"), "path": os.environ.get("AAD_KERNEL_PATH", "0")}
for shape in shapes:
    ch, bits = int(shape[1]), int(shape[3])
    Nc = N // ch
    prm = make_param(ch, 44100, bits, 1024, False, 0)
    b = gpu.batch(Nc, n, prm)
    pcm = torch.zeros((Nc, ch, n), dtype=torch.int16, device=dev)
    aad = torch.zeros((Nc, b.aad_stream_stride), dtype=torch.uint8, device=dev)
    assert gpu.lib.AADGpu_SynthBatchDevice(ctx, C.byref(b), 0, pcm.data_ptr(), s) == OK
    assert gpu.lib.AADGpu_EncodeBatchDevice(ctx, C.byref(b), pcm.data_ptr(), None, aad.data_ptr(), None, s) == OK
    out = torch.zeros_like(pcm)
    times = []
    for rep in range(3):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for it in range(8):
            if it == 3: ev[0].record()
            assert gpu.lib.AADGpu_DecodeBatchDevice(ctx, C.byref(b), aad.data_ptr(), None, out.data_ptr(), s) == OK
        ev[1].record(); torch.cuda.synchronize()
        times.append(round(ev[0].elapsed_time(ev[1]) / 5, 4))
    res[shape] = min(times)
    res[shape + "_all"] = times
    flat = out.view(-1)
    res[shape + "_sum"] = [int(torch.sum(flat, dtype=torch.int64).item()), int(torch.sum(flat[::7], dtype=torch.int64).item()),
                           int(torch.sum(flat[3::1013], dtype=torch.int64).item())]
    del pcm, aad, out
    torch.cuda.empty_cache()
print(json.dumps(res))
