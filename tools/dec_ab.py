"""A/B of the decoder's flush on the bench batch (GPU box): kernel path 0 = register flush, 4 = TMA bulk store (mono 4-bit
only; the other shapes run the same kernel on both paths).  python tools/dec_ab.py"""
import ctypes as C, sys, json
import torch
sys.path.insert(0, '.')
import aad_b200
from aad_b200.capi import OK, make_param
api, gpu = aad_b200.load()
ctx = gpu.create(0)
dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
N, n = 12500, 441000
res = {}
for ch, bits in ((1, 4), (2, 4), (1, 2), (1, 3)):
    Nc = N // ch
    prm = make_param(ch, 44100, bits, 1024, False, 0)
    b = gpu.batch(Nc, n, prm)
    pcm = torch.zeros((Nc, ch, n), dtype=torch.int16, device=dev)
    aad = torch.zeros((Nc, b.aad_stream_stride), dtype=torch.uint8, device=dev)
    assert gpu.lib.AADGpu_SynthBatchDevice(ctx, C.byref(b), 0, pcm.data_ptr(), s) == OK
    assert gpu.lib.AADGpu_EncodeBatchDevice(ctx, C.byref(b), pcm.data_ptr(), None, aad.data_ptr(), None, s) == OK
    outs = {}
    for path in (0, 4):
        gpu.lib.AADGpu_SetKernelPath(path)
        out = torch.zeros_like(pcm)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for it in range(8):
            if it == 3: ev[0].record()
            assert gpu.lib.AADGpu_DecodeBatchDevice(ctx, C.byref(b), aad.data_ptr(), None, out.data_ptr(), s) == OK
        ev[1].record(); torch.cuda.synchronize()
        res[f"c{ch}b{bits}_path{path}"] = round(ev[0].elapsed_time(ev[1]) / 5, 4)
        outs[path] = out
    res[f"c{ch}b{bits}_equal"] = bool(torch.equal(outs[0], outs[4]))
    gpu.lib.AADGpu_SetKernelPath(0)
    del pcm, aad, outs, out
    torch.cuda.empty_cache()
print(json.dumps(res))
