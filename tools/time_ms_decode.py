"""Decode time of a stereo batch with and without mid/side (GPU box only): python tools/time_ms_decode.py [clips]"""
import ctypes as C
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import aad_b200
from aad_b200.capi import OK, make_param

api, gpu = aad_b200.load()
ctx = gpu.create(0)
clips, n = (int(sys.argv[1]) if len(sys.argv) > 1 else 6250), 441000
dev = torch.device("cuda:0")
s = torch.cuda.current_stream().cuda_stream
for ms in (False, True):
    prm = make_param(2, 44100, 4, 1024, ms, 0)
    b = gpu.batch(clips, n, prm)
    pcm = torch.zeros((clips, 2, n), dtype=torch.int16, device=dev)
    aad = torch.zeros((clips, b.aad_stream_stride), dtype=torch.uint8, device=dev)
    out = torch.zeros_like(pcm)
    assert gpu.lib.AADGpu_SynthBatchDevice(ctx, C.byref(b), 0, pcm.data_ptr(), s) == OK
    assert gpu.lib.AADGpu_EncodeBatchDevice(ctx, C.byref(b), pcm.data_ptr(), None, aad.data_ptr(), None, s) == OK
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for it in range(6):
        if it == 3:
            ev[0].record()
        assert gpu.lib.AADGpu_DecodeBatchDevice(ctx, C.byref(b), aad.data_ptr(), None, out.data_ptr(), s) == OK
    ev[1].record()
    torch.cuda.synchronize()
    err = (out.float() - pcm.float()).pow(2).mean().sqrt().item()
    print(f"stereo 4-bit, ms={int(ms)}: decode {ev[0].elapsed_time(ev[1]) / 3:.3f} ms per batch of {clips} clips, round-trip rms error {err:.1f}")
    del pcm, aad, out
gpu.destroy(ctx)
