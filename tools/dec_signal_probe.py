"""How much of the decoder's time is shared-memory bank conflicts of the q-table lookup?  Decode the bench batch shape
(12,500 mono 10-s clips, 4-bit) for signals whose code / step-row statistics differ: the bench's synthetic signal,
silence (every lane reads the same table word: broadcast, conflict free), and white noise at two levels.
python tools/dec_signal_probe.py [bits]"""
import ctypes as C, sys, json
import torch
sys.path.insert(0, '.')
import aad_b200
from aad_b200.capi import OK, make_param
api, gpu = aad_b200.load()
ctx = gpu.create(0)
dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
N, n = 12500, 441000
bits = int(sys.argv[1]) if len(sys.argv) > 1 else 4
prm = make_param(1, 44100, bits, 1024, False, 0)
b = gpu.batch(N, n, prm)
res = {}
g = torch.Generator(device=dev); g.manual_seed(1)
for name in ("synth", "silence", "noise_full", "noise_small"):
    pcm = torch.zeros((N, 1, n), dtype=torch.int16, device=dev)
    if name == "synth":
        assert gpu.lib.AADGpu_SynthBatchDevice(ctx, C.byref(b), 0, pcm.data_ptr(), s) == OK
    elif name == "noise_full":
        pcm.random_(-32768, 32767, generator=g)
    elif name == "noise_small":
        pcm.random_(-200, 200, generator=g)
    aad = torch.zeros((N, b.aad_stream_stride), dtype=torch.uint8, device=dev)
    assert gpu.lib.AADGpu_EncodeBatchDevice(ctx, C.byref(b), pcm.data_ptr(), None, aad.data_ptr(), None, s) == OK
    out = torch.zeros_like(pcm)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for it in range(8):
        if it == 3: ev[0].record()
        assert gpu.lib.AADGpu_DecodeBatchDevice(ctx, C.byref(b), aad.data_ptr(), None, out.data_ptr(), s) == OK
    ev[1].record(); torch.cuda.synchronize()
    res[name] = round(ev[0].elapsed_time(ev[1]) / 5, 4)
    del pcm, aad, out
    torch.cuda.empty_cache()
print(json.dumps(res))
