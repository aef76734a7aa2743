"""Drop-in AADDecoder_DecodeWhole (malloc'd .aad in, malloc'd int32 rows out, src/main.c:94-108) on a 1-hour 48 kHz stereo
4-bit stream, timed for several host-thread counts and ring-slice sizes (GPU box).  The library reads
AAD_B200_HOST_THREADS / AAD_B200_RING_MIB once per process, so every point runs in a process of its own.
python tools/dropin_sweep.py            -> one JSON line with all points
python tools/dropin_sweep.py one        -> this process's point (used by the sweep)"""
import ctypes as C, json, os, subprocess, sys, time
import numpy as np
sys.path.insert(0, '.')


def one():
    import torch
    import aad_b200
    from aad_b200.capi import OK, make_param, _planar_pointers
    api, gpu = aad_b200.load()
    ctx = gpu.create(0)
    dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream
    ch, rate, bits, n = 2, 48000, 4, int(os.environ.get("SWEEP_SAMPLES", "172800000"))
    prm = make_param(ch, rate, bits, 1024, False, 2)
    b = gpu.batch(1, n, prm)
    size = gpu.stream_bytes(prm, n)
    pcm = torch.empty((1, ch, n), dtype=torch.int16, device=dev)
    aad = torch.zeros((1, int(b.aad_stream_stride)), dtype=torch.uint8, device=dev)
    assert gpu.lib.AADGpu_SynthBatchDevice(ctx, C.byref(b), 7, pcm.data_ptr(), s) == OK
    assert gpu.lib.AADGpu_SetEncodeSegmentBlocks(ctx, 64) == OK
    assert gpu.lib.AADGpu_EncodeBatchDevice(ctx, C.byref(b), pcm.data_ptr(), None, aad.data_ptr(), None, s) == OK
    assert gpu.lib.AADGpu_SetEncodeSegmentBlocks(ctx, 0) == OK
    torch.cuda.synchronize()
    blob = np.frombuffer(aad[0, :size].cpu().numpy().tobytes(), dtype=np.uint8)
    rows = [np.zeros(n, dtype=np.int32) for _ in range(ch)]
    ptrs = _planar_pointers(rows)
    handle = api.lib.AADDecoder_Create(None, 0)
    secs = []
    for _ in range(5):
        t0 = time.perf_counter()
        rc = api.lib.AADDecoder_DecodeWhole(handle, blob.ctypes.data_as(C.POINTER(C.c_uint8)), len(blob), ptrs, ch, n)
        secs.append(time.perf_counter() - t0)
        assert rc == OK
    api.lib.AADDecoder_Destroy(handle)
    print(json.dumps({"threads": os.environ.get("AAD_B200_HOST_THREADS", "default"), "ring_mib": os.environ.get("AAD_B200_RING_MIB", "default"),
                      "ms": [round(1e3 * x, 2) for x in secs], "best_ms": round(1e3 * min(secs[1:]), 2),
                      "sum": int(rows[0][::1009].astype(np.int64).sum() + rows[1][::1013].astype(np.int64).sum())}))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one()
    else:
        out = []
        K = {"threads": "AAD_B200_HOST_THREADS", "ring_mib": "AAD_B200_RING_MIB", "conv_log2": "AAD_B200_CONV_PIECE_LOG2",
             "copy_log2": "AAD_B200_COPY_PIECE_LOG2", "no_avx2": "AAD_B200_NO_AVX2"}
        points = [{}, {"no_avx2": 1}, {}, {"no_avx2": 1}, {"threads": 8}, {"threads": 8, "no_avx2": 1}, {"threads": 12}]
        for pt in points:
            env = dict(os.environ)
            env["AAD_B200_TRACE"] = "1"
            for k, v in pt.items():
                env[K[k]] = str(v)
            r = subprocess.run([sys.executable, __file__, "one"], env=env, capture_output=True, text=True)
            if r.returncode == 0:
                res = json.loads(r.stdout.strip().splitlines()[-1])
                trace = [l for l in r.stderr.splitlines() if "drop-in decode" in l]
                out.append({"point": pt or "shipped", "best_ms": res["best_ms"], "ms": res["ms"], "sum": res["sum"],
                            "phases": trace[-1].split(": ", 2)[-1] if trace else None})
            else:
                out.append({"point": pt, "error": r.stderr[-300:]})
        print(json.dumps({"cpus": os.cpu_count(), "points": out}))
