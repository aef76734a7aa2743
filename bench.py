#!/usr/bin/env python3
"""bench.py -- AAD ADPCM encode+decode throughput on B200, one process per GPU.

Workload (BASELINE.json configs[4]): a batch of synthetic 10-second 44.1 kHz mono clips,
4-bit, CLI-default encoder settings (block 1024, 2 encode trials, no MS), encoded and decoded
back.  The named config is 100,000 clips sharded over 8 GPUs = 12,500 clips per GPU; that
per-GPU share is what every rank processes (weak scaling, no data-path collective -- clips are
independent).  A "step" = encode the whole per-GPU batch + decode it back.

  value     round-trip Msamples/s with PCM / .aad resident in HBM (kernels only, CUDA events)
  e2e       the same round trip through the host C ABI from pinned host buffers, H2D + kernels + D2H
            inside the timed region: AADGpu_ReconstructBatch (one call, headline) and the pair
            AADGpu_EncodeBatch / AADGpu_DecodeBatch (e2e.separate_calls); e2e.link_probe_gbs = what plain
            pinned 1-D copies reach on every rank at the same time, e2e.frac_of_link = the e2e traffic over it
  roofline  dominant kernel (the encoder at 2 trials) against the measured HBM copy bandwidth
  cpu_baseline  the reference codec on this box's host cores, bounded sample, same clips;
            its output doubles as the parity check of the GPU result (every rank checks clips of its own)
  decode_sweep / encode_sweep   the kernels at the other bit depths / channel counts / trial counts
  other_configs   BASELINE configs[2] and [3]: ONE long stream (1 h 48 kHz stereo 4-bit; 30 min 96 kHz 8-channel
            3-bit) -- decode device-resident and end to end from pinned host memory, on one device and by block
            range over all local GPUs (AADGpuGroup_*, rank 0 drives the group while the other ranks idle);
            bit-exact encode (2 / 8 chains: chain-bound) and the labelled segment-mode encode; parity vs the
            reference / oracle

`--impl reference` times the reference's own CPU implementation (oracle/_ref when it was
compiled, else the oracle port) on all host threads, same metric and config.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "encode+decode round-trip throughput (bit-exact AAD ADPCM)"
UNIT = "Msamples/s"
RATE, CLIP_SAMPLES, CHANNELS, BITS, MAX_BLOCK, TRIALS = 44100, 441000, 1, 4, 1024, 2
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel (aad_encode_roles<4,0,0>) at the
# default workload, from an `ncu --set full` capture of this command kept under profiles/ -- NOT measured in
# the run, hence reported as traffic_from_profile (roofline.traffic itself stays null).
TRAFFIC_FROM_PROFILE = {"bytes": 21.07e9, "profile": "profiles/r02b_encode_roles.md"}
LONG_STREAMS = {
    "config3": {"channels": 2, "rate": 48000, "bits": 4, "samples": 172_800_000,
                "workload": "BASELINE configs[2]: synthetic 1-hour 48 kHz 16-bit stereo stream, 4-bit, block 1024"},
    "config4": {"channels": 8, "rate": 96000, "bits": 3, "samples": 172_800_000,
                "workload": "BASELINE configs[3]: synthetic 30-minute 96 kHz 8-channel stream, 3-bit, block 1024"},
}
SEGMENT_BLOCKS = 64


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--clips", type=int, default=12500, help="clips per GPU")
    ap.add_argument("--samples", type=int, default=CLIP_SAMPLES, help="samples per clip")
    ap.add_argument("--trials", type=int, default=TRIALS)
    ap.add_argument("--bits", type=int, default=BITS)
    ap.add_argument("--channels", type=int, default=CHANNELS)
    ap.add_argument("--no-pairing", action="store_true", help="encoder: never interleave two passes in one thread")
    ap.add_argument("--schedule", type=int, default=1, help="encoder pass schedule (AADGpu_SetEncoderSchedule): 1 = by shape")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-clips", type=int, default=96, help="clips in the bounded CPU-baseline sample")
    ap.add_argument("--no-sweeps", action="store_true", help="skip decode_sweep / encode_sweep")
    ap.add_argument("--no-long", action="store_true", help="skip other_configs (the long single streams)")
    ap.add_argument("--long-samples", type=int, default=0, help="samples per channel of the long streams (0 = BASELINE's)")
    return ap.parse_args()


def workload_config(args, n_gpus):
    return {
        "workload": f"batch of synthetic {args.samples / RATE:g}-second 44.1 kHz {args.channels}-channel clips, "
                    f"{args.bits}-bit, block {MAX_BLOCK}, {args.trials} encode trials, encode+decode "
                    f"(BASELINE configs[4]: 100k clips over 8 GPUs = 12.5k clips/GPU)",
        "clips_per_gpu": args.clips, "clips_total": args.clips * n_gpus, "samples_per_clip": args.samples,
        "channels": args.channels, "bits_per_sample": args.bits, "max_block_size": MAX_BLOCK,
        "num_encode_trials": args.trials, "sharding": f"clips x{n_gpus}, no collective",
        "l2": "inputs larger than L2 (per-GPU PCM batch >> 126 MB), no flush needed",
    }


# ---- reference / oracle on the CPU ---------------------------------------------------------------

class CpuCodec:
    """The reference codec on host cores: oracle/_ref/libaad_ref.so (the compiled, unmodified
    reference) when present, else the oracle port (oracle/liboracle.so)."""

    def __init__(self, channels=1):
        from aad_b200.capi import AADCApi
        # the stock reference stops at 2 channels (src/aad.h:13); libaad_ref8.so = the same sources with that limit at 8
        ref = ROOT / "oracle" / "_ref" / ("libaad_ref.so" if channels <= 2 else "libaad_ref8.so")
        if ref.exists():
            self.kind, self.api = "reference", AADCApi(ref)
        else:
            sys.path.insert(0, str(ROOT / "tests"))
            import aadtest
            port = ROOT / "oracle" / "liboracle.so"
            if not port.exists():
                import subprocess
                subprocess.run(["make", "-C", str(ROOT / "oracle"), "liboracle.so"], check=True, stdout=subprocess.DEVNULL)
            self.kind, self.port = "port", aadtest.Oracle(port)

    def encode(self, pcm16, rate, bits, trials):
        if self.kind == "reference":
            rc, data = self.api.encode_whole(pcm16.astype(np.int32), rate, bits, MAX_BLOCK, False, trials)
        else:
            rc, data = self.port.encode(pcm16, rate, bits, MAX_BLOCK, False, trials)
        assert rc == 0
        return data

    def decode(self, data):
        """-> int16 [channels, n]"""
        if self.kind == "reference":
            rc, dec, _ = self.api.decode_whole(data)
            assert rc == 0
            return dec.astype(np.int16)
        rc, dec, _ = self.port.decode(data)
        assert rc == 0
        return dec

    def roundtrip(self, pcm16, bits, trials):
        """pcm16 [channels, n] -> (aad bytes, decoded int16 [channels, n], seconds in the codec)."""
        if self.kind == "reference":
            pcm32 = pcm16.astype(np.int32)
            t0 = time.perf_counter()
            rc, data = self.api.encode_whole(pcm32, RATE, bits, MAX_BLOCK, False, trials)
            rc2, dec, _ = self.api.decode_whole(data)
            dt = time.perf_counter() - t0
            assert rc == 0 and rc2 == 0
            return data, dec.astype(np.int16), dt
        t0 = time.perf_counter()
        rc, data = self.port.encode(pcm16, RATE, bits, MAX_BLOCK, False, trials)
        rc2, dec, _ = self.port.decode(data)
        dt = time.perf_counter() - t0
        assert rc == 0 and rc2 == 0
        return data, dec, dt


def run_reference_arm(args):
    """All host threads, each on its own clips (the reference has no threading of its own)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from aad_b200.synth import synth_pcm16      # numpy mirror of the generator; libaad_b200.so is not loaded here
    codec = CpuCodec()
    cores = os.cpu_count() or 1
    per_step = cores * 4
    # the generator's sine table, as AADGpu_SynthLut builds it: lrint(32767 sin(2 pi k / 1024))
    lut = np.rint(32767.0 * np.sin(2.0 * 3.14159265358979323846 * np.arange(1024) / 1024.0)).astype(np.int16)
    clips = synth_pcm16(lut, 0, per_step, args.channels, args.samples, RATE)

    def one(i):
        return codec.roundtrip(clips[i], args.bits, args.trials)[2]

    def step():
        t0 = time.perf_counter()
        with ThreadPoolExecutor(cores) as pool:
            list(pool.map(one, range(per_step)))
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        step()
    t = sum(step() for _ in range(args.steps))
    samples = per_step * args.channels * args.samples * args.steps
    value = samples / t / 1e6
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * t / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": cores, "kind": codec.kind,
                         "sample": f"{per_step} clips per step ({per_step * args.samples * args.channels / 1e6:.1f} Msamples), "
                                   f"one clip per thread at a time, {cores} threads"},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- clocks ----------------------------------------------------------------------------------------

class ClockSampler:
    """SM clock and throttle reasons during the timed region (NVML, 100 ms period)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.reasons |= {name for bit, name in self.REASONS.items() if mask & bit}
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv:
            self.thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self.nv:
            self.thread.join()

    def summary(self):
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---- the B200 arm ----------------------------------------------------------------------------------

def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _timed_kernel(torch, fn, iters=3):
    """mean device time of fn() in ms over `iters` launches after one warm-up, CUDA events on the current stream"""
    fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(iters):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / iters


def run_sweeps(args, torch, gpu, ctx, dev, check, total_samples, peak):
    """decode_sweep: bits 2/3/4 x channels 1/2/8; encode_sweep: bits 2/3/4 x trials 0/2 (mono).  Same sample count
    per launch as the headline batch (the clip count shrinks with the channel count), device resident, kernels only."""
    from aad_b200.capi import make_param
    stream = torch.cuda.current_stream().cuda_stream
    n = args.samples
    dec, enc = {}, {}
    for ch in (1, 2, 8):
        N = max(1, total_samples // (ch * n))
        pcm = torch.empty((N, ch, n), dtype=torch.int16, device=dev)
        out = torch.empty_like(pcm)
        b_any = gpu.batch(N, n, make_param(ch, RATE, 4, MAX_BLOCK, False, 0))
        check(gpu.lib.AADGpu_SynthBatchDevice(ctx, C.byref(b_any), 0, pcm.data_ptr(), stream), "synth")
        for bits in (2, 3, 4):
            prm = make_param(ch, RATE, bits, MAX_BLOCK, False, 0)
            b = gpu.batch(N, n, prm)
            aad = torch.zeros((N, int(b.aad_stream_stride)), dtype=torch.uint8, device=dev)
            bps = 2.0 + gpu.stream_bytes(prm, n) / (ch * n)
            samples = N * ch * n
            for trials in ((0, 2) if ch == 1 else (0,)):
                bt = gpu.batch(N, n, make_param(ch, RATE, bits, MAX_BLOCK, False, trials))
                ms = _timed_kernel(torch, lambda: check(gpu.lib.AADGpu_EncodeBatchDevice(
                    ctx, C.byref(bt), pcm.data_ptr(), None, aad.data_ptr(), None, stream), "encode"), iters=2)
                if ch == 1:
                    enc[f"b{bits}_t{trials}"] = {"kernel_ms": round(ms, 3), "msamples_s": round(samples / ms / 1e3, 1),
                                                "hbm_frac": round(samples * bps / (ms * 1e-3) / 1e9 / peak, 5)}
            ms = _timed_kernel(torch, lambda: check(gpu.lib.AADGpu_DecodeBatchDevice(
                ctx, C.byref(b), aad.data_ptr(), None, out.data_ptr(), stream), "decode"))
            dec[f"b{bits}_c{ch}"] = {"kernel_ms": round(ms, 3), "msamples_s": round(samples / ms / 1e3, 1),
                                     "hbm_frac": round(samples * bps / (ms * 1e-3) / 1e9 / peak, 5), "clips": N}
            del aad
        del pcm, out
        torch.cuda.empty_cache()
    # short clips: 1-second mono clips are 22 blocks each, so a decoder whose warp tasks stay inside one stream idles a
    # third of its lanes; aad_decode_fast's tasks run on into the next stream (kernel path 6 = per-stream tasks, for the A/B)
    n1 = RATE
    N = max(1, total_samples // n1)
    prm = make_param(1, RATE, 4, MAX_BLOCK, False, 0)
    b = gpu.batch(N, n1, prm)
    pcm = torch.empty((N, 1, n1), dtype=torch.int16, device=dev)
    out = torch.empty_like(pcm)
    aad = torch.zeros((N, int(b.aad_stream_stride)), dtype=torch.uint8, device=dev)
    check(gpu.lib.AADGpu_SynthBatchDevice(ctx, C.byref(b), 0, pcm.data_ptr(), stream), "synth")
    check(gpu.lib.AADGpu_EncodeBatchDevice(ctx, C.byref(b), pcm.data_ptr(), None, aad.data_ptr(), None, stream), "encode")
    bps = 2.0 + gpu.stream_bytes(prm, n1) / n1
    decode = lambda: check(gpu.lib.AADGpu_DecodeBatchDevice(ctx, C.byref(b), aad.data_ptr(), None, out.data_ptr(), stream), "decode")
    ms = _timed_kernel(torch, decode)
    gpu.lib.AADGpu_SetKernelPath(6)
    try:
        ms_per_stream = _timed_kernel(torch, decode)
    finally:
        gpu.lib.AADGpu_SetKernelPath(0)
    dec["b4_c1_1s_clips"] = {"kernel_ms": round(ms, 3), "msamples_s": round(N * n1 / ms / 1e3, 1),
                             "hbm_frac": round(N * n1 * bps / (ms * 1e-3) / 1e9 / peak, 5), "clips": N,
                             "samples_per_clip": n1, "kernel_ms_with_per_stream_tasks": round(ms_per_stream, 3)}
    del pcm, out, aad
    torch.cuda.empty_cache()
    return dec, enc


def run_long_stream(name, spec, args, torch, api, gpu, ctx, dev, check, world, peak, link):
    """One long stream (BASELINE configs[2] / [3]) on rank 0; see the module docstring."""
    from aad_b200.capi import make_param
    stream = torch.cuda.current_stream().cuda_stream
    ch, rate, bits = spec["channels"], spec["rate"], spec["bits"]
    n = args.long_samples or spec["samples"]
    trials = TRIALS
    prm = make_param(ch, rate, bits, MAX_BLOCK, False, trials)
    b = gpu.batch(1, n, prm)
    size = gpu.stream_bytes(prm, n)
    rc, bs, spb = api.calculate_block_size(MAX_BLOCK, ch, bits)       # AADEncoder_CalculateBlockSize, src/aad_encoder.c:85-131
    assert rc == 0
    bps = 2.0 + size / (ch * n)
    samples = ch * n
    res = {"workload": spec["workload"], "channels": ch, "bits_per_sample": bits, "samples_per_channel": n,
           "aad_bytes": size, "pcm_bytes": samples * 2, "algorithmic_bytes_per_sample": round(bps, 4)}
    parity = {}

    pcm = torch.empty((1, ch, n), dtype=torch.int16, device=dev)
    aad = torch.zeros((1, int(b.aad_stream_stride)), dtype=torch.uint8, device=dev)
    out = torch.empty_like(pcm)
    check(gpu.lib.AADGpu_SynthBatchDevice(ctx, C.byref(b), 7, pcm.data_ptr(), stream), "synth")

    # ---- segment-mode encode, device resident (extension: NOT byte-identical to the reference encoder; every
    #      segment is what the reference produces for those samples on a fresh handle) -- also makes the stream
    #      the decode legs work on (a valid .aad stream: the stock decoder decodes it)
    check(gpu.lib.AADGpu_SetEncodeSegmentBlocks(ctx, SEGMENT_BLOCKS), "segments on")
    try:
        ms = _timed_kernel(torch, lambda: check(gpu.lib.AADGpu_EncodeBatchDevice(
            ctx, C.byref(b), pcm.data_ptr(), None, aad.data_ptr(), None, stream), "segment encode"), iters=2)
    finally:
        check(gpu.lib.AADGpu_SetEncodeSegmentBlocks(ctx, 0), "segments off")
    nblocks = -(-n // spb)
    res["encode_segment_mode"] = {
        "label": "extension, not byte-identical to the reference encoder (DESIGN.md 4.4)", "segment_blocks": SEGMENT_BLOCKS,
        "chains": -(-nblocks // SEGMENT_BLOCKS) * ch, "trials": trials,
        "kernel_ms": round(ms, 3), "msamples_s": round(samples / ms / 1e3, 1),
        "hbm_frac": round(samples * bps / (ms * 1e-3) / 1e9 / peak, 5)}

    # ---- decode, device resident (planar rows)
    ms = _timed_kernel(torch, lambda: check(gpu.lib.AADGpu_DecodeBatchDevice(
        ctx, C.byref(b), aad.data_ptr(), None, out.data_ptr(), stream), "decode"))
    res["decode_device"] = {"kernel_ms": round(ms, 3), "msamples_s": round(samples / ms / 1e3, 1),
                            "hbm_frac": round(samples * bps / (ms * 1e-3) / 1e9 / peak, 5),
                            "chains": nblocks * ch}

    # ---- bit-exact encode: `channels` serial chains (src/aad_encoder.c:853-886), timed on a prefix (the rate of
    #      a chain does not depend on its length) and compared byte for byte with the reference's encode of it
    n_pre = min(n, (2_000_000 // spb) * spb)
    bp = gpu.batch(1, n_pre, prm, pcm_channel_stride=n, pcm_stream_stride=ch * n)
    pre = torch.zeros((1, int(bp.aad_stream_stride)), dtype=torch.uint8, device=dev)
    ms = _timed_kernel(torch, lambda: check(gpu.lib.AADGpu_EncodeBatchDevice(
        ctx, C.byref(bp), pcm.data_ptr(), None, pre.data_ptr(), None, stream), "prefix encode"), iters=1)
    res["encode_bit_exact"] = {
        "chains": ch, "trials": trials, "prefix_samples_per_channel": n_pre, "kernel_ms": round(ms, 3),
        "msamples_s": round(ch * n_pre / ms / 1e3, 2),
        "whole_stream_seconds_at_this_rate": round(ms * 1e-3 * n / n_pre, 2),
        "note": "chain-bound: one GPU thread per channel, the reference's block-to-block state carry does not shard"}
    codec = CpuCodec(ch)
    pcm_pre = pcm[0, :, :n_pre].cpu().numpy()
    t0 = time.perf_counter()
    want = codec.encode(pcm_pre, rate, bits, trials)
    res["encode_bit_exact"]["cpu_one_thread_msamples_s"] = round(ch * n_pre / (time.perf_counter() - t0) / 1e6, 2)
    parity["encode_prefix_bit_exact"] = bool(pre[0, :len(want)].cpu().numpy().tobytes() == want)
    parity["against"] = codec.kind
    del pre

    # ---- decode parity: whole stream against the reference (config 3), sampled blocks one by one (config 4)
    data = aad[0, :size].cpu().numpy()
    dec_h = out[0].cpu().numpy()
    def decode_whole_call(capi_lib, reps):
        """AADDecoder_DecodeWhole of `capi_lib` on the stream with plain malloc'd buffers -> (seconds per call, int32 rows)"""
        from aad_b200.capi import _planar_pointers
        blob = np.frombuffer(data.tobytes(), dtype=np.uint8)           # a pageable copy
        rows = [np.zeros(n, dtype=np.int32) for _ in range(ch)]        # touched: page faults are not part of the call
        ptrs = _planar_pointers(rows)
        handle = capi_lib.AADDecoder_Create(None, 0)
        secs = []
        for _ in range(reps):
            t0 = time.perf_counter()
            rc = capi_lib.AADDecoder_DecodeWhole(handle, blob.ctypes.data_as(C.POINTER(C.c_uint8)), len(blob), ptrs, ch, n)
            secs.append(time.perf_counter() - t0)
            check(rc, "AADDecoder_DecodeWhole")
        capi_lib.AADDecoder_Destroy(handle)
        decode_whole_call.last_ms = [round(1e3 * x, 2) for x in secs]
        return min(secs[1:]) if len(secs) > 1 else secs[0], rows

    ref_decode_s = None
    if samples <= 400_000_000 and codec.kind == "reference":
        ref_decode_s, want_rows = decode_whole_call(codec.api.lib, 1)
        parity["decode"] = {"checked": "whole stream", "samples": samples,
                            "bit_exact": bool(all(np.array_equal(want_rows[c], dec_h[c]) for c in range(ch)))}
        del want_rows
    elif samples <= 400_000_000:
        want_pcm = codec.decode(data.tobytes())
        parity["decode"] = {"checked": "whole stream", "samples": samples, "bit_exact": bool(np.array_equal(want_pcm, dec_h))}
        del want_pcm
    else:
        import struct
        rng = np.random.default_rng(4)
        picks = sorted(set([0, 1, nblocks - 2, nblocks - 1] + rng.integers(0, nblocks, size=400).tolist()))
        ok = True
        for blk in picks:
            count = min(spb, n - blk * spb)
            mini = data[:14].tobytes() + struct.pack(">I", count) + data[18:31].tobytes() + \
                data[31 + blk * bs: 31 + (blk + 1) * bs].tobytes()
            ok &= bool(np.array_equal(codec.decode(mini), dec_h[:, blk * spb: blk * spb + count]))
        parity["decode"] = {"checked": f"{len(picks)} blocks spread over the stream, each decoded on its own",
                            "samples": int(sum(min(spb, n - k * spb) for k in picks)) * ch, "bit_exact": bool(ok)}
    # segment-mode encode: sampled segments against the reference's encode of the same samples on a fresh handle
    nseg = -(-nblocks // SEGMENT_BLOCKS)
    rng = np.random.default_rng(9)
    ok = True
    segs = sorted(set([0, nseg - 1] + rng.integers(0, nseg, size=6).tolist()))
    for k in segs:
        s0 = k * SEGMENT_BLOCKS * spb
        seg_pcm = pcm[0, :, s0:s0 + SEGMENT_BLOCKS * spb].cpu().numpy()
        want = codec.encode(seg_pcm, rate, bits, trials)
        b0 = 31 + k * SEGMENT_BLOCKS * bs
        ok &= bool(data[b0:b0 + len(want) - 31].tobytes() == want[31:])
    parity["segments"] = {"checked": len(segs), "bit_exact_per_segment": bool(ok)}

    # ---- end to end from pinned host memory, WAV order: one device, then all local GPUs by block range
    h_aad = gpu.pinned((size,), np.uint8)
    h_wav = gpu.pinned((n, ch), np.int16)
    h_aad[:] = data
    want_wav = np.ascontiguousarray(dec_h.T)
    del dec_h

    def time_host_call(fn, reps=3):
        fn()                                   # warm: first use allocates the context's device buffers
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps

    def link_floor(h2d, d2h):
        if not link:
            return None
        return {"duplex_ms": round(1e3 * max(h2d / (link["h2d"] * 1e9), d2h / (link["d2h"] * 1e9)), 3),
                "serial_ms": round(1e3 * (h2d / (link["h2d"] * 1e9) + d2h / (link["d2h"] * 1e9)), 3)}

    def e2e_entry(seconds, devices, h2d, d2h, same, path):
        e = {"devices": devices, "ms": round(seconds * 1e3, 3), "msamples_s": round(samples / seconds / 1e6, 1),
             "h2d_bytes": h2d, "d2h_bytes": d2h, "host_gbs": round((h2d + d2h) / seconds / 1e9, 2),
             "equals_device_resident_result": bool(same), "path": path}
        if devices == 1:
            f = link_floor(h2d, d2h)
            if f:
                e["link_floor"] = f
                e["vs_link_serial_floor"] = round(seconds * 1e3 / f["serial_ms"], 3)
        return e

    h_wav[...] = 0
    dt = time_host_call(lambda: check(gpu.lib.AADGpu_DecodeInterleaved16(ctx, h_aad.ctypes.data, size, h_wav.ctypes.data, n), "e2e decode"))
    res["decode_e2e"] = e2e_entry(dt, 1, size, samples * 2, np.array_equal(h_wav, want_wav),
                                  "AADGpu_DecodeInterleaved16: pinned .aad in, WAV-order int16 out; sliced H2D | decode | D2H")
    group = None
    if world > 1:
        devs = list(range(world))
        group = gpu.lib.AADGpuGroup_Create((C.c_int * world)(*devs), world)
        if not group:
            raise RuntimeError(f"AADGpuGroup_Create: {gpu.last_error()}")
        h_wav[...] = 0
        dt_g = time_host_call(lambda: check(gpu.lib.AADGpuGroup_DecodeInterleaved16(group, h_aad.ctypes.data, size, h_wav.ctypes.data, n),
                                            "e2e group decode"))
        res["decode_e2e_group"] = e2e_entry(dt_g, world, size, samples * 2, np.array_equal(h_wav, want_wav),
                                            "AADGpuGroup_DecodeInterleaved16: blocks shared out over all local GPUs, one host thread each")
        res["decode_e2e_group"]["speedup_vs_1_device"] = round(dt / dt_g, 3)
        from aad_b200.shard import decode_block_shard      # the same split AADGpuGroup_* makes (aad_gpu.c: split_range)
        res["decode_e2e_group"]["block_shards"] = [
            [sh.block_begin, sh.block_end] for sh in (decode_block_shard(n, spb, bs, size, world, r) for r in range(world))]

    # ---- the drop-in call itself: AADDecoder_DecodeWhole with plain malloc'd buffers, int32 samples (what
    #      src/main.c:94-108 does), next to the reference's own time for the same call on one host core
    if ref_decode_s is not None:
        dt_d, rows = decode_whole_call(api.lib, 5)
        same = all(np.array_equal(rows[c], want_wav[:, c]) for c in range(ch))
        res["dropin_e2e"] = {
            "call": "AADDecoder_DecodeWhole(malloc'd .aad, malloc'd int32 rows), as src/main.c:94-108 calls it",
            "ms": round(dt_d * 1e3, 3), "ms_all_calls": decode_whole_call.last_ms, "msamples_s": round(samples / dt_d / 1e6, 1),
            "equals_device_resident_result": bool(same),
            "vs_pinned_int16_path": round(res["decode_e2e"]["ms"] / (dt_d * 1e3), 3),
            "reference_same_call_ms": round(ref_decode_s * 1e3, 1), "speedup_vs_reference_one_core": round(ref_decode_s / dt_d, 1),
            "how": "slices through a ring of pinned buffers as int16; host threads widen to int32 while they copy"}
        del rows

    # ---- segment-mode encode end to end (extension): WAV-order PCM in pinned memory -> .aad in pinned memory
    h_wav[...] = pcm[0].t().contiguous().cpu().numpy()
    h_out = gpu.pinned((size + 64,), np.uint8)
    osz = C.c_uint32(0)
    check(gpu.lib.AADGpu_SetEncodeSegmentBlocks(ctx, SEGMENT_BLOCKS), "segments on")
    try:
        dt = time_host_call(lambda: check(gpu.lib.AADGpu_EncodeInterleaved16(ctx, C.byref(prm), h_wav.ctypes.data, n, h_out.ctypes.data,
                                                                               size + 64, C.byref(osz)), "e2e segment encode"))
    finally:
        check(gpu.lib.AADGpu_SetEncodeSegmentBlocks(ctx, 0), "segments off")
    res["encode_segment_e2e"] = e2e_entry(dt, 1, samples * 2, size, osz.value == size and np.array_equal(h_out[:size], data),
                                          "AADGpu_EncodeInterleaved16 with 64-block segments (extension)")
    if group:
        h_out[...] = 0
        dt_g = time_host_call(lambda: check(gpu.lib.AADGpuGroup_EncodeInterleaved16(group, C.byref(prm), SEGMENT_BLOCKS, h_wav.ctypes.data, n,
                                                                                      h_out.ctypes.data, size + 64, C.byref(osz)), "e2e group encode"))
        res["encode_segment_e2e_group"] = e2e_entry(dt_g, world, samples * 2, size, osz.value == size and np.array_equal(h_out[:size], data),
                                                    "AADGpuGroup_EncodeInterleaved16: segments shared out over all local GPUs (extension)")
        res["encode_segment_e2e_group"]["speedup_vs_1_device"] = round(dt / dt_g, 3)
        gpu.lib.AADGpuGroup_Destroy(group)
    for a in (h_aad, h_wav, h_out):
        gpu.free_pinned(a)
    parity["e2e_equal_device_resident"] = all(v.get("equals_device_resident_result", True) for v in res.values() if isinstance(v, dict))
    res["parity"] = parity
    del pcm, aad, out
    torch.cuda.empty_cache()
    return res


def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    import aad_b200
    from aad_b200.capi import make_param

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    idle_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        idle_group = dist.new_group(backend="gloo")     # host-side waits that keep the waiting ranks' GPUs idle

    api, gpu = aad_b200.load()
    gpu.lib.AADGpu_SetEncoderSchedule(0 if args.no_pairing else args.schedule)
    ctx = gpu.create(local)
    local_cpus = gpu.lib.AADGpu_BindHostThread(ctx)     # pinned buffers of this rank on the GPU's own NUMA node
    N, ch, n = args.clips, args.channels, args.samples
    prm = make_param(ch, RATE, args.bits, MAX_BLOCK, False, args.trials)
    batch = gpu.batch(N, n, prm)
    astride = int(batch.aad_stream_stride)
    samples_per_step = N * ch * n
    stream_bytes = gpu.stream_bytes(prm, n)
    # algorithmic bytes per sample: int16 PCM once + the encoded stream once
    bytes_per_sample = 2.0 + stream_bytes / (ch * n)
    peak, peak_src = hbm_peak()

    pcm = torch.empty((N, ch, n), dtype=torch.int16, device=dev)
    aad = torch.zeros((N, astride), dtype=torch.uint8, device=dev)
    out = torch.empty((N, ch, n), dtype=torch.int16, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    bref = C.byref(batch)

    def check(rc, what):
        if rc != 0:
            raise RuntimeError(f"{what}: AADApiResult={rc} {gpu.last_error()}")

    # this rank's clips of the whole job (weak scaling: N clips per rank, contiguous ranges, no exchange)
    from aad_b200.shard import decode_block_shard, encode_stream_shard
    first_clip, last_clip = encode_stream_shard(N * world, world, rank)
    assert last_clip - first_clip == N
    check(gpu.lib.AADGpu_SynthBatchDevice(ctx, bref, first_clip, pcm.data_ptr(), stream), "synth")

    def step(events=None):
        if events:
            events[0].record()
        check(gpu.lib.AADGpu_EncodeBatchDevice(ctx, bref, pcm.data_ptr(), None, aad.data_ptr(), None, stream), "encode")
        if events:
            events[1].record()
        check(gpu.lib.AADGpu_DecodeBatchDevice(ctx, bref, aad.data_ptr(), None, out.data_ptr(), stream), "decode")
        if events:
            events[2].record()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    launches0 = gpu.launch_count()
    with ClockSampler(local) as clocks:
        t_start = torch.cuda.Event(enable_timing=True)
        t_end = torch.cuda.Event(enable_timing=True)
        t_start.record()
        for k in range(args.steps):
            step(ev[k])
        t_end.record()
        barrier()
    launches = gpu.launch_count() - launches0
    total_ms = t_start.elapsed_time(t_end)
    enc_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps
    dec_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps
    t = torch.tensor([total_ms, enc_ms, dec_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, enc_ms, dec_ms = (float(x) for x in t.cpu())
    ms_per_step = total_ms / args.steps
    value = samples_per_step * world / (ms_per_step * 1e-3) / 1e6

    # ---- the encoder without the start-state search (num_encode_trials = 0: one pass per block), for
    #      reference next to the CLI-default setting the headline uses --------------------
    trials0 = None
    if args.trials != 0:
        b0 = gpu.batch(N, n, make_param(ch, RATE, args.bits, MAX_BLOCK, False, 0))
        scratch = torch.zeros_like(aad)
        ms0 = _timed_kernel(torch, lambda: check(gpu.lib.AADGpu_EncodeBatchDevice(
            ctx, C.byref(b0), pcm.data_ptr(), None, scratch.data_ptr(), None, stream), "encode t0"), iters=1)
        trials0 = {"kernel_ms": round(ms0, 3), "encode_msamples_s": round(samples_per_step / (ms0 * 1e-3) / 1e6, 3),
                   "hbm_frac": round(samples_per_step * bytes_per_sample / (ms0 * 1e-3) / 1e9 / peak, 5)}
        del scratch

    # ---- the host link, every rank at the same time: what plain pinned 1-D copies reach on this box ----
    link = None
    if not args.no_e2e or not args.no_long:
        gbs = (C.c_double * 3)()
        barrier()
        check(gpu.lib.AADGpu_LinkProbe(ctx, 2 << 30, 2, gbs), "link probe")
        mine = torch.tensor([gbs[0], gbs[1], gbs[2]], dtype=torch.float64, device=dev)
        every = [torch.zeros_like(mine) for _ in range(world)]
        if world > 1:
            dist.all_gather(every, mine)
        else:
            every = [mine]
        every = torch.stack(every).cpu()           # [rank][h2d, d2h, both]
        link = {"h2d": round(gbs[0], 2), "d2h": round(gbs[1], 2), "both": round(gbs[2], 2),
                "sum_over_ranks": {k: round(float(every[:, i].sum()), 2) for i, k in enumerate(("h2d", "d2h", "both"))},
                "slowest_rank": {k: round(float(every[:, i].min()), 2) for i, k in enumerate(("h2d", "d2h", "both"))},
                "per_rank_both": [round(float(x), 1) for x in every[:, 2]],
                "what": "pinned 1-D 2 GiB copies, all ranks at the same time (AADGpu_LinkProbe); rank 0's own rates first; "
                        "bursts of 4 GiB per direction -- the sustained figure for the e2e traffic is e2e.copies_only"}

    # ---- end to end through the host C ABI, pinned host buffers ------------------------------------
    e2e = None
    if not args.no_e2e:
        # pinned host buffers for the whole per-GPU batch (2 x PCM + .aad); if the host cannot pin that much
        # for every rank of this box, the end-to-end leg runs on the largest prefix of the batch that fits
        per_clip = 2 * ch * n * 2 + astride
        Ne = N
        try:
            import psutil
            budget = int(0.6 * psutil.virtual_memory().available / max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world))))
            Ne = max(1, min(N, budget // per_clip))
        except Exception:
            pass
        while True:
            got = []
            try:
                for shape, dt in (((Ne, ch, n), np.int16), ((Ne, astride), np.uint8), ((Ne, ch, n), np.int16)):
                    got.append(gpu.pinned(shape, dt))
                h_pcm, h_aad, h_out = got
                break
            except Exception:
                for a in got:
                    gpu.free_pinned(a)
                if Ne == 1:
                    raise
                Ne = max(1, Ne // 2)
        h_pcm[...] = pcm[:Ne].cpu().numpy()
        hb = gpu.batch(Ne, n, prm)

        def two_calls():
            check(gpu.lib.AADGpu_EncodeBatch(ctx, C.byref(hb), h_pcm.ctypes.data, None, h_aad.ctypes.data, None), "e2e encode")
            check(gpu.lib.AADGpu_DecodeBatch(ctx, C.byref(hb), h_aad.ctypes.data, None, h_out.ctypes.data), "e2e decode")

        def one_call():
            check(gpu.lib.AADGpu_ReconstructBatch(ctx, C.byref(hb), h_pcm.ctypes.data, None, h_aad.ctypes.data, None,
                                                  h_out.ctypes.data), "e2e round trip")

        def copies_only():     # the same host <-> device copies with no kernel between them: the floor on this box
            check(gpu.lib.AADGpu_CopyProbeBatch(ctx, C.byref(hb), h_pcm.ctypes.data, h_aad.ctypes.data, h_out.ctypes.data), "copy probe")

        def timed(fn):
            """-> (seconds per step as max over ranks, results equal to the device-resident ones)"""
            h_aad[...] = 0
            h_out[...] = 0
            fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                fn()
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            same = bool(np.array_equal(h_out, out[:Ne].cpu().numpy())) and \
                bool(np.array_equal(h_aad[:, :stream_bytes], aad[:Ne, :stream_bytes].cpu().numpy()))
            return float(dt.cpu()[0]) / e2e_steps, same

        e2e_steps = max(1, min(args.steps, 3))
        cl = torch.tensor([Ne], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(cl, op=dist.ReduceOp.SUM)
        total_clips = int(cl.cpu()[0])
        total_samples = total_clips * ch * n
        pcm_bytes, aad_bytes = Ne * ch * n * 2, Ne * astride
        s2, same2 = timed(two_calls)
        s1, same1 = timed(one_call)
        check(gpu.lib.AADGpu_ReconstructBatch(ctx, C.byref(hb), h_pcm.ctypes.data, None, h_aad.ctypes.data, None, h_out.ctypes.data),
              "e2e round trip")                       # leaves the device buffers holding the results the probe copies down
        s0, same0 = timed(copies_only)
        same = torch.tensor([int(same1 and same0), int(same2)], dtype=torch.int32, device=dev)
        if world > 1:
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
        same1, same2 = (bool(x) for x in same.cpu())
        # headline: the round trip as ONE call of the public C ABI (the batch form of the reference's
        # execute_reconstruction_core, src/main.c:275-346): PCM in, .aad and reconstructed PCM out
        moved_gbs = total_clips * (2 * ch * n * 2 + astride) / s1 / 1e9      # all ranks, both directions
        e2e = {"value": round(total_samples / s1 / 1e6, 3), "unit": UNIT,
               "h2d_bytes_per_step": pcm_bytes, "d2h_bytes_per_step": aad_bytes + pcm_bytes,
               "steps": e2e_steps, "ms_per_step": round(s1 * 1e3, 3), "clips_per_gpu": Ne,
               "matches_device_resident_result": same1, "host_cpus_bound": int(local_cpus),
               "host_traffic_gbs_all_ranks": round(moved_gbs, 2), "link_probe_gbs": link,
               "copies_only": {"ms_per_step": round(s0 * 1e3, 3),
                               "host_traffic_gbs_all_ranks": round(total_clips * (2 * ch * n * 2 + astride) / s0 / 1e9, 2),
                               "what": "AADGpu_CopyProbeBatch: exactly this call's host <-> device copies (same pinned buffers, slices "
                                       "and row shapes, both directions at once, every rank at the same time), no kernels: the "
                                       "floor of the end-to-end call on this box, max over ranks"},
               "frac_of_link": round(s0 / s1, 4),
               "frac_of_probe_sum": round(moved_gbs / link["sum_over_ranks"]["both"], 4) if link else None,
               "path": "AADGpu_ReconstructBatch (host C ABI, pinned host buffers; per slice H2D pcm | encode | decode | "
                       "D2H .aad + pcm, both link directions busy at once)",
               "separate_calls": {"value": round(total_samples / s2 / 1e6, 3), "ms_per_step": round(s2 * 1e3, 3),
                                  "h2d_bytes_per_step": pcm_bytes + aad_bytes, "d2h_bytes_per_step": aad_bytes + pcm_bytes,
                                  "matches_device_resident_result": same2,
                                  "path": "AADGpu_EncodeBatch then AADGpu_DecodeBatch (the .aad goes to the host and back)"}}
        for a in (h_pcm, h_aad, h_out):
            gpu.free_pinned(a)

    # ---- CPU baseline on a bounded sample + parity of the GPU result: EVERY rank checks clips of its own ------
    cpu = None
    parity = None
    if not args.no_cpu:
        codec = CpuCodec(ch)
        k = min(args.cpu_clips if rank == 0 else 8, N)
        idx = sorted(set(np.linspace(0, N - 1, k).astype(int).tolist()))
        sel = torch.tensor(idx, device=dev)
        pcm_h = pcm[sel].cpu().numpy()
        aad_h = aad[sel].cpu().numpy()
        out_h = out[sel].cpu().numpy()
        secs, ok = 0.0, True
        for j in range(len(idx)):
            data, dec, dt = codec.roundtrip(pcm_h[j], args.bits, args.trials)
            secs += dt
            ok &= (aad_h[j, :len(data)].tobytes() == data) and bool(np.array_equal(out_h[j], dec))
        tally = torch.tensor([int(ok), 1, len(idx)], dtype=torch.int64, device=dev)   # all ok, ranks, streams
        if world > 1:
            lo = tally[:1].clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(tally, op=dist.ReduceOp.SUM)
            tally[0] = lo[0]
        ok_all, ranks, streams = (int(x) for x in tally.cpu())
        cpu = {"value": round(len(idx) * ch * n / secs / 1e6, 3), "unit": UNIT, "cores": 1, "kind": codec.kind,
               "sample": f"{len(idx)} of the {N} clips of rank 0 (evenly spaced), encode+decode, one thread, {secs:.1f} s"}
        parity = {"checked_streams": streams, "checked_ranks": ranks, "world": world, "bit_exact": bool(ok_all),
                  "against": codec.kind, "how": "every rank: evenly spaced clips of its own batch, .aad bytes and decoded PCM "
                                                "against the CPU codec (rank 0: the cpu_baseline sample, other ranks: 8 clips)"}

    # ---- the other shapes of the same kernels, and the long single streams: rank 0, the others wait on the host ----
    del out
    decode_sweep = encode_sweep = None
    other = None
    if rank == 0:
        if not args.no_sweeps:
            del pcm, aad
            torch.cuda.empty_cache()
            decode_sweep, encode_sweep = run_sweeps(args, torch, gpu, ctx, dev, check, samples_per_step, peak)
        if not args.no_long:
            pcm = aad = None
            torch.cuda.empty_cache()
            other = {}
            solo = link
            if world > 1:       # the one-device legs run while the other ranks idle: their link floor is this device alone
                gbs = (C.c_double * 3)()
                check(gpu.lib.AADGpu_LinkProbe(ctx, 2 << 30, 2, gbs), "link probe")
                solo = {"h2d": round(gbs[0], 2), "d2h": round(gbs[1], 2), "both": round(gbs[2], 2)}
            other["link_probe_gbs_one_device_alone"] = {k: solo[k] for k in ("h2d", "d2h", "both")} if solo else None
            for name, spec in LONG_STREAMS.items():
                other[name] = run_long_stream(name, spec, args, torch, api, gpu, ctx, dev, check, world, peak, solo)
    else:
        del pcm, aad
        torch.cuda.empty_cache()
    if idle_group is not None:
        dist.barrier(group=idle_group)

    if rank == 0:
        default_workload = (N, n, ch, args.bits, args.trials) == (12500, CLIP_SAMPLES, CHANNELS, BITS, TRIALS)
        enc_gbs = samples_per_step * bytes_per_sample / (enc_ms * 1e-3) / 1e9
        dec_gbs = samples_per_step * bytes_per_sample / (dec_ms * 1e-3) / 1e9
        dominant, dom_gbs = ("aad_encode", enc_gbs) if enc_ms >= dec_ms else ("aad_decode", dec_gbs)
        line = {
            "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": workload_config(args, world),
            "encode_msamples_s": round(samples_per_step * world / (enc_ms * 1e-3) / 1e6, 3),
            "decode_msamples_s": round(samples_per_step * world / (dec_ms * 1e-3) / 1e6, 3),
            "kernel_ms": {"encode": round(enc_ms, 3), "decode": round(dec_ms, 3)},
            "roofline": {"bound": "hbm", "kernel": dominant, "achieved": round(dom_gbs, 2), "peak": peak, "unit": "GB/s",
                         "frac": round(dom_gbs / peak, 5), "traffic": None,
                         "traffic_from_profile": TRAFFIC_FROM_PROFILE if default_workload and dominant == "aad_encode" else None,
                         "algorithmic_bytes_per_launch": round(samples_per_step * bytes_per_sample), "peak_source": peak_src,
                         "algorithmic_bytes_per_sample": round(bytes_per_sample, 4),
                         "note": "12,500 serial chains per GPU: latency/issue bound, not HBM bound (DESIGN.md 4.2)"},
            "roofline_decode": {"bound": "hbm", "kernel": "aad_decode", "achieved": round(dec_gbs, 2), "peak": peak,
                                "unit": "GB/s", "frac": round(dec_gbs / peak, 5)},
            "encode_without_search": trials0, "decode_sweep": decode_sweep, "encode_sweep": encode_sweep,
            "cpu_baseline": cpu, "parity": parity, "e2e": e2e, "other_configs": other,
            "gpu_launches": launches, "clocks": clocks.summary(),
        }
        print(json.dumps(line), flush=True)
    gpu.destroy(ctx)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
