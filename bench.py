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
            AADGpu_EncodeBatch / AADGpu_DecodeBatch (e2e.separate_calls)
  roofline  dominant kernel (the encoder at 2 trials) against the measured HBM copy bandwidth
  cpu_baseline  the reference codec on this box's host cores, bounded sample, same clips;
            its output doubles as the parity check of the GPU result

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
# ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel
# (aad_encode_fast<4,0>) at the default workload: profiles/r01_v6_encode.md.  Reported only for
# that exact workload.
NCU_TRAFFIC_BYTES = 23.5e9


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
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-clips", type=int, default=96, help="clips in the bounded CPU-baseline sample")
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

    def __init__(self):
        from aad_b200.capi import AADCApi
        ref = ROOT / "oracle" / "_ref" / "libaad_ref.so"
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
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    api, gpu = aad_b200.load()
    gpu.lib.AADGpu_SetEncoderPairing(0 if args.no_pairing else 1)
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

    pcm = torch.empty((N, ch, n), dtype=torch.int16, device=dev)
    aad = torch.zeros((N, astride), dtype=torch.uint8, device=dev)
    out = torch.empty((N, ch, n), dtype=torch.int16, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    bref = C.byref(batch)

    def check(rc, what):
        if rc != 0:
            raise RuntimeError(f"{what}: AADApiResult={rc} {gpu.last_error()}")

    check(gpu.lib.AADGpu_SynthBatchDevice(ctx, bref, rank * N, pcm.data_ptr(), stream), "synth")

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
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for k in range(2):      # first launch warms up
            evs[0].record()
            check(gpu.lib.AADGpu_EncodeBatchDevice(ctx, C.byref(b0), pcm.data_ptr(), None, scratch.data_ptr(), None, stream), "encode t0")
            evs[1].record()
        torch.cuda.synchronize()
        ms0 = evs[0].elapsed_time(evs[1])
        trials0 = {"kernel_ms": round(ms0, 3), "encode_msamples_s": round(samples_per_step / (ms0 * 1e-3) / 1e6, 3),
                   "hbm_frac": round(samples_per_step * bytes_per_sample / (ms0 * 1e-3) / 1e9 / hbm_peak()[0], 5)}
        del scratch

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
        total_samples = int(cl.cpu()[0]) * ch * n
        pcm_bytes, aad_bytes = Ne * ch * n * 2, Ne * astride
        s2, same2 = timed(two_calls)
        s1, same1 = timed(one_call)
        # headline: the round trip as ONE call of the public C ABI (the batch form of the reference's
        # execute_reconstruction_core, src/main.c:275-346): PCM in, .aad and reconstructed PCM out
        e2e = {"value": round(total_samples / s1 / 1e6, 3), "unit": UNIT,
               "h2d_bytes_per_step": pcm_bytes, "d2h_bytes_per_step": aad_bytes + pcm_bytes,
               "steps": e2e_steps, "ms_per_step": round(s1 * 1e3, 3), "clips_per_gpu": Ne,
               "matches_device_resident_result": same1, "host_cpus_bound": int(local_cpus),
               "path": "AADGpu_ReconstructBatch (host C ABI, pinned host buffers; per slice H2D pcm | encode | decode | "
                       "D2H .aad + pcm, both link directions busy at once)",
               "separate_calls": {"value": round(total_samples / s2 / 1e6, 3), "ms_per_step": round(s2 * 1e3, 3),
                                  "h2d_bytes_per_step": pcm_bytes + aad_bytes, "d2h_bytes_per_step": aad_bytes + pcm_bytes,
                                  "matches_device_resident_result": same2,
                                  "path": "AADGpu_EncodeBatch then AADGpu_DecodeBatch (the .aad goes to the host and back)"}}
        for a in (h_pcm, h_aad, h_out):
            gpu.free_pinned(a)

    # ---- CPU baseline on a bounded sample + parity of the GPU result on those clips -----------------
    cpu = None
    parity = None
    if rank == 0 and not args.no_cpu:
        codec = CpuCodec()
        k = min(args.cpu_clips, N)
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
        cpu = {"value": round(len(idx) * ch * n / secs / 1e6, 3), "unit": UNIT, "cores": 1, "kind": codec.kind,
               "sample": f"{len(idx)} of the {N} clips of rank 0 (evenly spaced), encode+decode, one thread, {secs:.1f} s"}
        parity = {"checked_streams": len(idx), "bit_exact": bool(ok), "against": codec.kind}

    if rank == 0:
        peak, peak_src = hbm_peak()
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
                         "frac": round(dom_gbs / peak, 5),
                         "traffic": NCU_TRAFFIC_BYTES if default_workload and dominant == "aad_encode" else None,
                         "algorithmic_bytes_per_launch": round(samples_per_step * bytes_per_sample), "peak_source": peak_src,
                         "algorithmic_bytes_per_sample": round(bytes_per_sample, 4),
                         "note": "12,500 serial chains per GPU: latency/issue bound, not HBM bound (DESIGN.md 4.2, profiles/r01_v6_encode.md)"},
            "roofline_decode": {"bound": "hbm", "kernel": "aad_decode", "achieved": round(dec_gbs, 2), "peak": peak,
                                "unit": "GB/s", "frac": round(dec_gbs / peak, 5)},
            "encode_without_search": trials0,
            "cpu_baseline": cpu, "parity": parity, "e2e": e2e, "gpu_launches": launches, "clocks": clocks.summary(),
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
