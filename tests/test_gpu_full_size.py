"""BASELINE configs 3 and 4 at their FULL sizes on the device.

config 3  synthetic 1-hour 48 kHz stereo stream, 4-bit: 172.8 M samples per channel, 174,194 blocks
config 4  synthetic 30-minute 96 kHz 8-channel stream, 3-bit: 172.8 M samples per channel, 591,781 blocks

One stream is a serial encode chain per channel (src/aad_encoder.c:853-886), so these encode as 2 / 8
GPU threads (flagged chain-bound in DESIGN.md) and run here with 0 trials to keep the test short;
their DECODE is block parallel and shards by block range.  Checks:
  config 3: whole-stream bit-exactness of encode AND decode against the CPU oracle, sharded decode.
  config 4: size-independent properties -- the first blocks against the oracle's encode of the same
            prefix, blocks sampled over the whole stream decoded one by one by the oracle, sharded
            decode == unsharded decode, round-trip error within the reference's own tolerance.
"""
import ctypes as C
import struct
import time

import numpy as np
import pytest

import aadtest
from aad_b200.capi import OK, make_param

pytestmark = pytest.mark.gpu


def run_stream(gpu, ctx, channels, rate, bits, n, trials=0):
    """synthesise, encode and decode ONE stream on the device; returns host arrays and timings"""
    import torch
    prm = make_param(channels, rate, bits, 1024, False, trials)
    b = gpu.batch(1, n, prm)
    dev = torch.device("cuda:0")
    pcm = torch.zeros((1, channels, n), dtype=torch.int16, device=dev)
    aad = torch.zeros((1, b.aad_stream_stride), dtype=torch.uint8, device=dev)
    out = torch.zeros_like(pcm)
    s = torch.cuda.current_stream().cuda_stream
    assert gpu.lib.AADGpu_SynthBatchDevice(ctx, C.byref(b), 7, pcm.data_ptr(), s) == OK
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    assert gpu.lib.AADGpu_EncodeBatchDevice(ctx, C.byref(b), pcm.data_ptr(), None, aad.data_ptr(), None, s) == OK
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    assert gpu.lib.AADGpu_DecodeBatchDevice(ctx, C.byref(b), aad.data_ptr(), None, out.data_ptr(), s) == OK
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    size = gpu.stream_bytes(prm, n)
    res = pcm[0].cpu().numpy(), aad[0, :size].cpu().numpy(), out[0].cpu().numpy(), (t1 - t0, t2 - t1)
    del pcm, aad, out
    torch.cuda.empty_cache()
    return res


def sharded_decode(gpu, data, channels, n, devices):
    g = gpu.lib.AADGpuGroup_Create((C.c_int * len(devices))(*devices), len(devices))
    assert g, gpu.last_error()
    try:
        out = gpu.pinned((n, channels), np.int16)
        blob = gpu.pinned((len(data),), np.uint8)
        blob[:] = data
        t0 = time.perf_counter()
        rc = gpu.lib.AADGpuGroup_DecodeInterleaved16(g, blob.ctypes.data, len(blob), out.ctypes.data, n)
        dt = time.perf_counter() - t0
        assert rc == OK, gpu.last_error()
        res = np.ascontiguousarray(out.T)
        gpu.free_pinned(out)
        gpu.free_pinned(blob)
        return res, dt
    finally:
        gpu.lib.AADGpuGroup_Destroy(g)


def device_list(gpu, shards):
    n = gpu.device_count()
    return [i % n for i in range(shards)]


def test_config3_one_hour_stereo_is_bit_exact_end_to_end(product, gpu_ctx, oracle):
    _, gpu = product
    ch, rate, bits, n = 2, 48000, 4, 172_800_000
    pcm, data, dec, (t_enc, t_dec) = run_stream(gpu, gpu_ctx, ch, rate, bits, n)
    print(f"\nconfig 3 on one B200: encode (2 chains, 0 trials) {t_enc:.2f} s = {ch * n / t_enc / 1e6:.0f} Msamples/s, "
          f"decode {t_dec * 1e3:.1f} ms = {ch * n / t_dec / 1e6:.0f} Msamples/s")
    t0 = time.perf_counter()
    rc, want = oracle.encode(pcm, rate, bits, 1024, False, 0)
    t_cpu = time.perf_counter() - t0
    assert rc == 0 and len(want) == len(data) == 31 + 174_193 * 1024 + aad_tail_bytes(n, 992, ch, bits)
    assert data.tobytes() == want                                        # 178 MB, byte for byte
    rc, want_pcm, _ = oracle.decode(want)
    assert rc == 0 and np.array_equal(dec, want_pcm)                     # 345.6 M samples
    print(f"config 3 oracle encode on one host core: {t_cpu:.1f} s = {ch * n / t_cpu / 1e6:.0f} Msamples/s")
    for shards in (4, 7):
        got, dt = sharded_decode(gpu, data, ch, n, device_list(gpu, shards))
        assert np.array_equal(got, dec), shards
        print(f"config 3 decode from host memory in {shards} block-range shards on {gpu.device_count()} GPU(s): {dt * 1e3:.0f} ms")


def aad_tail_bytes(n, spb, ch, bits):
    """bytes of the last, partial block (0 when the stream ends on a block boundary)"""
    tail = n % spb
    if tail == 0:
        return 0
    gs, gb = {4: (2, 1), 3: (8, 3), 2: (4, 1)}[bits]
    groups = (tail - 4 + gs - 1) // gs if tail > 4 else 0
    return ch * (18 + groups * gb)


def test_config4_thirty_minutes_eight_channels_properties(product, gpu_ctx, oracle):
    _, gpu = product
    ch, rate, bits, n = 8, 96000, 3, 172_800_000
    spb, bs = 292, 1008
    pcm, data, dec, (t_enc, t_dec) = run_stream(gpu, gpu_ctx, ch, rate, bits, n)
    print(f"\nconfig 4 on one B200: encode (8 chains, 0 trials) {t_enc:.2f} s = {ch * n / t_enc / 1e6:.0f} Msamples/s, "
          f"decode {t_dec * 1e3:.1f} ms = {ch * n / t_dec / 1e6:.0f} Msamples/s")
    nblocks = (n + spb - 1) // spb
    assert nblocks == 591_781 and struct.unpack(">HIIHHI", data[12:30].tobytes()) == (ch, n, rate, bits, bs, spb)
    # (1) the encoder is a serial chain: its first K blocks depend on the first K*spb samples only
    k = 4000
    rc, want = oracle.encode(pcm[:, :k * spb], rate, bits, 1024, False, 0)
    assert rc == 0 and data[31:31 + k * bs].tobytes() == want[31:]
    # (2) every block decodes on its own: sampled blocks through the oracle, one block per call
    rng = np.random.default_rng(4)
    picks = sorted(set([0, 1, nblocks - 2, nblocks - 1] + rng.integers(0, nblocks, size=400).tolist()))
    for b in picks:
        count = min(spb, n - b * spb)
        blk = data[31 + b * bs: 31 + (b + 1) * bs].tobytes()
        hdr = data[:14].tobytes() + struct.pack(">I", count) + data[18:31].tobytes()
        rc, want_pcm, _ = oracle.decode(hdr + blk)
        assert rc == 0 and np.array_equal(dec[:, b * spb: b * spb + count], want_pcm), b
    # (3) block-range shards reassemble to the unsharded result
    got, dt = sharded_decode(gpu, data, ch, n, device_list(gpu, 5))
    assert np.array_equal(got, dec)
    print(f"config 4 decode from host memory in 5 block-range shards on {gpu.device_count()} GPU(s): {dt * 1e3:.0f} ms")
    # (4) round trip within the reference's 3-bit tolerance for tonal signals (test/test_aad_encode_decode.c:310-315: RMSE / 32767 < 6e-2)
    step = 1 << 20
    err2 = sum(float(np.sum((pcm[:, i:i + step].astype(np.float64) - dec[:, i:i + step]) ** 2)) for i in range(0, n, step))
    assert np.sqrt(err2 / (ch * n)) / 32767.0 < 6.0e-2


@pytest.mark.parametrize("config", ["config3", "config4"])
def test_long_streams_encode_in_parallel_with_segments(product, gpu_ctx, oracle, config):
    """The extension that makes ONE long stream a parallel encode (AADGpu_SetEncodeSegmentBlocks; not byte-identical
    to the reference encoder, tests/test_gpu_segments.py): configs 3 and 4 at full size with the CLI's 2 trials.
    Sampled segments are compared with the oracle's encode of the same samples on a fresh handle, the whole stream
    is decoded (block-parallel) and the round trip held to the reference's own tolerance."""
    _, gpu = product
    ch, rate, bits, n, spb, bs = (2, 48000, 4, 172_800_000, 992, 1024) if config == "config3" else (8, 96000, 3, 172_800_000, 292, 1008)
    seg_blocks, trials = 64, 2
    assert gpu.lib.AADGpu_SetEncodeSegmentBlocks(gpu_ctx, seg_blocks) == OK
    try:
        pcm, data, dec, (t_enc, t_dec) = run_stream(gpu, gpu_ctx, ch, rate, bits, n, trials=trials)
    finally:
        gpu.lib.AADGpu_SetEncodeSegmentBlocks(gpu_ctx, 0)
    nblocks = (n + spb - 1) // spb
    nseg = (nblocks + seg_blocks - 1) // seg_blocks
    print(f"\n{config} on one B200, segments of {seg_blocks} blocks ({nseg * ch} chains), {trials} trials: encode {t_enc * 1e3:.1f} ms = "
          f"{ch * n / t_enc / 1e6:.0f} Msamples/s, decode {t_dec * 1e3:.1f} ms")
    rng = np.random.default_rng(9)
    for k in sorted(set([0, 1, nseg - 2, nseg - 1] + rng.integers(0, nseg, size=60).tolist())):
        s0 = k * seg_blocks * spb
        rc, want = oracle.encode(pcm[:, s0:s0 + seg_blocks * spb], rate, bits, 1024, False, trials)
        assert rc == 0
        b0 = 31 + k * seg_blocks * bs
        assert data[b0:b0 + len(want) - 31].tobytes() == want[31:], k
    assert len(data) == 31 + (nblocks - 1) * bs + (aad_tail_bytes(n, spb, ch, bits) or bs)
    step = 1 << 20
    err2 = sum(float(np.sum((pcm[:, i:i + step].astype(np.float64) - dec[:, i:i + step]) ** 2)) for i in range(0, n, step))
    assert np.sqrt(err2 / (ch * n)) / 32767.0 < 6.0e-2
    if config == "config3":
        # the same stream from host memory (WAV order), its segments shared out over a device group: what
        # `aad -e --segment-blocks 64 --device ...` does; identical bytes to the one-device encode above
        prm = make_param(ch, rate, bits, 1024, False, trials)
        inter = gpu.pinned((n, ch), np.int16)
        inter[:] = pcm.T
        blob = gpu.pinned((len(data) + 64,), np.uint8)
        for shards in (1, 4):
            devices = device_list(gpu, shards)
            g = gpu.lib.AADGpuGroup_Create((C.c_int * len(devices))(*devices), len(devices))
            assert g, gpu.last_error()
            try:
                size = C.c_uint32(0)
                t0 = time.perf_counter()
                rc = gpu.lib.AADGpuGroup_EncodeInterleaved16(g, C.byref(prm), seg_blocks, inter.ctypes.data, n, blob.ctypes.data,
                                                             len(blob), C.byref(size))
                dt = time.perf_counter() - t0
                assert rc == OK, gpu.last_error()
                assert size.value == len(data) and np.array_equal(blob[:size.value], data), shards
                print(f"config3 encode from host memory in {shards} segment-range shard(s) on {gpu.device_count()} GPU(s): {dt * 1e3:.0f} ms "
                      f"= {ch * n / dt / 1e6:.0f} Msamples/s end to end")
            finally:
                gpu.lib.AADGpuGroup_Destroy(g)
        gpu.free_pinned(inter)
        gpu.free_pinned(blob)
