"""One stream decoded straight into WAV order (aadk_decode_params::interleaved): the decoder's flush writes
whole frames for mono and 2 / 4 / 8 channels, other shapes decode to planes and take one more pass.  All of
them through the public one-stream entry points (AADGpu_DecodeInterleaved16, AADGpuGroup_DecodeInterleaved16,
AADGpu_ReconstructInterleaved16) against the oracle's whole-stream decode, bit for bit; plus the round-2
hardening of the decoder kernels (header sample count capped by the batch, corrupt step index clamped on every
kernel path, truncated streams report INSUFFICIENT_DATA like src/aad_decoder.c:522-527).
"""
import ctypes as C

import numpy as np
import pytest

import aadtest
from aad_b200 import capi
from aad_b200.capi import OK, make_param

pytestmark = pytest.mark.gpu


def _decode_interleaved(gpu, ctx, data, n, channels, fill=12345):
    blob = np.frombuffer(bytes(data), dtype=np.uint8).copy()
    out = np.full((max(n, 1), channels), fill, dtype=np.int16)
    rc = gpu.lib.AADGpu_DecodeInterleaved16(ctx, blob.ctypes.data, len(blob), out.ctypes.data, n)
    return rc, out


@pytest.mark.parametrize("bits", [2, 3, 4])
@pytest.mark.parametrize("channels,ms", [(1, False), (2, False), (2, True), (3, False), (4, False), (4, True), (5, True),
                                         (8, False)])
def test_wav_order_decode_matches_oracle(product, gpu_ctx, oracle, bits, channels, ms):
    _, gpu = product
    for block, n in ((1024, 70001), (1024, 5), (256 * channels, 40000), (1024, 2 * 4028 + 4), (4096, 33333)):
        pcm = aadtest.signal(("music", "noise", "steps")[(bits + channels + n) % 3], channels, n, bits * 31 + channels)
        rc, data = oracle.encode(pcm, 48000, bits, block, ms, 1)
        assert rc == 0
        _, bs, spb = oracle.geometry(block, channels, bits)
        # whole stream, cut inside a block's codes, cut inside a block's channel headers, cut at a block edge
        cuts = [len(data), len(data) - (len(data) - 31) // 3, 31 + bs * ((len(data) - 31) // bs // 2) + 7,
                31 + bs * ((len(data) - 31) // bs // 2)]
        for cut in cuts:
            if cut < 31:
                continue
            want_rc, want, _ = oracle.decode(data[:cut], fill=0)
            rc, out = _decode_interleaved(gpu, gpu_ctx, data[:cut], n, channels)
            assert rc == want_rc, (bits, channels, ms, block, n, cut, gpu.last_error())
            assert np.array_equal(out.T, want), (bits, channels, ms, block, n, cut)


def test_tma_bulk_store_flush_is_bit_exact(product, gpu_ctx, oracle):
    """kernel path 4: mono 4-bit rows leave shared memory through cp.async.bulk (measured slower, off by default):
    ragged batch, truncated streams, against the default path and the oracle"""
    _, gpu = product
    rng = np.random.default_rng(77)
    n_streams, n_max = 70, 30000
    lens = rng.integers(1, n_max + 1, size=n_streams).astype(np.uint32)
    lens[0], lens[1], lens[2] = n_max, 4, 2016 * 3
    pcm = np.zeros((n_streams, 1, n_max), dtype=np.int16)
    for i in range(n_streams):
        pcm[i, :, :lens[i]] = aadtest.signal(aadtest.SIGNALS[i % len(aadtest.SIGNALS)], 1, int(lens[i]), i)
    aad, sizes = gpu.encode_batch(gpu_ctx, pcm, 44100, 4, 1024, False, 0, num_samples=lens)
    cut = sizes.copy()
    for i in range(3, n_streams, 5):
        cut[i] = max(31, int(sizes[i]) - int(rng.integers(0, 2048)))
    res = []
    for path in (0, 4):
        gpu.lib.AADGpu_SetKernelPath(path)
        try:
            res.append((gpu.decode_batch(gpu_ctx, aad, n_max, 44100, 1, 4, 1024, False, sizes=sizes),
                        gpu.decode_batch(gpu_ctx, aad, n_max, 44100, 1, 4, 1024, False, sizes=cut)))
        finally:
            gpu.lib.AADGpu_SetKernelPath(0)
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    for i in range(0, n_streams, 9):
        _, want, _ = oracle.decode(aad[i, :sizes[i]].tobytes())
        assert np.array_equal(res[1][0][i, :, :lens[i]], want), i


@pytest.mark.parametrize("bits", [4, 2])
def test_tma_tensor_map_staging_is_bit_exact(product, gpu_ctx, oracle, bits):
    """kernel paths 7 / 8: mono 4-bit / 2-bit blocks staged by the TMA unit through a tensor map (aad_decode_tma; 8 = 24
    warps per SM, output rows flushed twice per window): ragged
    batch (stream lengths from the headers, last blocks partial, streams shorter than a warp task), several block
    sizes, against the default path and the oracle; the launch counter proves the TMA kernel really ran"""
    _, gpu = product
    rng = np.random.default_rng(78 + bits)
    for block, n_streams, n_max in ((1024, 70, 150000), (256, 9, 20000), (1024, 3, 500)):
        lens = rng.integers(1, n_max + 1, size=n_streams).astype(np.uint32)
        lens[0] = n_max
        lens[1] = 4
        pcm = np.zeros((n_streams, 1, n_max), dtype=np.int16)
        for i in range(n_streams):
            pcm[i, :, :lens[i]] = aadtest.signal(aadtest.SIGNALS[i % len(aadtest.SIGNALS)], 1, int(lens[i]), i)
        aad, sizes = gpu.encode_batch(gpu_ctx, pcm, 44100, bits, block, False, 0, num_samples=lens)
        res = []
        for path in (0, 7, 8):
            gpu.lib.AADGpu_SetKernelPath(path)
            before = int(gpu.lib.AADGpu_TmaLaunchCount())
            try:
                res.append(gpu.decode_batch(gpu_ctx, aad, n_max, 44100, 1, bits, block, False))
            finally:
                gpu.lib.AADGpu_SetKernelPath(0)
            assert (int(gpu.lib.AADGpu_TmaLaunchCount()) > before) == (path != 0), (path, block)
        assert np.array_equal(res[0], res[1]) and np.array_equal(res[0], res[2]), (bits, block)
        for i in range(0, n_streams, 9):
            _, want, _ = oracle.decode(aad[i, :sizes[i]].tobytes())
            assert np.array_equal(res[1][i, :, :lens[i]], want), (bits, block, i)


@pytest.mark.parametrize("path", [0, 1, 2])
def test_wav_order_decode_on_every_kernel_path(product, gpu_ctx, oracle, path):
    """kernel path 1 (generic kernels) has no WAV-order flush: planes + one interleave pass, same samples"""
    _, gpu = product
    gpu.lib.AADGpu_SetKernelPath(path)
    try:
        for channels, bits, ms in ((2, 4, True), (8, 3, False), (1, 2, False)):
            pcm = aadtest.signal("music", channels, 50000, 7)
            _, data = oracle.encode(pcm, 44100, bits, 1024, ms, 0)
            _, want, _ = oracle.decode(data)
            rc, out = _decode_interleaved(gpu, gpu_ctx, data, 50000, channels)
            assert rc == OK and np.array_equal(out.T, want), (path, channels, bits)
    finally:
        gpu.lib.AADGpu_SetKernelPath(0)


def test_reconstruct_one_stream_in_wav_order(product, gpu_ctx, oracle):
    _, gpu = product
    for channels, bits, ms in ((2, 4, False), (2, 3, True), (8, 3, False), (4, 2, False), (6, 4, False)):
        n = 123457
        pcm = aadtest.signal("music", channels, n, channels)
        prm = make_param(channels, 48000, bits, 1024, ms, 1)
        inter = np.ascontiguousarray(pcm.T)
        out = np.zeros_like(inter)
        size = C.c_uint32(0)
        rc = gpu.lib.AADGpu_ReconstructInterleaved16(gpu_ctx, C.byref(prm), inter.ctypes.data, n, out.ctypes.data, C.byref(size))
        assert rc == OK, gpu.last_error()
        _, data = oracle.encode(pcm, 48000, bits, 1024, ms, 1)
        _, want, _ = oracle.decode(data)
        assert size.value == len(data) and np.array_equal(out.T, want), (channels, bits, ms)


def test_header_cannot_claim_more_samples_than_the_batch_rows_hold(product, gpu_ctx, oracle):
    """AADGpu_DecodeBatchDevice takes each stream's length from its own header on the device; a header that
    claims more than batch->num_samples must not write past the rows the batch describes."""
    import torch
    _, gpu = product
    for channels, bits in ((1, 4), (2, 4), (8, 3)):
        n, n_rows = 9000, 5000
        n_streams = 3
        prm = make_param(channels, 44100, bits, 1024, False, 0)
        pcm = np.stack([aadtest.signal("music", channels, n, i) for i in range(n_streams)])
        streams = [oracle.encode(pcm[i], 44100, bits, 1024, False, 0)[1] for i in range(n_streams)]
        stride = max(len(s) for s in streams)
        aad_h = np.zeros((n_streams, stride), dtype=np.uint8)
        for i, s in enumerate(streams):
            aad_h[i, :len(s)] = np.frombuffer(s, dtype=np.uint8)
        b = gpu.batch(n_streams, n_rows, prm, aad_stream_stride=stride)      # rows hold 5000 samples, headers say 9000
        for kernel_path in (0, 1):
            gpu.lib.AADGpu_SetKernelPath(kernel_path)
            try:
                dev = torch.device("cuda:0")
                aad = torch.from_numpy(aad_h).to(dev)
                guard = 64
                out = torch.full((n_streams * channels * n_rows + guard,), -7, dtype=torch.int16, device=dev)
                s = torch.cuda.current_stream().cuda_stream
                assert gpu.lib.AADGpu_DecodeBatchDevice(gpu_ctx, C.byref(b), aad.data_ptr(), None, out.data_ptr(), s) == OK
                torch.cuda.synchronize()
                got = out.cpu().numpy()
            finally:
                gpu.lib.AADGpu_SetKernelPath(0)
            assert np.all(got[-guard:] == -7), (channels, bits, kernel_path)
            rows = got[:-guard].reshape(n_streams, channels, n_rows)
            for i in range(n_streams):
                _, want, _ = oracle.decode(streams[i])
                assert np.array_equal(rows[i], want[:, :n_rows]), (channels, bits, kernel_path, i)


@pytest.mark.parametrize("channels,bits", [(1, 4), (2, 3), (8, 3)])
def test_corrupt_step_index_is_clamped_on_every_kernel_path(product, gpu_ctx, oracle, channels, bits):
    """A block header's step index is 12 bits on the wire (up to 4095) but the table ends at 4080
    (src/aad_tables.h:38-39): all three decoder kernels clamp it, so they agree on a corrupt stream."""
    _, gpu = product
    n = 30000
    pcm = aadtest.signal("music", channels, n, 5)
    _, data = oracle.encode(pcm, 44100, bits, 1024, False, 0)
    bad = bytearray(data)
    _, bs, _ = oracle.geometry(1024, channels, bits)
    for b in (0, 2, 5):                       # index 4095, shift kept
        for c in range(channels):
            off = 31 + b * bs + 18 * c
            bad[off] = 0xFF
            bad[off + 1] = 0xF0 | (bad[off + 1] & 0x0F)
    res = []
    bound = gpu.stream_bytes_bound(make_param(channels, 44100, bits, 1024, False, 0), n)
    aad = np.zeros((1, bound), dtype=np.uint8)
    aad[0, :len(bad)] = np.frombuffer(bytes(bad), dtype=np.uint8)
    sizes = np.array([len(bad)], dtype=np.uint32)
    for path in (0, 1, 2):
        gpu.lib.AADGpu_SetKernelPath(path)
        try:
            res.append(gpu.decode_batch(gpu_ctx, aad, n, 44100, channels, bits, 1024, False, sizes=sizes))
        finally:
            gpu.lib.AADGpu_SetKernelPath(0)
    assert np.array_equal(res[0], res[1]) and np.array_equal(res[0], res[2])


def test_group_over_distinct_devices(product, oracle):
    """the SCALE box has 8 GPUs: one stream's blocks shared out over all of them (skipped on a 1-GPU box)"""
    _, gpu = product
    ndev = gpu.device_count()
    if ndev < 2:
        pytest.skip("one visible device: distinct-device groups are covered on the multi-GPU box")
    devices = list(range(min(ndev, 8)))
    g = gpu.lib.AADGpuGroup_Create((C.c_int * len(devices))(*devices), len(devices))
    assert g, gpu.last_error()
    try:
        for channels, bits in ((2, 4), (8, 3)):
            n = 900001
            pcm = aadtest.signal("music", channels, n, 9)
            _, data = oracle.encode(pcm, 48000, bits, 1024, False, 0)
            _, want, _ = oracle.decode(data)
            blob = np.frombuffer(data, dtype=np.uint8).copy()
            out = np.zeros((n, channels), dtype=np.int16)
            assert gpu.lib.AADGpuGroup_DecodeInterleaved16(g, blob.ctypes.data, len(blob), out.ctypes.data, n) == OK, gpu.last_error()
            assert np.array_equal(out.T, want), (channels, bits)
    finally:
        gpu.lib.AADGpuGroup_Destroy(g)
