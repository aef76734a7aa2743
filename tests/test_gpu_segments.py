"""Segment-parallel encoding (AADGpu_SetEncodeSegmentBlocks) -- GPU only.

An extension, NOT byte-identical to the reference encoder: every run of S blocks is
encoded as a stream of its own.  What IS checked bit for bit:
  * each segment's bytes == the oracle's (= the reference encoder's) output for that run of samples on a
    fresh handle, file header stripped;
  * the stream decodes with the stock decoder (the oracle) and this library's decoder to the same PCM;
  * host pipelines that cut the work into block-range slices (slice edges inside segments) produce the
    same bytes as one device-resident launch;
  * with the setting back at 0 the output is again the reference's, byte for byte.
"""
import ctypes as C
import subprocess

import numpy as np
import pytest

import aadtest
from aad_b200.capi import OK, make_param

pytestmark = pytest.mark.gpu


def _segmented_reference(oracle, pcm, rate, bits, block, ms, trials, seg_blocks):
    """the stream the segment mode must produce: file header of the whole stream + the oracle's blocks of
    every run of seg_blocks blocks encoded on a fresh handle"""
    ch, n = pcm.shape
    _, bs, spb = oracle.geometry(block, ch, bits)
    rc, whole = oracle.encode(pcm, rate, bits, block, ms, 0)
    assert rc == 0
    parts = [whole[:31]]
    for s0 in range(0, n, seg_blocks * spb):
        rc, seg = oracle.encode(pcm[:, s0:s0 + seg_blocks * spb], rate, bits, block, ms, trials)
        assert rc == 0
        parts.append(seg[31:])
    return b"".join(parts)


@pytest.mark.parametrize("bits", [2, 3, 4])
@pytest.mark.parametrize("channels,ms", [(1, False), (2, True), (8, False)])
def test_segments_equal_fresh_handle_encodes(product, gpu_ctx, oracle, bits, channels, ms):
    _, gpu = product
    n_streams, n_max = 23, 9000
    rng = np.random.default_rng(40 + bits + 10 * channels)
    lens = rng.integers(1, n_max + 1, size=n_streams).astype(np.uint32)
    lens[0], lens[1], lens[2] = n_max, 3, 5
    pcm = np.zeros((n_streams, channels, n_max), dtype=np.int16)
    for i in range(n_streams):
        pcm[i, :, :lens[i]] = aadtest.signal(aadtest.SIGNALS[i % len(aadtest.SIGNALS)], channels, int(lens[i]), 11 * i)
    for block, seg_blocks, trials in ((256 * channels // (2 if channels == 8 else 1), 3, 2), (1024, 1, 1), (1024, 2, 0), (1024, 1000, 2)):
        assert gpu.lib.AADGpu_SetEncodeSegmentBlocks(gpu_ctx, seg_blocks) == OK
        try:
            assert gpu.lib.AADGpu_GetEncodeSegmentBlocks(gpu_ctx) == seg_blocks
            aad, sizes = gpu.encode_batch(gpu_ctx, pcm, 44100, bits, block, ms, trials, num_samples=lens)
        finally:
            gpu.lib.AADGpu_SetEncodeSegmentBlocks(gpu_ctx, 0)
        dec = gpu.decode_batch(gpu_ctx, aad, n_max, 44100, channels, bits, block, ms, sizes=sizes)
        for i in range(n_streams):
            x = pcm[i, :, :lens[i]]
            want = _segmented_reference(oracle, x, 44100, bits, block, ms, trials, seg_blocks)
            got = aad[i, :sizes[i]].tobytes()
            assert got == want, (block, seg_blocks, trials, i, int(lens[i]))
            rc, want_pcm, _ = oracle.decode(got)                     # the stock decoder takes it
            assert rc == 0 and np.array_equal(dec[i, :, :lens[i]], want_pcm), (block, seg_blocks, i)
        # one segment spanning the stream == the reference encoder; and so is the setting at 0
        if seg_blocks == 1000:
            plain, psizes = gpu.encode_batch(gpu_ctx, pcm, 44100, bits, block, ms, trials, num_samples=lens)
            assert np.array_equal(psizes, sizes)
            for i in range(n_streams):
                assert np.array_equal(plain[i, :sizes[i]], aad[i, :sizes[i]]), i


@pytest.mark.parametrize("generic", [0, 1])
def test_segments_through_sliced_pipeline_and_generic_kernel(product, gpu_ctx, oracle, generic):
    """AADGpu_ReconstructBatch cuts the batch into block-range slices whose edges fall inside segments: the
    chain state has to cross them per (stream, segment, channel)."""
    _, gpu = product
    n_streams, n_max, channels, bits, seg_blocks = 600, 30000, 2, 4, 4
    rng = np.random.default_rng(5)
    lens = rng.integers(1, n_max + 1, size=n_streams).astype(np.uint32)
    lens[0] = n_max
    pcm = np.zeros((n_streams, channels, n_max), dtype=np.int16)
    for i in range(n_streams):
        pcm[i, :, :lens[i]] = aadtest.signal(aadtest.SIGNALS[i % len(aadtest.SIGNALS)], channels, int(lens[i]), i)
    b = gpu.batch(n_streams, n_max, make_param(channels, 32000, bits, 1024, False, 2))
    aad = np.zeros((n_streams, b.aad_stream_stride), dtype=np.uint8)
    sizes = np.zeros(n_streams, dtype=np.uint32)
    out = np.zeros_like(pcm)
    gpu.lib.AADGpu_SetKernelPath(generic)
    assert gpu.lib.AADGpu_SetEncodeSegmentBlocks(gpu_ctx, seg_blocks) == OK
    try:
        rc = gpu.lib.AADGpu_ReconstructBatch(gpu_ctx, C.byref(b), pcm.ctypes.data, lens.ctypes.data, aad.ctypes.data,
                                             sizes.ctypes.data, out.ctypes.data)
        assert rc == OK, gpu.last_error()
    finally:
        gpu.lib.AADGpu_SetEncodeSegmentBlocks(gpu_ctx, 0)
        gpu.lib.AADGpu_SetKernelPath(0)
    for i in range(0, n_streams, 7):
        x = pcm[i, :, :lens[i]]
        want = _segmented_reference(oracle, x, 32000, bits, 1024, False, 2, seg_blocks)
        assert aad[i, :sizes[i]].tobytes() == want, i
        rc, want_pcm, _ = oracle.decode(want)
        assert rc == 0 and np.array_equal(out[i, :, :lens[i]], want_pcm), i


def test_cli_segment_blocks_output_decodes_with_the_reference_cli(product, gpu_ctx, tmp_path):
    """`aad -e --segment-blocks N` writes a stream the UNMODIFIED reference command line decodes, to the same
    samples as `aad -d`; the reconstruction stays as close to the input as the reference encoder's."""
    ref_cli = aadtest.ROOT / "oracle" / "_ref" / "aad_ref_cli"
    cli = aadtest.ROOT / "aad_b200" / "aad"
    if not ref_cli.exists():
        pytest.skip("oracle/_ref/aad_ref_cli did not travel")
    src = aadtest.GOLDEN / "pi_15-25sec.wav" if (aadtest.GOLDEN / "pi_15-25sec.wav").exists() else aadtest.GOLDEN / "sin300Hz.wav"
    seg, plain = tmp_path / "seg.aad", tmp_path / "plain.aad"
    subprocess.run([str(cli), "-e", "--segment-blocks", "5", str(src), str(seg)], check=True, capture_output=True)
    subprocess.run([str(cli), "-e", str(src), str(plain)], check=True, capture_output=True)
    assert seg.stat().st_size == plain.stat().st_size and seg.read_bytes() != plain.read_bytes()
    subprocess.run([str(ref_cli), "-d", str(seg), str(tmp_path / "seg_ref.wav")], check=True, capture_output=True)
    subprocess.run([str(cli), "-d", str(seg), str(tmp_path / "seg_b200.wav")], check=True, capture_output=True)
    assert (tmp_path / "seg_ref.wav").read_bytes() == (tmp_path / "seg_b200.wav").read_bytes()
    subprocess.run([str(ref_cli), "-d", str(plain), str(tmp_path / "plain_ref.wav")], check=True, capture_output=True)
    x, _ = aadtest.read_wav16(src)
    a, _ = aadtest.read_wav16(tmp_path / "seg_ref.wav")
    p, _ = aadtest.read_wav16(tmp_path / "plain_ref.wav")
    rms = lambda e: float(np.sqrt(np.mean(e.astype(np.float64) ** 2)))
    assert rms(a - x) < 1.5 * rms(p - x) + 1.0


def _device_sets(gpu):
    sets = [[0], [0, 0], [0, 0, 0]]
    if gpu.device_count() >= 2:
        sets.append(list(range(min(gpu.device_count(), 8))))
    return sets


@pytest.mark.parametrize("bits,channels,ms,block,seg_blocks", [(4, 2, False, 1024, 4), (3, 2, True, 1024, 1), (4, 1, False, 1024, 7),
                                                               (2, 1, False, 333, 3), (3, 8, False, 1024, 5)])
def test_one_stream_encodes_in_segment_range_shards(product, gpu_ctx, oracle, bits, channels, ms, block, seg_blocks):
    """AADGpuGroup_EncodeInterleaved16: the segments of ONE stream shared out over the devices of a group (a device may
    be named twice, so one GPU covers it) == the single-device segment encode == the oracle's fresh-handle encodes.
    Also a stream with fewer segments than devices."""
    _, gpu = product
    prm = make_param(channels, 48000, bits, block, ms, 2)
    for n in (150001, 700):
        pcm = aadtest.signal("music", channels, n, bits)
        inter = np.ascontiguousarray(pcm.T)
        want = _segmented_reference(oracle, pcm, 48000, bits, block, ms, 2, seg_blocks)
        cap = len(want) + 64
        for devices in _device_sets(gpu):
            arr = (C.c_int * len(devices))(*devices)
            g = gpu.lib.AADGpuGroup_Create(arr, len(devices))
            assert g, gpu.last_error()
            try:
                data = np.full(cap, 0xAB, dtype=np.uint8)
                size = C.c_uint32(0)
                rc = gpu.lib.AADGpuGroup_EncodeInterleaved16(g, C.byref(prm), seg_blocks, inter.ctypes.data, n, data.ctypes.data, cap, C.byref(size))
                assert rc == OK, gpu.last_error()
                assert size.value == len(want) and data[:size.value].tobytes() == want, (n, devices)
                assert (data[size.value:] == 0xAB).all()
                # refused without segments: one stream is a serial chain
                assert gpu.lib.AADGpuGroup_EncodeInterleaved16(g, C.byref(prm), 0, inter.ctypes.data, n, data.ctypes.data, cap, C.byref(size)) == 1
                assert "segment_blocks" in gpu.last_error()
                assert gpu.lib.AADGpuGroup_EncodeInterleaved16(g, C.byref(prm), seg_blocks, inter.ctypes.data, n, data.ctypes.data, 40, C.byref(size)) == 3
            finally:
                gpu.lib.AADGpuGroup_Destroy(g)
    # the single-device call with the context setting gives the same bytes
    assert gpu.lib.AADGpu_SetEncodeSegmentBlocks(gpu_ctx, seg_blocks) == OK
    try:
        data = np.zeros(cap, dtype=np.uint8)
        size = C.c_uint32(0)
        assert gpu.lib.AADGpu_EncodeInterleaved16(gpu_ctx, C.byref(prm), inter.ctypes.data, n, data.ctypes.data, cap, C.byref(size)) == OK
        assert data[:size.value].tobytes() == want
    finally:
        gpu.lib.AADGpu_SetEncodeSegmentBlocks(gpu_ctx, 0)


def test_cli_segment_encode_on_a_device_list(tmp_path):
    cli = aadtest.ROOT / "aad_b200" / "aad"
    src = aadtest.GOLDEN / "pi_15-25sec.wav"
    one, many = tmp_path / "one.aad", tmp_path / "many.aad"
    for out, dev in ((one, "0"), (many, "0,0,0")):
        r = subprocess.run([str(cli), "-e", "-S", "6", "--device", dev, str(src), str(out)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
    assert one.read_bytes() == many.read_bytes()
