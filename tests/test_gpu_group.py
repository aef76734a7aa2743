"""Several devices of one box: AADGpuGroup_* shard batches by stream and the
decode of one long stream by block range, one host thread per device, no collective.

A group may name the same device more than once (two contexts, two threads, two shards), so the
sharding and reassembly are covered on a one-GPU box too; with >= 2 GPUs the same tests also run
over distinct devices.  Bar: identical bytes / samples to the single-device calls and the oracle.
"""
import ctypes as C
import subprocess

import numpy as np
import pytest

import aadtest
import aad_b200
from aad_b200.capi import make_param

pytestmark = pytest.mark.gpu


def device_sets(gpu):
    sets = [[0], [0, 0], [0, 0, 0]]
    n = gpu.device_count()
    if n >= 2:
        sets.append(list(range(min(n, 8))))
    return sets


def make_group(gpu, devices):
    arr = (C.c_int * len(devices))(*devices)
    g = gpu.lib.AADGpuGroup_Create(arr, len(devices))
    assert g and gpu.lib.AADGpuGroup_Size(g) == len(devices), gpu.last_error()
    return g


@pytest.mark.parametrize("bits,channels,ms", [(4, 1, False), (3, 2, True), (2, 8, False)])
def test_group_batches_equal_single_device(product, gpu_ctx, bits, channels, ms):
    _, gpu = product
    n_streams, n_max = 23, 6100
    rng = np.random.default_rng(bits + channels)
    lens = rng.integers(1, n_max + 1, size=n_streams).astype(np.uint32)
    lens[0] = n_max
    pcm = np.zeros((n_streams, channels, n_max), dtype=np.int16)
    for i in range(n_streams):
        pcm[i, :, :lens[i]] = aadtest.signal(aadtest.SIGNALS[i % len(aadtest.SIGNALS)], channels, int(lens[i]), i)
    want_aad, want_sizes = gpu.encode_batch(gpu_ctx, pcm, 44100, bits, 1024, ms, 2, num_samples=lens)
    want_pcm = gpu.decode_batch(gpu_ctx, want_aad, n_max, 44100, channels, bits, 1024, ms, sizes=want_sizes)
    prm = make_param(channels, 44100, bits, 1024, ms, 2)
    b = gpu.batch(n_streams, n_max, prm)
    for devices in device_sets(gpu):
        g = make_group(gpu, devices)
        try:
            aad = np.zeros_like(want_aad)
            sizes = np.zeros(n_streams, dtype=np.uint32)
            rc = gpu.lib.AADGpuGroup_EncodeBatch(g, C.byref(b), pcm.ctypes.data, lens.ctypes.data, aad.ctypes.data, sizes.ctypes.data)
            assert rc == 0, gpu.last_error()
            assert np.array_equal(sizes, want_sizes) and np.array_equal(aad, want_aad), devices
            out = np.zeros_like(want_pcm)
            rc = gpu.lib.AADGpuGroup_DecodeBatch(g, C.byref(b), aad.ctypes.data, sizes.ctypes.data, out.ctypes.data)
            assert rc == 0, gpu.last_error()
            for i in range(n_streams):
                assert np.array_equal(out[i, :, :lens[i]], want_pcm[i, :, :lens[i]]), (devices, i)
        finally:
            gpu.lib.AADGpuGroup_Destroy(g)


@pytest.mark.parametrize("bits,channels,ms,block", [(4, 2, False, 1024), (3, 2, True, 1024), (3, 1, False, 1024),
                                                    (2, 1, False, 333), (3, 8, False, 1024), (4, 4, False, 1024),
                                                    (2, 5, True, 1024)])
def test_one_stream_decodes_in_block_range_shards(product, oracle, bits, channels, ms, block):
    """Every block header reloads the chain state (src/aad_decoder.c:364-380): shards by block range,
    interleaved output, against the oracle's whole-stream decode.  Also a stream with fewer blocks
    than shards and one cut short in the middle of a block."""
    _, gpu = product
    for n in (150001, 700):
        pcm = aadtest.signal("music", channels, n, bits)
        rc, data = oracle.encode(pcm, 48000, bits, block, ms, 1)
        assert rc == 0
        for cut in (len(data), len(data) - (len(data) - 31) // 3):
            blob = np.frombuffer(data[:cut], dtype=np.uint8).copy()
            want_rc, want, _ = oracle.decode(data[:cut], fill=0)
            assert want_rc in (0, 4)                 # 4 = INSUFFICIENT_DATA when the last block lost its header
            for devices in device_sets(gpu):
                g = make_group(gpu, devices)
                try:
                    out = np.full((n, channels), 12345, dtype=np.int16)
                    rc = gpu.lib.AADGpuGroup_DecodeInterleaved16(g, blob.ctypes.data, len(blob), out.ctypes.data, n)
                    assert rc == want_rc, gpu.last_error()   # the earlier blocks are decoded either way (src/aad_decoder.c:522-527)
                    assert np.array_equal(out.T, want), (n, cut, devices)
                finally:
                    gpu.lib.AADGpuGroup_Destroy(g)


def test_group_argument_errors(product):
    _, gpu = product
    assert not gpu.lib.AADGpuGroup_Create((C.c_int * 1)(99), 1)
    assert "no such CUDA device" in gpu.last_error()
    g = make_group(gpu, [0, 0])
    try:
        junk = np.zeros(64, dtype=np.uint8)
        out = np.zeros(64, dtype=np.int16)
        assert gpu.lib.AADGpuGroup_DecodeInterleaved16(g, junk.ctypes.data, 20, out.ctypes.data, 32) == 4      # INSUFFICIENT_DATA
        assert gpu.lib.AADGpuGroup_DecodeInterleaved16(g, junk.ctypes.data, 64, out.ctypes.data, 32) == 2      # INVALID_FORMAT
        assert gpu.lib.AADGpuGroup_DecodeInterleaved16(g, None, 64, out.ctypes.data, 32) == 1                  # INVALID_ARGUMENT
        data = (aadtest.GOLDEN / "sin300Hz.aad").read_bytes()
        blob = np.frombuffer(data, dtype=np.uint8).copy()
        small = np.zeros(100, dtype=np.int16)
        assert gpu.lib.AADGpuGroup_DecodeInterleaved16(g, blob.ctypes.data, len(blob), small.ctypes.data, 50) == 3   # INSUFFICIENT_BUFFER
    finally:
        gpu.lib.AADGpuGroup_Destroy(g)


def test_cli_on_a_device_list(tmp_path):
    cli = aad_b200.PACKAGE_DIR / "aad"
    for stem in ("sin300Hz", "sin300Hz_mono"):
        r = subprocess.run([str(cli), "-d", "--device", "0,0,0", str(aadtest.GOLDEN / f"{stem}.aad"), str(tmp_path / "o.wav")],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        assert (tmp_path / "o.wav").read_bytes() == (aadtest.GOLDEN / f"{stem}_decoded.wav").read_bytes()
    r = subprocess.run([str(cli), "-d", "-D", "all", str(aadtest.GOLDEN / "sin300Hz.aad"), str(tmp_path / "a.wav")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and (tmp_path / "a.wav").read_bytes() == (aadtest.GOLDEN / "sin300Hz_decoded.wav").read_bytes()
