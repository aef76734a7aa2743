"""Host-side API behaviour, differentially against the UNMODIFIED compiled reference (oracle/_ref):
every entry point that answers without touching the device -- block-size arithmetic, header
encode / decode / validation, work-size and handle creation rules, parameter validation, call-order
errors -- must return what the reference returns, field for field, over seeded random and boundary
inputs.  CPU only; skipped where the reference sources (and hence oracle/_ref) are absent.

(The stock reference handles 1-2 channels, so the product is switched to its stock limit here;
NULL-pointer rows are covered by tests/test_host_api.py.)
"""
import ctypes as C

import numpy as np
import pytest

from aad_b200.capi import HeaderInfo, EncodeParameter, OK, HEADER_SIZE

U8P = C.POINTER(C.c_uint8)


@pytest.fixture()
def pair(product, ref):
    api, gpu = product
    gpu.lib.AADGpu_SetMaxChannels(2)
    yield api, ref
    gpu.lib.AADGpu_SetMaxChannels(8)


def interesting_u16(rng):
    return int(rng.choice([0, 1, 2, 17, 18, 19, 35, 36, 37, 40, 63, 64, 128, 255, 256, 1023, 1024, 4096, 32767, 32768, 65535,
                           int(rng.integers(0, 65536))]))


def test_calculate_block_size_everywhere(pair):
    a, r = pair
    rng = np.random.default_rng(1)
    cases = [(mb, ch, bits) for mb in (0, 17, 18, 19, 20, 35, 36, 37, 38, 39, 40, 41, 64, 100, 1023, 1024, 1025, 65535)
             for ch in (0, 1, 2, 3) for bits in (0, 1, 2, 3, 4, 5, 8)]
    cases += [(interesting_u16(rng), int(rng.integers(0, 4)), int(rng.integers(0, 7))) for _ in range(3000)]
    for mb, ch, bits in cases:
        assert a.calculate_block_size(mb, ch, bits) == r.calculate_block_size(mb, ch, bits), (mb, ch, bits)
        # num_samples_per_block may be NULL (src/aad_encoder.c:90-93)
        ba, br = C.c_uint16(7), C.c_uint16(7)
        ra = a.lib.AADEncoder_CalculateBlockSize(mb, ch, bits, C.byref(ba), None)
        rr = r.lib.AADEncoder_CalculateBlockSize(mb, ch, bits, C.byref(br), None)
        assert (ra, ba.value) == (rr, br.value), (mb, ch, bits)


def random_header(rng):
    h = HeaderInfo()
    h.format_version = int(rng.choice([4, 4, 4, 0, 3, 5]))
    h.codec_version = int(rng.choice([18, 18, 18, 0, 17, 19]))
    h.num_channels = int(rng.choice([1, 2, 1, 2, 0, 3, 9]))
    h.num_samples = int(rng.choice([0, 1, 3, 4, 5, 1000, 24000, 2 ** 32 - 1]))
    h.sampling_rate = int(rng.choice([0, 1, 8000, 44100, 48000, 2 ** 32 - 1]))
    h.bits_per_sample = int(rng.choice([4, 3, 2, 4, 0, 1, 5, 16]))
    h.block_size = interesting_u16(rng)
    h.num_samples_per_block = int(rng.choice([0, 1, 4, 5, 292, 992, 2016, 4028, 2 ** 31]))
    h.ch_process_method = int(rng.choice([0, 0, 1, 1, 2, 3]))
    return h


def test_encode_and_decode_header_everywhere(pair):
    a, r = pair
    rng = np.random.default_rng(2)
    for k in range(4000):
        h = random_header(rng)
        size = int(rng.choice([HEADER_SIZE, HEADER_SIZE, 64, 30, 0, 12]))
        bufs = []
        for lib in (a.lib, r.lib):
            buf = np.full(64, 0xA5, dtype=np.uint8)
            rc = lib.AADEncoder_EncodeHeader(C.byref(h), buf.ctypes.data_as(U8P), size)
            bufs.append((rc, buf.tobytes()))
        assert bufs[0] == bufs[1], (k, h.as_dict(), size)      # same result AND same bytes (nothing written on error)
        if bufs[0][0] == OK:
            raw = np.frombuffer(bufs[0][1], dtype=np.uint8).copy()
            # decode what was written, and mutations of it (bad signature / versions / fields / sizes)
            for mutate in range(6):
                m = raw.copy()
                if mutate == 1:
                    m[int(rng.integers(0, 4))] ^= 0x20
                elif mutate >= 2:
                    m[int(rng.integers(4, HEADER_SIZE))] = int(rng.integers(0, 256))
                dsize = int(rng.choice([HEADER_SIZE, 64, 30, 4]))
                outs = []
                for lib in (a.lib, r.lib):
                    out = HeaderInfo()
                    rc = lib.AADDecoder_DecodeHeader(m.ctypes.data_as(U8P), dsize, C.byref(out))
                    outs.append((rc, out.as_dict() if rc == OK else None))
                assert outs[0] == outs[1], (k, mutate, dsize)
                if outs[0][0] == OK:       # SetHeader = CheckHeaderFormat (src/aad_decoder.c:173-225)
                    hd = HeaderInfo(**outs[0][1])
                    res = []
                    for lib in (a.lib, r.lib):
                        dec = lib.AADDecoder_Create(None, 0)
                        res.append(lib.AADDecoder_SetHeader(dec, C.byref(hd)))
                        lib.AADDecoder_Destroy(dec)
                    assert res[0] == res[1], (k, mutate, hd.as_dict())


def test_set_header_on_raw_structs(pair):
    a, r = pair
    rng = np.random.default_rng(3)
    for k in range(3000):
        h = random_header(rng)
        res = []
        for lib in (a.lib, r.lib):
            dec = lib.AADDecoder_Create(None, 0)
            res.append(lib.AADDecoder_SetHeader(dec, C.byref(h)))
            lib.AADDecoder_Destroy(dec)
        assert res[0] == res[1], (k, h.as_dict())


def test_set_encode_parameter_everywhere(pair):
    a, r = pair
    rng = np.random.default_rng(4)
    for k in range(4000):
        prm = EncodeParameter(int(rng.choice([1, 2, 0, 3])), int(rng.choice([0, 8000, 48000])),
                              int(rng.choice([2, 3, 4, 0, 1, 5, 9])), interesting_u16(rng),
                              int(rng.choice([0, 1, 2, 5])), int(rng.integers(0, 4)))
        create_block = int(rng.choice([prm.max_block_size, 1024, 64, 40]))
        res = []
        for lib in (a.lib, r.lib):
            enc = lib.AADEncoder_Create(create_block, None, 0)
            if not enc:
                res.append("no handle")
                continue
            res.append(lib.AADEncoder_SetEncodeParameter(enc, C.byref(prm)))
            lib.AADEncoder_Destroy(enc)
        assert res[0] == res[1], (k, create_block, [getattr(prm, f) for f, _ in prm._fields_])


def test_work_size_and_create_rules(pair):
    a, r = pair
    for mb in (0, 1, 17, 18, 19, 36, 64, 1024, 65535):
        wa, wr = a.lib.AADEncoder_CalculateWorkSize(mb), r.lib.AADEncoder_CalculateWorkSize(mb)
        assert (wa < 0) == (wr < 0), mb          # the sizes themselves differ (no host sample buffers here)
        for lib, need in ((a.lib, wa), (r.lib, wr)):
            if need < 0:
                assert not lib.AADEncoder_Create(mb, None, 0)
                continue
            buf = (C.c_uint8 * (need + 64))()
            assert not lib.AADEncoder_Create(mb, None, need)            # NULL with a size
            assert not lib.AADEncoder_Create(mb, buf, 0)                # memory with no size
            assert not lib.AADEncoder_Create(mb, buf, need - 1)         # too small
            h = lib.AADEncoder_Create(mb, buf, need)
            assert h and C.addressof(buf) <= h < C.addressof(buf) + need and h % 16 == 0
            lib.AADEncoder_Destroy(h)                                   # must not free caller memory
            h2 = lib.AADEncoder_Create(mb, None, 0)
            assert h2
            lib.AADEncoder_Destroy(h2)
    for lib in (a.lib, r.lib):
        need = lib.AADDecoder_CalculateWorkSize()
        assert need > 0
        buf = (C.c_uint8 * (need + 64))()
        assert not lib.AADDecoder_Create(None, need) and not lib.AADDecoder_Create(buf, 0)
        assert not lib.AADDecoder_Create(buf, need - 1)
        h = lib.AADDecoder_Create(buf, need)
        assert h and h % 16 == 0
        lib.AADDecoder_Destroy(h)


def test_call_order_errors(pair):
    """PARAMETER_NOT_SET before SetEncodeParameter / SetHeader; argument and buffer checks that answer
    before any sample is touched (src/aad_encoder.c:824-846, src/aad_decoder.c:331-361, :487-509)."""
    a, r = pair
    i32pp = C.POINTER(C.POINTER(C.c_int32))
    pcm = np.zeros(64, dtype=np.int32)
    rows = (C.POINTER(C.c_int32) * 2)(pcm.ctypes.data_as(C.POINTER(C.c_int32)), pcm.ctypes.data_as(C.POINTER(C.c_int32)))
    out = np.zeros(4096, dtype=np.uint8)
    size = C.c_uint32(0)
    nd = C.c_uint32(0)
    hdr_ok = HeaderInfo(4, 18, 2, 64, 48000, 4, 1024, 992, 0)
    for lib in (a.lib, r.lib):
        enc = lib.AADEncoder_Create(1024, None, 0)
        got = [lib.AADEncoder_EncodeWhole(enc, rows, 64, out.ctypes.data_as(U8P), 4096, C.byref(size))]
        lib.AADEncoder_SetEncodeParameter(enc, C.byref(EncodeParameter(2, 48000, 4, 1024, 0, 0)))
        got.append(lib.AADEncoder_EncodeWhole(enc, rows, 64, out.ctypes.data_as(U8P), 30, C.byref(size)))   # no room for the header
        got.append(lib.AADEncoder_EncodeWhole(enc, rows, 0, out.ctypes.data_as(U8P), 4096, C.byref(size)))   # no samples
        lib.AADEncoder_Destroy(enc)
        dec = lib.AADDecoder_Create(None, 0)
        got.append(lib.AADDecoder_DecodeBlock(dec, out.ctypes.data_as(U8P), 1024, rows, 2, 64, C.byref(nd)))
        lib.AADDecoder_SetHeader(dec, C.byref(hdr_ok))
        got.append(lib.AADDecoder_DecodeBlock(dec, out.ctypes.data_as(U8P), 35, rows, 2, 64, C.byref(nd)))   # shorter than the block header
        got.append(lib.AADDecoder_DecodeBlock(dec, out.ctypes.data_as(U8P), 1024, rows, 1, 64, C.byref(nd)))  # too few channel rows
        hb = np.zeros(64, dtype=np.uint8)
        lib.AADEncoder_EncodeHeader(C.byref(hdr_ok), hb.ctypes.data_as(U8P), 64)
        got.append(lib.AADDecoder_DecodeWhole(dec, hb.ctypes.data_as(U8P), 30, rows, 2, 64))                 # header cut short
        got.append(lib.AADDecoder_DecodeWhole(dec, hb.ctypes.data_as(U8P), 64, rows, 1, 64))                 # too few rows
        got.append(lib.AADDecoder_DecodeWhole(dec, hb.ctypes.data_as(U8P), 64, rows, 2, 63))                 # too few samples
        lib.AADDecoder_Destroy(dec)
        if lib is a.lib:
            first = got
    assert first == got, (first, got)
