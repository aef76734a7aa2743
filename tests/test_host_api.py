"""Host-side logic of libaad_b200.so on a machine WITHOUT a GPU: the library loads, exports
every symbol include/*.h declares, and its O(1) host functions (block geometry, stream header,
handles, parameter validation) behave exactly like the reference's -- the same assertions as
test/test_aad_encoder.c:23-333 and test/test_aad_decoder.c:33-253, run side by side against the
compiled reference when it is available.  No compute entry point is called here."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from aad_b200 import capi
from aad_b200.capi import (INSUFFICIENT_DATA, INVALID_ARGUMENT, INVALID_FORMAT, OK, HeaderInfo, block_header_size,
                           make_param)
from test_oracle import BLOCK_SIZE_KAT

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol(product):
    api, gpu = product
    declared = set()
    for header in (ROOT / "include").glob("*.h"):
        text = re.sub(r"/\*.*?\*/", "", header.read_text(), flags=re.S)
        declared |= set(re.findall(r"\b(AAD(?:Encoder|Decoder|Gpu|GpuGroup)_\w+)\s*\(", text))
    assert len(declared) >= 14 + 15
    for name in sorted(declared):
        assert hasattr(api.lib, name), f"{name} declared in include/ but not exported"
    assert set(api.SYMBOLS) <= declared
    assert set(gpu.SYMBOLS) <= declared


def test_ctypes_bindings_match_the_header_prototypes(product):
    """Every prototype of include/aad_b200.h that aad_b200/gpu.py binds takes as many arguments as the binding
    passes (a drifted binding would corrupt the call instead of failing)."""
    _, gpu = product
    text = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "aad_b200.h").read_text(), flags=re.S)
    arity = {}
    for name, params in re.findall(r"\b(AADGpu(?:Group)?_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        params = params.strip()
        arity[name] = 0 if params in ("", "void") else params.count(",") + 1
    checked = 0
    for name in gpu.SYMBOLS:
        fn = getattr(gpu.lib, name)
        if fn.argtypes is None or name not in arity:
            continue
        assert len(fn.argtypes) == arity[name], (name, len(fn.argtypes), arity[name])
        checked += 1
    assert checked >= 25


@pytest.mark.parametrize("max_block,ch,bits,bs,spb", BLOCK_SIZE_KAT)
def test_calculate_block_size_known_answers(product, max_block, ch, bits, bs, spb):
    assert product[0].calculate_block_size(max_block, ch, bits) == (OK, bs, spb)


def test_calculate_block_size_errors(product):
    api, gpu = product
    lib = api.lib
    bs, spb = C.c_uint16(), C.c_uint32()
    assert lib.AADEncoder_CalculateBlockSize(32, 1, 4, None, None) == INVALID_ARGUMENT
    assert lib.AADEncoder_CalculateBlockSize(32, 1, 4, None, C.byref(spb)) == INVALID_ARGUMENT
    assert lib.AADEncoder_CalculateBlockSize(32, 1, 4, C.byref(bs), None) == OK and bs.value == 32
    assert api.calculate_block_size(block_header_size(1) - 1, 1, 4)[0] == INVALID_FORMAT
    assert api.calculate_block_size(32, 0, 4)[0] == INVALID_FORMAT
    assert api.calculate_block_size(1024, gpu.lib.AADGpu_GetMaxChannels() + 1, 4)[0] == INVALID_FORMAT
    assert api.calculate_block_size(32, 1, 0)[0] == INVALID_FORMAT
    assert api.calculate_block_size(32, 1, capi.MAX_BITS + 1)[0] == INVALID_FORMAT     # 5 bits: not a format


def test_stock_channel_limit_is_selectable(product):
    api, gpu = product
    try:
        gpu.lib.AADGpu_SetMaxChannels(2)
        assert api.calculate_block_size(1024, 3, 4)[0] == INVALID_FORMAT     # src/aad.h:13
        assert api.calculate_block_size(1024, 2, 4)[0] == OK
    finally:
        gpu.lib.AADGpu_SetMaxChannels(8)
    assert api.calculate_block_size(1024, 8, 3) == (OK, 1008, 292)


def _valid_header(**kw):
    h = HeaderInfo(0, 0, 1, 1024, 44100, capi.MAX_BITS, 32, 32, capi.CH_NONE)
    for k, v in kw.items():
        setattr(h, k, v)
    return h


def _encode_header(api, h, size=capi.HEADER_SIZE):
    data = np.zeros(64, dtype=np.uint8)
    rc = api.lib.AADEncoder_EncodeHeader(C.byref(h), data.ctypes.data_as(C.POINTER(C.c_uint8)), size)
    return rc, data


HEADER_REJECTS = [
    dict(num_channels=0), dict(num_channels=9), dict(num_samples=0), dict(sampling_rate=0), dict(bits_per_sample=0),
    dict(bits_per_sample=1), dict(bits_per_sample=5), dict(block_size=0), dict(block_size=17), dict(block_size=18),
    dict(num_samples_per_block=0), dict(ch_process_method=capi.CH_INVALID),
    dict(num_channels=1, ch_process_method=capi.CH_MS),
]


def test_encode_header_layout_and_errors(product, request):
    api, _ = product
    libs = [api]
    if (ROOT / "oracle" / "_ref" / "libaad_ref.so").exists():
        libs.append(request.getfixturevalue("ref"))
    h = _valid_header(num_channels=2, num_samples=0x01020304, sampling_rate=48000, bits_per_sample=3, block_size=1020,
                      num_samples_per_block=1316, ch_process_method=capi.CH_MS, format_version=99, codec_version=77)
    images = []
    for lib in libs:
        rc, data = _encode_header(lib, h)
        assert rc == OK
        images.append(bytes(data[:31]))
    d = images[0]
    assert all(img == d for img in images)
    # offsets checked by test/test_aad_decoder.c:95-184; versions come from the macros, not the struct
    assert d[:4] == b"AAD\0" and d[4:8] == bytes([0, 0, 0, 4]) and d[8:12] == bytes([0, 0, 0, 18])
    assert d[12:14] == bytes([0, 2]) and d[14:18] == bytes([1, 2, 3, 4]) and d[18:22] == (48000).to_bytes(4, "big")
    assert d[22:24] == bytes([0, 3]) and d[24:26] == (1020).to_bytes(2, "big") and d[26:30] == (1316).to_bytes(4, "big")
    assert d[30] == 1
    for lib in libs:
        hh = _valid_header()
        data = np.zeros(64, dtype=np.uint8)
        p8 = data.ctypes.data_as(C.POINTER(C.c_uint8))
        assert lib.lib.AADEncoder_EncodeHeader(None, p8, 31) == INVALID_ARGUMENT
        assert lib.lib.AADEncoder_EncodeHeader(C.byref(hh), None, 31) == INVALID_ARGUMENT
        assert _encode_header(lib, hh, 30)[0] == INSUFFICIENT_DATA
        for bad in HEADER_REJECTS:
            if lib is not api and bad.get("num_channels") == 9:
                bad = dict(num_channels=3)
            rc, data = _encode_header(lib, _valid_header(**bad))
            assert rc == INVALID_FORMAT, bad
            assert not data.any(), "nothing may be written before validation passes"


def test_decode_header_roundtrip_and_errors(product):
    api, _ = product
    h = _valid_header(num_channels=2, bits_per_sample=2, block_size=128, num_samples_per_block=188)
    rc, data = _encode_header(api, h)
    assert rc == OK
    rc, got = api.decode_header(bytes(data[:31]))
    assert rc == OK
    want = h.as_dict()
    want.update(format_version=capi.FORMAT_VERSION, codec_version=capi.CODEC_VERSION)
    assert got.as_dict() == want
    out = HeaderInfo()
    p8 = data.ctypes.data_as(C.POINTER(C.c_uint8))
    assert api.lib.AADDecoder_DecodeHeader(None, 31, C.byref(out)) == INVALID_ARGUMENT
    assert api.lib.AADDecoder_DecodeHeader(p8, 31, None) == INVALID_ARGUMENT
    assert api.lib.AADDecoder_DecodeHeader(p8, 30, C.byref(out)) == INSUFFICIENT_DATA
    bad = bytearray(data[:31])
    bad[0] = ord("a")
    assert api.decode_header(bytes(bad))[0] == INVALID_FORMAT
    # a parseable header with bad fields is accepted by DecodeHeader and rejected by SetHeader
    dec = api.lib.AADDecoder_Create(None, 0)
    assert dec
    for off, width, value in [(4, 4, 0), (4, 4, 5), (8, 4, 0), (8, 4, 19), (12, 2, 0), (12, 2, 9), (14, 4, 0),
                              (18, 4, 0), (22, 2, 0), (22, 2, 5), (24, 2, 0), (24, 2, 35), (26, 4, 0), (30, 1, 2)]:
        img = bytearray(data[:31])
        img[off:off + width] = value.to_bytes(width, "big")
        rc, parsed = api.decode_header(bytes(img))
        assert rc == OK
        assert api.lib.AADDecoder_SetHeader(dec, C.byref(parsed)) == INVALID_FORMAT, (off, value)
    img = bytearray(data[:31])
    img[12:14] = (1).to_bytes(2, "big")
    img[30] = capi.CH_MS
    rc, parsed = api.decode_header(bytes(img))
    assert rc == OK and api.lib.AADDecoder_SetHeader(dec, C.byref(parsed)) == INVALID_FORMAT
    rc, parsed = api.decode_header(bytes(data[:31]))
    assert api.lib.AADDecoder_SetHeader(dec, C.byref(parsed)) == OK
    assert api.lib.AADDecoder_SetHeader(None, C.byref(parsed)) == INVALID_ARGUMENT
    assert api.lib.AADDecoder_SetHeader(dec, None) == INVALID_ARGUMENT
    api.lib.AADDecoder_Destroy(dec)


def test_encoder_handle_lifecycle(product):
    lib = product[0].lib
    ws = lib.AADEncoder_CalculateWorkSize(1024)
    assert ws > 0
    assert lib.AADEncoder_CalculateWorkSize(0) == -1
    work = C.create_string_buffer(ws)
    enc = lib.AADEncoder_Create(1024, work, ws)
    assert enc and C.addressof(work) <= enc < C.addressof(work) + 16
    lib.AADEncoder_Destroy(enc)                       # caller memory: frees nothing
    own = lib.AADEncoder_Create(1024, None, 0)
    assert own
    lib.AADEncoder_Destroy(own)
    lib.AADEncoder_Destroy(None)
    assert not lib.AADEncoder_Create(0, None, 0)
    assert not lib.AADEncoder_Create(0, work, ws)
    assert not lib.AADEncoder_Create(1024, None, ws)
    assert not lib.AADEncoder_Create(1024, work, 0)
    assert not lib.AADEncoder_Create(1024, work, ws - 1)


def test_decoder_handle_lifecycle(product):
    lib = product[0].lib
    ws = lib.AADDecoder_CalculateWorkSize()
    assert ws > 0
    work = C.create_string_buffer(ws)
    dec = lib.AADDecoder_Create(work, ws)
    assert dec
    lib.AADDecoder_Destroy(dec)
    own = lib.AADDecoder_Create(None, 0)
    assert own
    lib.AADDecoder_Destroy(own)
    assert not lib.AADDecoder_Create(None, ws)
    assert not lib.AADDecoder_Create(work, 0)
    assert not lib.AADDecoder_Create(work, ws - 1)


def test_set_encode_parameter_and_call_order_errors(product):
    api, _ = product
    lib = api.lib
    enc = lib.AADEncoder_Create(256, None, 0)
    ok = make_param(1, 8000, 4, 256, False, 1)
    assert lib.AADEncoder_SetEncodeParameter(None, C.byref(ok)) == INVALID_ARGUMENT
    assert lib.AADEncoder_SetEncodeParameter(enc, None) == INVALID_ARGUMENT
    for bad in (make_param(1, 8000, 0, 256), make_param(1, 8000, 5, 256), make_param(1, 8000, 4, 0),
                make_param(1, 8000, 4, 17), make_param(0, 8000, 4, 256), make_param(9, 8000, 4, 1024)):
        assert lib.AADEncoder_SetEncodeParameter(enc, C.byref(bad)) == INVALID_FORMAT
    bad = make_param(1, 8000, 4, 256)
    bad.ch_process_method = capi.CH_INVALID
    assert lib.AADEncoder_SetEncodeParameter(enc, C.byref(bad)) == INVALID_FORMAT
    # EncodeWhole before SetEncodeParameter, and NULL arguments (src/aad_encoder.c:826-834)
    pcm = np.zeros((1, 64), dtype=np.int32)
    out = np.zeros(4096, dtype=np.uint8)
    size = C.c_uint32()
    rows = capi._planar_pointers([pcm[0]])
    p8 = out.ctypes.data_as(C.POINTER(C.c_uint8))
    assert lib.AADEncoder_EncodeWhole(enc, rows, 64, p8, 4096, C.byref(size)) == capi.PARAMETER_NOT_SET
    assert lib.AADEncoder_SetEncodeParameter(enc, C.byref(ok)) == OK
    assert lib.AADEncoder_EncodeWhole(None, rows, 64, p8, 4096, C.byref(size)) == INVALID_ARGUMENT
    assert lib.AADEncoder_EncodeWhole(enc, None, 64, p8, 4096, C.byref(size)) == INVALID_ARGUMENT
    assert lib.AADEncoder_EncodeWhole(enc, rows, 64, None, 4096, C.byref(size)) == INVALID_ARGUMENT
    assert lib.AADEncoder_EncodeWhole(enc, rows, 64, p8, 4096, None) == INVALID_ARGUMENT
    # header problems surface before any device work: too small a buffer, zero samples
    assert lib.AADEncoder_EncodeWhole(enc, rows, 64, p8, 30, C.byref(size)) == INSUFFICIENT_DATA
    assert lib.AADEncoder_EncodeWhole(enc, rows, 0, p8, 4096, C.byref(size)) == INVALID_FORMAT
    lib.AADEncoder_Destroy(enc)
    # decoder: DecodeBlock before SetHeader, NULLs
    dec = lib.AADDecoder_Create(None, 0)
    n = C.c_uint32()
    assert lib.AADDecoder_DecodeBlock(dec, p8, 64, rows, 1, 64, C.byref(n)) == capi.PARAMETER_NOT_SET
    assert lib.AADDecoder_DecodeBlock(None, p8, 64, rows, 1, 64, C.byref(n)) == INVALID_ARGUMENT
    assert lib.AADDecoder_DecodeBlock(dec, None, 64, rows, 1, 64, C.byref(n)) == INVALID_ARGUMENT
    assert lib.AADDecoder_DecodeBlock(dec, p8, 64, None, 1, 64, C.byref(n)) == INVALID_ARGUMENT
    assert lib.AADDecoder_DecodeBlock(dec, p8, 64, rows, 1, 64, None) == INVALID_ARGUMENT
    assert lib.AADDecoder_DecodeWhole(None, p8, 64, rows, 1, 64) == INVALID_ARGUMENT
    assert lib.AADDecoder_DecodeWhole(dec, None, 64, rows, 1, 64) == INVALID_ARGUMENT
    assert lib.AADDecoder_DecodeWhole(dec, p8, 64, None, 1, 64) == INVALID_ARGUMENT
    assert lib.AADDecoder_DecodeWhole(dec, p8, 10, rows, 1, 64) == INSUFFICIENT_DATA
    assert lib.AADDecoder_DecodeWhole(dec, p8, 64, rows, 1, 64) == INVALID_FORMAT      # no "AAD\0" signature
    # a valid header with too small a PCM buffer (src/aad_decoder.c:506-509)
    rc, img = _encode_header(api, _valid_header(num_channels=2, block_size=64, num_samples_per_block=32))
    p_img = img.ctypes.data_as(C.POINTER(C.c_uint8))
    rows2 = capi._planar_pointers([np.zeros(2048, dtype=np.int32), np.zeros(2048, dtype=np.int32)])
    assert lib.AADDecoder_DecodeWhole(dec, p_img, 64, rows2, 1, 2048) == capi.INSUFFICIENT_BUFFER
    assert lib.AADDecoder_DecodeWhole(dec, p_img, 64, rows2, 2, 1023) == capi.INSUFFICIENT_BUFFER
    lib.AADDecoder_Destroy(dec)


def test_no_cpu_fallback_without_a_device(product):
    """On a box without CUDA the compute entry points must fail loudly, not compute on the CPU."""
    api, gpu = product
    if gpu.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(Exception):
        gpu.create(0)
    assert "no CPU fallback" in gpu.last_error() or gpu.last_error()
    pcm = np.zeros((1, 64), dtype=np.int32)
    rc, data = api.encode_whole(pcm, 8000, 4)
    assert rc == capi.NG and data == b""


def test_stream_size_helpers(product, oracle):
    _, gpu = product
    import aadtest
    for ch, bits, block, n in [(1, 4, 1024, 441000), (2, 3, 1024, 5000), (2, 2, 256, 4), (1, 3, 64, 125), (8, 3, 1024, 3000)]:
        prm = make_param(ch, 44100, bits, block)
        pcm = aadtest.signal("music", ch, n, 3)
        rc, data = oracle.encode(pcm, 44100, bits, block, False, 0)
        assert rc == 0
        assert gpu.stream_bytes(prm, n) == len(data)
        assert gpu.stream_bytes_bound(prm, n) >= len(data)


def test_synth_generator_mirror_is_deterministic(product):
    _, gpu = product
    from aad_b200.synth import synth_pcm16
    lut = gpu.synth_lut()
    assert lut[0] == 0 and lut[256] == 32767 and lut[768] == -32767
    a = synth_pcm16(lut, 5, 2, 2, 1000, 44100)
    b = synth_pcm16(lut, 6, 1, 2, 1000, 44100)
    assert np.array_equal(a[1], b[0]) and not np.array_equal(a[0], a[1])
    assert a.dtype == np.int16 and abs(int(a.max())) > 10000


def test_shift_free_quantiser_tables_are_exact():
    """The production quantiser (aad_encode_fast.cuh, enc_sample): min(umulhi(|d|, D), maxmag) for the regular rows and
    min(umulhi(|d| << (b-1), D), maxmag) for the first AADK_TINY_ROWSb rows equals min((|d| << (b-2)) / step, maxmag)
    (src/aad_encoder.c:372) for every table step and every reachable |d| (|sample - predict| < 2^17)."""
    import re
    text = (ROOT / "aad_b200" / "csrc" / "aad_tables_data.h").read_text().replace("\\\n", " ")

    def table(name):
        body = re.search(name + r" \{(.*?)\}", text, re.S).group(1)
        return [int(v.rstrip("u")) for v in re.findall(r"-?\d+u?", body)]

    steps = table("AADK_STEP_TABLE_INIT")
    d = np.arange(0, 1 << 17, dtype=np.uint64)
    for bits in (2, 3, 4):
        direct = table(f"AADK_STEP_DIRECT{bits}_INIT")
        tiny = int(re.search(rf"#define AADK_TINY_ROWS{bits} (\d+)", text).group(1))
        assert len(direct) == 256 and all(0 < m < (1 << 32) for m in direct)
        assert tiny == sum(1 for s in steps if s <= (1 << (bits - 2)))
        maxmag = (1 << (bits - 1)) - 1
        for row, (s, m) in enumerate(zip(steps, direct)):
            operand = d << np.uint64(bits - 1) if row < tiny else d
            fast = np.minimum((operand * np.uint64(m)) >> np.uint64(32), maxmag)
            assert np.array_equal(fast, np.minimum((d << np.uint64(bits - 2)) // np.uint64(s), maxmag)), (s, bits)


def test_magic_division_tables_are_exact():
    """The encoder's divide-free quantiser: umulhi(|d| << (b-1), M) >> L == (|d| << (b-2)) / step for
    every table step and every reachable |d| (|sample - predict| < 2^17), all bit depths."""
    import re
    text = (ROOT / "aad_b200" / "csrc" / "aad_tables_data.h").read_text()

    def table(name):
        body = re.search(name + r" \{(.*?)\}", text.replace("\\\n", " "), re.S).group(1)
        return [int(v.rstrip("u")) for v in re.findall(r"-?\d+u?", body)]

    steps, magic, shift = table("AADK_STEP_TABLE_INIT"), table("AADK_STEP_MAGIC_INIT"), table("AADK_STEP_SHIFT_INIT")
    assert len(steps) == len(magic) == len(shift) == 256
    d = np.arange(0, 1 << 17, dtype=np.uint64)
    for s, m, l in set(zip(steps, magic, shift)):
        assert (1 << 31) <= m < (1 << 32) and l == (s - 1).bit_length()
        for bits in (2, 3, 4):
            fast = (((d << np.uint64(bits - 1)) * np.uint64(m)) >> np.uint64(32)) >> np.uint64(l)
            assert np.array_equal(fast, (d << np.uint64(bits - 2)) // np.uint64(s)), (s, bits)


def test_pipeline_slice_boundaries_are_monotone_and_complete(tmp_path):
    """aad_gpu.c: slice_bound (block-range slices of AADGpu_ReconstructBatch, short first / last slices): for every
    (blocks >= slices) pair the boundaries start at 0, end at the block count and never go backwards -- and are strictly
    increasing (no empty slice) with the ramp.  The function is static: its text is compiled into a small checker."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    text = (Path(__file__).resolve().parents[1] / "aad_b200" / "csrc" / "aad_gpu.c").read_text()
    start, end = text.index("static uint32_t slice_bound("), text.index("static int slice_ramp_default(void)")
    src = "#include <stdint.h>\n#include <stdio.h>\n" + text[start:end] + r"""
int main(void)
{
  unsigned long bad = 0, checked = 0;
  for (uint32_t slices = 1; slices <= 64; slices++)
    for (uint32_t nblk = slices; nblk <= 3000; nblk += (nblk < 300 ? 1 : 37))
      for (int ramp = 0; ramp < 2; ramp++) {
        uint32_t prev = slice_bound(nblk, slices, 0, ramp);
        if (prev != 0) bad++;
        for (uint32_t k = 1; k <= slices; k++, checked++) {
          const uint32_t b = slice_bound(nblk, slices, k, ramp);
          if (b < prev || (b == prev && (ramp || slices <= nblk))) bad++;
          prev = b;
        }
        if (prev != nblk) bad++;
      }
  printf("%lu %lu\n", checked, bad);
  return 0;
}
"""
    c_file, exe = tmp_path / "slice_bound.c", tmp_path / "slice_bound"
    c_file.write_text(src)
    subprocess.run(["gcc", "-O2", "-o", str(exe), str(c_file)], check=True)
    checked, bad = map(int, subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split())
    assert checked > 1_000_000 and bad == 0


def test_avx2_widening_loop_equals_the_scalar_one(tmp_path):
    """aad_gpu.c: widen_avx2 (int16 ring -> the caller's int32 rows of the drop-in DecodeWhole) against a plain loop, every
    start offset of a 32-byte aligned destination, lengths around the 16-sample stride.  Compiled from the file's own text."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    text = (Path(__file__).resolve().parents[1] / "aad_b200" / "csrc" / "aad_gpu.c").read_text()
    start = text.index('__attribute__((target("avx2"))) static uint64_t widen_avx2')
    end = text.index("static int have_avx2(void)")
    src = "#include <stdint.h>\n#include <stdio.h>\n#include <stdlib.h>\n#include <immintrin.h>\n" + text[start:end] + r"""
int main(void)
{
  if (!__builtin_cpu_supports("avx2")) { printf("skip\n"); return 0; }
  enum { N = 4096 };
  int16_t *narrow = malloc(2 * N + 64);
  int32_t *wide = aligned_alloc(64, 4 * N + 64), *want = malloc(4 * N + 64);
  unsigned long bad = 0, checked = 0;
  uint32_t x = 12345;
  for (int i = 0; i < N; i++) { x = x * 1664525u + 1013904223u; narrow[i] = (int16_t)(x >> 16); want[i] = narrow[i]; }
  narrow[0] = -32768; narrow[1] = 32767; want[0] = -32768; want[1] = 32767;
  for (uint64_t t0 = 0; t0 < 64; t0 += 8)               /* wide + t0 is 32-byte aligned */
    for (uint64_t b = t0; b < t0 + 100; b++) {
      for (int i = 0; i < N; i++) wide[i] = 0x55555555;
      const uint64_t t = widen_avx2(narrow, wide, t0, b);
      _mm_sfence();
      if (t > b || t < t0 || (b - t) >= 16 || (t - t0) % 16) bad++;
      for (uint64_t i = 0; i < N; i++, checked++)
        if (wide[i] != ((i >= t0 && i < t) ? want[i] : 0x55555555)) bad++;
    }
  printf("%lu %lu\n", checked, bad);
  return 0;
}
"""
    c_file, exe = tmp_path / "widen.c", tmp_path / "widen"
    c_file.write_text(src)
    subprocess.run(["gcc", "-O2", "-o", str(exe), str(c_file)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    if out == ["skip"]:
        pytest.skip("this CPU has no AVX2")
    checked, bad = map(int, out)
    assert checked > 1_000_000 and bad == 0
