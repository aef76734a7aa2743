"""Host I/O path and command line: aad_wav.c and the `aad` binary.

CPU part: the WAV parser / writer against hand-built files (every bit depth, extra chunks, fmt
extension, truncation) and the reference's fixtures; the option parser's accepted spellings and
error exits (src/command_line_parser.c:173-331, src/main.c:518-625); `-i`.
GPU part (marked): `aad -e / -d / -r / -g / -c / --batch` produce the reference CLI's files byte for
byte -- test/sin300Hz*.aad and *_decoded.wav are outputs of the reference CLI itself
(test/make_test_data.sh:4-7).
"""
import ctypes as C
import struct
import subprocess

import numpy as np
import pytest

import aadtest
import aad_b200

CLI = aad_b200.PACKAGE_DIR / "aad"


class WavInfo(C.Structure):
    _fields_ = [("num_channels", C.c_uint32), ("sampling_rate", C.c_uint32), ("bits_per_sample", C.c_uint32),
                ("num_samples", C.c_uint32), ("data_offset", C.c_size_t)]


@pytest.fixture(scope="module")
def wavlib():
    lib = C.CDLL(str(aad_b200.LIBRARY_PATH))
    lib.aadwav_parse.restype = C.c_int
    lib.aadwav_parse.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(WavInfo)]
    lib.aadwav_sample32.restype = C.c_int32
    lib.aadwav_sample32.argtypes = [C.c_void_p, C.c_uint32, C.c_size_t]
    lib.aadwav_to_pcm16.restype = None
    lib.aadwav_to_pcm16.argtypes = [C.c_void_p, C.c_uint32, C.c_size_t, C.c_void_p]
    lib.aadwav_write_header.restype = C.c_size_t
    lib.aadwav_write_header.argtypes = [C.c_void_p] + [C.c_uint32] * 4
    lib.aadwav_store32.restype = None
    lib.aadwav_store32.argtypes = [C.c_void_p, C.c_uint32, C.c_size_t, C.c_int32]
    return lib


def build_wav(channels, rate, bits, frames, extra_chunks=(), fmt_extra=b"", declared_data=None):
    """frames: int array [n, channels] of `bits`-wide sample values (unsigned for 8 bit)."""
    frames = np.asarray(frames, dtype=np.int64).reshape(-1, channels)
    width = bits // 8
    data = b"".join(int(v & ((1 << bits) - 1)).to_bytes(width, "little") for v in frames.reshape(-1))
    fmt = struct.pack("<HHIIHH", 1, channels, rate, rate * width * channels, width * channels, bits) + fmt_extra
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt
    for cid, payload in extra_chunks:
        body += cid + struct.pack("<I", len(payload)) + payload
    body += b"data" + struct.pack("<I", len(data) if declared_data is None else declared_data) + data
    return b"RIFF" + struct.pack("<I", len(body)) + body


def parse(lib, image):
    buf = np.frombuffer(image, dtype=np.uint8).copy()
    info = WavInfo()
    rc = lib.aadwav_parse(buf.ctypes.data, len(buf), C.byref(info))
    return rc, info, buf


@pytest.mark.parametrize("bits", [8, 16, 24, 32])
@pytest.mark.parametrize("channels", [1, 2, 3])
def test_wav_parse_and_widen_every_depth(wavlib, bits, channels):
    rng = np.random.default_rng(bits + channels)
    n = 57
    if bits == 8:
        frames = rng.integers(0, 256, size=(n, channels))
        want32 = ((frames - 128) << 24).astype(np.int64)
    else:
        frames = rng.integers(-(1 << (bits - 1)), 1 << (bits - 1), size=(n, channels))
        want32 = frames << (32 - bits)
    image = build_wav(channels, 44100, bits, frames)
    rc, info, buf = parse(wavlib, image)
    assert rc == 0
    assert (info.num_channels, info.sampling_rate, info.bits_per_sample, info.num_samples) == (channels, 44100, bits, n)
    assert info.data_offset == 44
    data = buf.ctypes.data + info.data_offset
    got32 = np.array([wavlib.aadwav_sample32(data, bits, i) for i in range(n * channels)], dtype=np.int64).reshape(n, channels)
    assert np.array_equal(got32, want32)                       # src/wav.c:391-415
    pcm16 = np.zeros(n * channels, dtype=np.int16)
    wavlib.aadwav_to_pcm16(data, bits, n * channels, pcm16.ctypes.data)
    assert np.array_equal(pcm16.reshape(n, channels), want32 >> 16)   # src/main.c:175-179
    # writer: same bytes back (src/wav.c:418-436, :562-627)
    out = np.zeros(len(image), dtype=np.uint8)
    assert wavlib.aadwav_write_header(out.ctypes.data, channels, 44100, bits, n) == 44
    for i, v in enumerate(want32.reshape(-1)):
        wavlib.aadwav_store32(out.ctypes.data + 44, bits, i, int(v))
    assert out.tobytes() == image


def test_wav_skips_fmt_extension_and_unknown_chunks(wavlib):
    frames = np.arange(-20, 20).reshape(20, 2)
    image = build_wav(2, 8000, 16, frames, extra_chunks=[(b"LIST", b"x" * 26), (b"fact", b"\0" * 4)], fmt_extra=b"\0" * 6)
    rc, info, buf = parse(wavlib, image)
    assert rc == 0 and info.num_samples == 20 and info.data_offset == 44 + 6 + (8 + 26) + (8 + 4)
    pcm16 = np.zeros(40, dtype=np.int16)
    wavlib.aadwav_to_pcm16(buf.ctypes.data + info.data_offset, 16, 40, pcm16.ctypes.data)
    assert np.array_equal(pcm16, frames.reshape(-1))


def test_wav_rejects_what_the_reference_rejects(wavlib):
    good = build_wav(1, 8000, 16, np.zeros((4, 1)))
    assert parse(wavlib, good)[0] == 0
    assert parse(wavlib, b"RIFX" + good[4:])[0] == 2                     # not RIFF
    assert parse(wavlib, good[:8] + b"WAVX" + good[12:])[0] == 2         # not WAVE
    assert parse(wavlib, good[:12] + b"LIST" + good[16:])[0] == 2        # "fmt " must come first (src/wav.c:137)
    assert parse(wavlib, good[:20] + struct.pack("<H", 3) + good[22:])[0] == 2   # float format tag
    assert parse(wavlib, good[:34] + struct.pack("<H", 12) + good[36:])[0] == 2  # 12 bits per sample
    assert parse(wavlib, good[:22] + struct.pack("<H", 0) + good[24:])[0] == 2   # no channels
    assert parse(wavlib, good[:-3])[0] == 3                              # data shorter than declared
    assert parse(wavlib, good[:30])[0] == 3
    assert parse(wavlib, build_wav(1, 8000, 16, np.zeros((4, 1)), declared_data=6))[1].num_samples == 3


@pytest.mark.parametrize("stem", ["sin300Hz", "sin300Hz_mono", "bunny1", "pi_15-25sec", "sin300Hz_decoded"])
def test_wav_fixtures_parse_like_the_wave_module(wavlib, stem):
    image = (aadtest.GOLDEN / f"{stem}.wav").read_bytes()
    rc, info, buf = parse(wavlib, image)
    pcm, rate = aadtest.read_wav16(aadtest.GOLDEN / f"{stem}.wav")
    assert rc == 0 and info.sampling_rate == rate and (info.num_channels, info.num_samples) == pcm.shape
    got = np.zeros(pcm.size, dtype=np.int16)
    wavlib.aadwav_to_pcm16(buf.ctypes.data + info.data_offset, 16, pcm.size, got.ctypes.data)
    assert np.array_equal(got.reshape(-1, pcm.shape[0]).T, pcm)


# ---- command line, no device needed -----------------------------------------------------------------

def run_cli(*args, **kw):
    return subprocess.run([str(CLI), *map(str, args)], capture_output=True, text=True, timeout=300, **kw)


def test_cli_usage_help_version():
    r = run_cli()
    assert r.returncode == 1 and "Usage:" in r.stdout and "-h" in r.stdout
    r = run_cli("-h")
    assert r.returncode == 0 and "--bits-per-sample" in r.stdout and "--num-encode-trials" in r.stdout
    r = run_cli("--version")
    assert r.returncode == 0 and "Version.18" in r.stdout


@pytest.mark.parametrize("args,message", [
    (["in.wav", "out.aad"], "must specify at least one mode"),
    (["-e", "-d", "a", "b"], "multiple modes cannot specify simultaneously"),
    (["-ed", "a", "b"], "multiple modes cannot specify simultaneously"),
    (["-e"], "input file must be specified"),
    (["-e", "a.wav"], "output file must be specified"),
    (["-e", "-e", "a", "b"], "multiply specified"),
    (["--encode", "--encode", "a", "b"], "multiply specified"),
    (["-e", "-b"], "needs argument"),
    (["-e", "-b", "-m", "a", "b"], "needs argument"),
    (["-be", "4", "a", "b"], "tail of short option sequence"),
    (["-e", "--bits-per-sample"], "needs argument"),
    (["-x", "a", "b"], "Unknown short option"),
    (["--bogus", "a", "b"], "Unknown long option"),
    (["-e", "a", "b", "c"], "Too many strings"),
    (["-r", "--batch", "list.txt"], "--batch goes with -e or -d"),
])
def test_cli_argument_errors(args, message):
    r = run_cli(*args)
    assert r.returncode == 1 and message in r.stderr, (r.stdout, r.stderr)


def test_cli_information_matches_the_reference_layout():
    r = run_cli("-i", aadtest.GOLDEN / "sin300Hz.aad")
    assert r.returncode == 0
    lines = {l.split(":")[0]: l.split(":")[1].strip() for l in r.stdout.splitlines()}
    assert lines == {"Format Version": "4", "Codec Version": "18", "Number of Channels": "2",
                     "Number of Samples per Channel": "24000", "Sampling Rate": "48000", "Bits per Sample": "4",
                     "Block size": "1024", "Number of Samples per Block": "992", "Channel Processing": "None",
                     "Bits per Second(bps)": "396387.1"}
    # the exact format strings of src/main.c:260-269
    assert r.stdout.splitlines()[0] == "%-30s %-9d   " % ("Format Version:", 4)
    assert run_cli("--information", aadtest.GOLDEN / "sin300Hz.wav").returncode == 1       # not an .aad header
    assert "Failed to open" in run_cli("-i", "/nonexistent.aad").stderr


def test_cli_without_a_device_fails_loudly(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    r = run_cli("-e", aadtest.GOLDEN / "sin300Hz.wav", tmp_path / "o.aad")
    assert r.returncode == 1 and "no CPU fallback" in r.stderr and not (tmp_path / "o.aad").exists()


# ---- command line on the GPU ---------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("stem", ["sin300Hz", "sin300Hz_mono"])
def test_cli_encode_decode_reproduce_the_reference_files(tmp_path, stem):
    """test/make_test_data.sh:4-7: `aad -e X.wav X.aad`, `aad -d X.aad X_decoded.wav` with the stock CLI."""
    r = run_cli("-e", aadtest.GOLDEN / f"{stem}.wav", tmp_path / "o.aad")
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "o.aad").read_bytes() == (aadtest.GOLDEN / f"{stem}.aad").read_bytes()
    r = run_cli("--decode", aadtest.GOLDEN / f"{stem}.aad", tmp_path / "o.wav")
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "o.wav").read_bytes() == (aadtest.GOLDEN / f"{stem}_decoded.wav").read_bytes()


@pytest.mark.gpu
def test_cli_options_reach_the_encoder(tmp_path, oracle):
    pcm, rate = aadtest.read_wav16(aadtest.GOLDEN / "pi_15-25sec.wav")
    for args, (bits, block, ms, trials) in [(["-b", "3", "-s", "256", "-m", "-t", "1"], (3, 256, True, 1)),
                                            (["--bits-per-sample=2", "--num-encode-trials", "0"], (2, 1024, False, 0))]:
        r = run_cli("-e", *args, aadtest.GOLDEN / "pi_15-25sec.wav", tmp_path / "o.aad")
        assert r.returncode == 0, r.stderr
        rc, want = oracle.encode(pcm, rate, bits, block, ms, trials)
        assert rc == 0 and (tmp_path / "o.aad").read_bytes() == want
    r = run_cli("-e", "-b", "5", aadtest.GOLDEN / "sin300Hz.wav", tmp_path / "bad.aad")      # src/aad_encoder.c:743-746
    assert r.returncode == 1 and "Failed to set encode parameter" in r.stderr


REF_CLI = aadtest.ROOT / "oracle" / "_ref" / "aad_ref_cli"


def analysis_case(tmp_path, oracle, wavlib, bits):
    """A stereo input of the given bit depth and what -r / -g / -c must produce for it, derived from the
    oracle codec and src/main.c:349-503: the codec sees the top 16 bits, -r / -g write the input's own
    format, -c prints the reference's (peculiar) statistic."""
    rng = np.random.default_rng(bits)
    n, ch, rate = 3000, 2, 22050
    base = aadtest.signal("music", ch, n, 3).T.astype(np.int64)                  # [n, ch] int16 values
    if bits == 8:
        frames = ((base >> 8) + 128)
        pcm32 = ((frames - 128) << 24)
    else:
        frames = (base << (bits - 16)) + (rng.integers(0, 1 << (bits - 16), size=base.shape) if bits > 16 else 0)
        pcm32 = frames << (32 - bits)
    src = tmp_path / f"in{bits}.wav"
    src.write_bytes(build_wav(ch, rate, bits, frames))
    pcm16 = (pcm32 >> 16).astype(np.int16).T                                     # what the codec is fed
    rc, data = oracle.encode(pcm16, rate, 4, 1024, False, 2)
    rc2, dec, _ = oracle.decode(data)
    assert rc == 0 and rc2 == 0
    recon32 = dec.T.astype(np.int64) << 16

    def image(values32):
        out = np.zeros(44 + n * ch * (bits // 8), dtype=np.uint8)
        wavlib.aadwav_write_header(out.ctypes.data, ch, rate, bits, n)
        wrapped = ((values32 + (1 << 31)) % (1 << 32)) - (1 << 31)
        for i, v in enumerate(wrapped.reshape(-1)):
            wavlib.aadwav_store32(out.ctypes.data + 44, bits, i, int(v))
        return out.tobytes()

    residual = ((pcm32 - recon32 + (1 << 31)) % (1 << 32)) - (1 << 31)
    err = residual / 2147483647.0 - dec.T.astype(np.float64) / 2147483647.0
    stats = "RMSE:%f MSD:%f MaxAE:%f" % (np.sqrt(np.mean(err ** 2)), np.mean(np.abs(err)), np.max(np.abs(err)))
    return src, data, image(recon32), image(pcm32 - recon32), stats


def check_analysis_modes(cli, tmp_path, case):
    src, data, want_r, want_g, want_c = case
    run = lambda *a: subprocess.run([str(cli), *map(str, a)], capture_output=True, text=True, timeout=300)
    assert run("-e", src, tmp_path / "e.aad").returncode == 0
    assert (tmp_path / "e.aad").read_bytes() == data
    assert run("-r", src, tmp_path / "r.wav").returncode == 0
    assert (tmp_path / "r.wav").read_bytes() == want_r
    assert run("-g", src, tmp_path / "g.wav").returncode == 0
    assert (tmp_path / "g.wav").read_bytes() == want_g
    r = run("-c", src)
    assert r.returncode == 0 and r.stdout.strip() == want_c


@pytest.mark.parametrize("bits", [8, 16, 24, 32])
def test_the_model_of_the_analysis_modes_is_the_reference_cli(tmp_path, oracle, wavlib, bits):
    """Pins analysis_case() -- and with it aad_wav.c's conversions -- against the UNMODIFIED reference
    command line compiled into oracle/_ref (CPU only, where /root/reference exists)."""
    if not REF_CLI.exists():
        pytest.skip("oracle/_ref/aad_ref_cli not built (the reference sources are not on this machine)")
    check_analysis_modes(REF_CLI, tmp_path, analysis_case(tmp_path, oracle, wavlib, bits))


@pytest.mark.gpu
@pytest.mark.parametrize("bits", [8, 16, 24, 32])
def test_cli_reconstruct_gap_calculate(tmp_path, oracle, wavlib, bits):
    check_analysis_modes(CLI, tmp_path, analysis_case(tmp_path, oracle, wavlib, bits))


@pytest.mark.gpu
@pytest.mark.parametrize("bits,channels", [(8, 1), (16, 2), (24, 3), (32, 2)])
def test_analyze_wav_api_against_numpy_model(product, gpu_ctx, oracle, bits, channels):
    """AADGpu_AnalyzeWav on a data chunk long enough that every reduction thread sums many samples: -r / -g images
    byte for byte, MaxAE exactly, the two sums of -c to 1e-12 relative (their order of summation is the one documented
    difference from src/main.c:470-497)."""
    import ctypes as C
    from aad_b200.capi import make_param, OK
    _, gpu = product
    rng = np.random.default_rng(100 + bits)
    n, rate = 700_001, 32000
    base = aadtest.signal("music", channels, n, bits).T.astype(np.int64)         # [n, ch]
    if bits == 8:
        frames = (base >> 8) + 128
        pcm32 = (frames - 128) << 24
        raw = frames.astype(np.uint8).tobytes()
    else:
        frames = (base << (bits - 16)) + (rng.integers(0, 1 << (bits - 16), size=base.shape) if bits > 16 else 0)
        pcm32 = frames << (32 - bits)
        le = np.ascontiguousarray(frames.astype("<i8")).view(np.uint8).reshape(n, channels, 8)[:, :, :bits // 8]
        raw = np.ascontiguousarray(le).tobytes()
    pcm16 = (pcm32 >> 16).astype(np.int16).T
    rc, data = oracle.encode(pcm16, rate, 3, 1024, False, 1)
    rc2, dec, _ = oracle.decode(data)
    assert rc == 0 and rc2 == 0
    recon32 = dec.T.astype(np.int64) << 16
    wrap = lambda v: ((v + (1 << 31)) % (1 << 32)) - (1 << 31)

    def narrow(values32):                                                        # src/wav.c:418-436
        v = wrap(values32) >> (32 - bits)
        if bits == 8:
            return ((v + 128) & 0xFF).astype(np.uint8).tobytes()
        return np.ascontiguousarray(np.ascontiguousarray(v.astype("<i8")).view(np.uint8).reshape(n, channels, 8)[:, :, :bits // 8]).tobytes()

    prm = make_param(channels, rate, 3, 1024, False, 1)
    src = np.frombuffer(raw, dtype=np.uint8).copy()
    out = np.zeros_like(src)
    size = C.c_uint32(0)
    assert gpu.lib.AADGpu_AnalyzeWav(gpu_ctx, C.byref(prm), src.ctypes.data, bits, n, 0, out.ctypes.data, None, C.byref(size)) == OK, gpu.last_error()
    assert size.value == len(data) and out.tobytes() == narrow(recon32)
    enc = np.zeros(len(data) + 16, dtype=np.uint8)                               # AADGpu_EncodeWav: same chunk -> the oracle's stream
    assert gpu.lib.AADGpu_EncodeWav(gpu_ctx, C.byref(prm), src.ctypes.data, bits, n, enc.ctypes.data, len(enc), C.byref(size)) == OK
    assert enc[:size.value].tobytes() == data
    assert gpu.lib.AADGpu_EncodeWav(gpu_ctx, C.byref(prm), src.ctypes.data, bits, n, enc.ctypes.data, 100, C.byref(size)) == 3   # INSUFFICIENT_BUFFER
    assert gpu.lib.AADGpu_AnalyzeWav(gpu_ctx, C.byref(prm), src.ctypes.data, bits, n, 1, out.ctypes.data, None, None) == OK
    assert out.tobytes() == narrow(pcm32 - recon32)
    stats = (C.c_double * 3)()
    assert gpu.lib.AADGpu_AnalyzeWav(gpu_ctx, C.byref(prm), src.ctypes.data, bits, n, 2, None, stats, None) == OK
    err = wrap(pcm32 - recon32) / 2147483647.0 - dec.T.astype(np.float64) / 2147483647.0
    assert stats[2] == np.max(np.abs(err))
    assert abs(stats[0] - np.sqrt(np.mean(err ** 2))) <= 1e-12 * stats[0]
    assert abs(stats[1] - np.mean(np.abs(err))) <= 1e-12 * stats[1]
    # argument checks
    assert gpu.lib.AADGpu_AnalyzeWav(gpu_ctx, C.byref(prm), src.ctypes.data, 12, n, 0, out.ctypes.data, None, None) == 2     # INVALID_FORMAT
    assert gpu.lib.AADGpu_AnalyzeWav(gpu_ctx, C.byref(prm), src.ctypes.data, bits, n, 2, None, None, None) == 1               # INVALID_ARGUMENT
    assert gpu.lib.AADGpu_AnalyzeWav(gpu_ctx, C.byref(prm), src.ctypes.data, bits, n, 0, None, stats, None) == 1


@pytest.mark.gpu
def test_cli_differential_against_the_reference_cli(tmp_path):
    """Every mode, fixtures of the reference: same files, same stdout as the stock CLI run beside it."""
    if not REF_CLI.exists():
        pytest.skip("oracle/_ref/aad_ref_cli did not travel")
    for stem in ("sin300Hz", "bunny1", "pi_15-25sec"):
        src = aadtest.GOLDEN / f"{stem}.wav"
        for opts in ([], ["-b", "3", "-t", "0"], ["-b", "2", "-s", "300"] + (["-m"] if stem != "bunny1" else [])):
            out = {}
            for name, cli in (("ref", REF_CLI), ("b200", CLI)):
                d = tmp_path / name
                d.mkdir(exist_ok=True)
                run = lambda *a: subprocess.run([str(cli), *map(str, a)], capture_output=True, text=True, timeout=600)
                assert run("-e", *opts, src, d / "e.aad").returncode == 0
                assert run("-d", d / "e.aad", d / "d.wav").returncode == 0
                assert run("-r", *opts, src, d / "r.wav").returncode == 0
                assert run("-g", *opts, src, d / "g.wav").returncode == 0
                c = run("-c", *opts, src)
                i = run("-i", d / "e.aad")
                assert c.returncode == 0 and i.returncode == 0
                out[name] = [(d / f).read_bytes() for f in ("e.aad", "d.wav", "r.wav", "g.wav")] + [c.stdout, i.stdout]
            assert out["ref"] == out["b200"], (stem, opts)


@pytest.mark.gpu
def test_the_unmodified_reference_main_runs_on_the_b200_library(tmp_path):
    """INTEGRATION.md section 1, executed: src/main.c + wav.c + command_line_parser.c compiled untouched
    and linked against libaad_b200.so instead of the reference codec (oracle/Makefile:
    _ref/aad_ref_cli_b200) reproduce the reference's own output files."""
    drop_in = aadtest.ROOT / "oracle" / "_ref" / "aad_ref_cli_b200"
    if not drop_in.exists():
        pytest.skip("oracle/_ref/aad_ref_cli_b200 did not travel")
    run = lambda *a: subprocess.run([str(drop_in), *map(str, a)], capture_output=True, text=True, timeout=300)
    for stem in ("sin300Hz", "sin300Hz_mono"):
        r = run("-e", aadtest.GOLDEN / f"{stem}.wav", tmp_path / "o.aad")
        assert r.returncode == 0, r.stderr
        assert (tmp_path / "o.aad").read_bytes() == (aadtest.GOLDEN / f"{stem}.aad").read_bytes()
        r = run("-d", aadtest.GOLDEN / f"{stem}.aad", tmp_path / "o.wav")
        assert r.returncode == 0, r.stderr
        assert (tmp_path / "o.wav").read_bytes() == (aadtest.GOLDEN / f"{stem}_decoded.wav").read_bytes()
    r = run("-c", "-b", "3", aadtest.GOLDEN / "bunny1.wav")
    want = subprocess.run([str(REF_CLI), "-c", "-b", "3", str(aadtest.GOLDEN / "bunny1.wav")], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout == want.stdout


@pytest.mark.gpu
def test_cli_batch_mode_equals_file_by_file(tmp_path, oracle):
    """--batch: files of one shape share a launch; every output equals the single-file result."""
    rng = np.random.default_rng(5)
    pairs, expect = [], {}
    for i in range(9):
        ch, rate = (1, 16000) if i % 3 == 0 else ((2, 44100) if i % 3 == 1 else (2, 48000))
        n = int(rng.integers(5, 9000))
        pcm = aadtest.signal(aadtest.SIGNALS[i % len(aadtest.SIGNALS)], ch, n, i)
        src = tmp_path / f"in{i}.wav"
        src.write_bytes(build_wav(ch, rate, 16, pcm.T))
        rc, want = oracle.encode(pcm, rate, 3, 512, False, 2)
        assert rc == 0
        pairs.append((src, tmp_path / f"out{i}.aad"))
        expect[i] = (want, pcm.shape, rate)
    manifest = tmp_path / "enc.txt"
    manifest.write_text("# input output\n" + "".join(f"{a} {b}\n" for a, b in pairs) + "\n")
    r = run_cli("-e", "-b", "3", "-s", "512", "--batch", manifest)
    assert r.returncode == 0, r.stderr
    for i, (_, out) in enumerate(pairs):
        assert out.read_bytes() == expect[i][0], i
    manifest = tmp_path / "dec.txt"
    manifest.write_text("".join(f"{b} {tmp_path / f'dec{i}.wav'}\n" for i, (_, b) in enumerate(pairs)))
    r = run_cli("-d", "--batch", manifest)
    assert r.returncode == 0, r.stderr
    for i in range(9):
        assert run_cli("-d", pairs[i][1], tmp_path / "single.wav").returncode == 0
        assert (tmp_path / f"dec{i}.wav").read_bytes() == (tmp_path / "single.wav").read_bytes(), i
        _, dec, _ = oracle.decode(expect[i][0])
        got, rate = aadtest.read_wav16(tmp_path / f"dec{i}.wav")
        assert rate == expect[i][2] and np.array_equal(got, dec), i
