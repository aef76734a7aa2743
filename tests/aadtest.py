"""Shared test helpers: oracle binding, signal generators, WAV fixtures.

The oracle (oracle/liboracle.so, and oracle/_ref/*.so = the compiled reference) is test
infrastructure: it is imported here and nowhere in the product package."""
import ctypes as C
import hashlib
import json
import wave
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
GOLDEN = ROOT / "tests" / "golden"


class OracleInfo(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "format_version", "codec_version", "channels", "num_samples", "sampling_rate", "bits", "block_size",
        "samples_per_block", "ms")]


class OracleChain(C.Structure):
    _fields_ = [("weight", C.c_int32 * 4), ("stepsize_index", C.c_int32)]


class Oracle:
    """oracle/aad_oracle.h"""

    def __init__(self, path):
        lib = self.lib = C.CDLL(str(path))
        lib.aad_oracle_geometry.restype = C.c_int
        lib.aad_oracle_geometry.argtypes = [C.c_uint32] * 3 + [C.POINTER(C.c_uint32)] * 2
        lib.aad_oracle_encode.restype = C.c_int64
        lib.aad_oracle_encode.argtypes = [C.c_void_p, C.c_size_t] + [C.c_uint32] * 7 + [C.c_void_p, C.c_void_p, C.c_size_t]
        lib.aad_oracle_decode.restype = C.c_int
        lib.aad_oracle_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32,
                                          C.POINTER(OracleInfo)]
        lib.aad_oracle_read_header.restype = C.c_int
        lib.aad_oracle_read_header.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(OracleInfo), C.c_int]
        lib.aad_oracle_write_header.restype = C.c_int
        lib.aad_oracle_write_header.argtypes = [C.POINTER(OracleInfo), C.c_void_p, C.c_size_t]
        lib.aad_oracle_encode_batch.restype = C.c_int
        lib.aad_oracle_encode_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t] + [C.c_uint32] * 8 + [
            C.c_void_p, C.c_size_t, C.c_void_p]
        lib.aad_oracle_decode_batch.restype = C.c_int
        lib.aad_oracle_decode_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t,
                                                C.c_size_t, C.c_uint32, C.c_uint32]

    def geometry(self, max_block, channels, bits):
        bs, spb = C.c_uint32(0), C.c_uint32(0)
        rc = self.lib.aad_oracle_geometry(max_block, channels, bits, C.byref(bs), C.byref(spb))
        return rc, bs.value, spb.value

    def encode(self, pcm, rate, bits, max_block=1024, ms=False, trials=2, state=None):
        """pcm int16 [channels, n] -> (rc, bytes).  state: list of (w0..w3, idx) per channel, updated in place."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        ch, n = pcm.shape
        cap = 31 + 4 * ch * n + 4096
        out = np.zeros(cap, dtype=np.uint8)
        st = None
        if state is not None:
            st = (OracleChain * ch)()
            for c in range(ch):
                for k in range(4):
                    st[c].weight[k] = int(state[c][k])
                st[c].stepsize_index = int(state[c][4])
        r = self.lib.aad_oracle_encode(pcm.ctypes.data, n, ch, n, rate, bits, max_block, int(ms), trials,
                                       C.cast(st, C.c_void_p) if st is not None else None, out.ctypes.data, cap)
        if state is not None:
            for c in range(ch):
                state[c][:] = [st[c].weight[k] for k in range(4)] + [st[c].stepsize_index]
        if r < 0:
            return int(-r), b""
        return 0, out[:r].tobytes()

    def decode(self, data, buf_channels=None, buf_samples=None, fill=0):
        """-> (rc, int16 [channels, buf_samples], OracleInfo)"""
        buf = np.frombuffer(bytes(data), dtype=np.uint8)
        info = OracleInfo()
        rc = self.lib.aad_oracle_read_header(buf.ctypes.data, len(buf), C.byref(info), 1)
        if rc != 0:
            return rc, None, info
        ch = buf_channels if buf_channels is not None else info.channels
        n = buf_samples if buf_samples is not None else info.num_samples
        out = np.full((max(ch, 1), max(n, 1)), fill, dtype=np.int16)
        rc = self.lib.aad_oracle_decode(buf.ctypes.data, len(buf), out.ctypes.data, out.shape[1], ch, n, C.byref(info))
        return rc, out[:ch, :n], info


# ---- fixtures on disk -------------------------------------------------------------------------

def read_wav16(path):
    """-> (int16 [channels, samples], rate).  The reference CLI feeds the codec (PCM >> 16) of its
    32-bit-widened samples, i.e. exactly these int16 values (src/main.c:175-179)."""
    with wave.open(str(path)) as w:
        assert w.getsampwidth() == 2
        ch, rate, n = w.getnchannels(), w.getframerate(), w.getnframes()
        data = np.frombuffer(w.readframes(n), dtype="<i2").reshape(n, ch).T
    return np.ascontiguousarray(data), rate


def golden_table():
    return json.loads((GOLDEN / "golden.json").read_text())


def sha(data):
    return hashlib.sha256(bytes(data)).hexdigest()


def pcm_sha(pcm):
    """hash of decoded PCM as little-endian int16, channel-major"""
    return sha(np.ascontiguousarray(pcm, dtype="<i2").tobytes())


# ---- signals ----------------------------------------------------------------------------------

def signal(kind, channels, n, seed=0):
    """int16 [channels, n].  sine / noise / nyquist follow test/test_aad_encode_decode.c:428-471."""
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    if kind == "sine":
        x = np.stack([(32767 * 0.5 * np.sin(2.0 * 3.1415 * 440.0 * (c + 1) * t / 48000.0)).astype(np.int64)
                      for c in range(channels)])
    elif kind == "noise":          # full-scale white noise: overflows the predictor MAC
        x = rng.integers(-32767, 32768, size=(channels, n))
    elif kind == "nyquist":        # full-scale square at fs/2
        x = np.stack([np.where(t % 2 == 1, -32768, 32767)] * channels)
    elif kind == "silence":
        x = np.zeros((channels, n), dtype=np.int64)
    elif kind == "impulse":
        x = np.zeros((channels, n), dtype=np.int64)
        x[:, min(10, n - 1)] = 32767
    elif kind == "music":          # two partials + noise, different per channel
        x = np.stack([(9000 * np.sin(2 * np.pi * (220.0 * (c + 1)) * t / 44100.0)
                       + 5000 * np.sin(2 * np.pi * 3001.0 * t / 44100.0)).astype(np.int64)
                      + rng.integers(-800, 801, size=n) for c in range(channels)])
    elif kind == "steps":          # slow square between extremes: saturates weights / shift field
        x = np.stack([np.where((t // 37) % 2 == 0, 30000, -30000)] * channels) + rng.integers(-50, 51, size=(channels, n))
    elif kind == "fades":          # noise whose level sweeps 0 .. loud .. 0 in bursts: the step index walks through
        level = np.abs(np.sin(np.pi * t / 700.0)) ** 6 * (3.0 ** (seed % 9))   # the first rows of the step table and back
        x = np.rint(rng.standard_normal((channels, n)) * level).astype(np.int64)
    elif kind == "whisper":        # +-1 .. +-6 LSB noise with silent gaps: steps 1 .. 8
        x = rng.integers(-(1 + seed % 6), 2 + seed % 6, size=(channels, n)) * ((t // 53) % 3 != 0)
    else:
        raise ValueError(kind)
    return np.clip(x, -32768, 32767).astype(np.int16)


SIGNALS = ("sine", "noise", "nyquist", "silence", "impulse", "music", "steps")
QUIET_SIGNALS = ("fades", "whisper", "silence", "impulse")
