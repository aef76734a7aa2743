"""ctypes binding of the AAD codec C API (include/aad.h, aad_encoder.h, aad_decoder.h).

The same binding works for any shared library exporting the reference's 14 entry points
(src/aad_encoder.h:25-50, src/aad_decoder.h:15-42): libaad_b200.so (this repo's product) and,
in the tests only, the reference compiled into oracle/_ref/.  That is what makes the parity
tests read like the reference's own: identical calls against two libraries.
"""
import ctypes as C

import numpy as np

# AADApiResult, src/aad.h:25-33
OK, INVALID_ARGUMENT, INVALID_FORMAT, INSUFFICIENT_BUFFER, INSUFFICIENT_DATA, PARAMETER_NOT_SET, NG = range(7)
# AADChannelProcessMethod, src/aad.h:36-40
CH_NONE, CH_MS, CH_INVALID = range(3)

HEADER_SIZE = 31
FORMAT_VERSION = 4
CODEC_VERSION = 18
MIN_BITS, MAX_BITS = 2, 4


def block_header_size(channels):
    """AAD_BLOCK_HEADER_SIZE, src/aad_internal.h:37"""
    return 18 * channels


class HeaderInfo(C.Structure):
    """struct AADHeaderInfo, src/aad.h:43-53"""
    _fields_ = [
        ("format_version", C.c_uint32),
        ("codec_version", C.c_uint32),
        ("num_channels", C.c_uint16),
        ("num_samples", C.c_uint32),
        ("sampling_rate", C.c_uint32),
        ("bits_per_sample", C.c_uint16),
        ("block_size", C.c_uint16),
        ("num_samples_per_block", C.c_uint32),
        ("ch_process_method", C.c_int),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class EncodeParameter(C.Structure):
    """struct AADEncodeParameter, src/aad_encoder.h:8-15"""
    _fields_ = [
        ("num_channels", C.c_uint16),
        ("sampling_rate", C.c_uint32),
        ("bits_per_sample", C.c_uint16),
        ("max_block_size", C.c_uint16),
        ("ch_process_method", C.c_int),
        ("num_encode_trials", C.c_uint8),
    ]


def make_param(channels, rate, bits, max_block, ms=False, trials=2):
    return EncodeParameter(channels, rate, bits, max_block, CH_MS if ms else CH_NONE, trials)


def _planar_pointers(arrays):
    """int32_t *buf[C] from a list of contiguous int32 numpy arrays"""
    ptrs = (C.POINTER(C.c_int32) * len(arrays))()
    for i, a in enumerate(arrays):
        assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
        ptrs[i] = a.ctypes.data_as(C.POINTER(C.c_int32))
    return ptrs


class AADCApi:
    """The reference's public C API, bound from `path`."""

    SYMBOLS = (
        "AADEncoder_CalculateBlockSize", "AADEncoder_EncodeHeader", "AADEncoder_CalculateWorkSize",
        "AADEncoder_Create", "AADEncoder_Destroy", "AADEncoder_SetEncodeParameter", "AADEncoder_EncodeWhole",
        "AADDecoder_DecodeHeader", "AADDecoder_CalculateWorkSize", "AADDecoder_Create", "AADDecoder_Destroy",
        "AADDecoder_SetHeader", "AADDecoder_DecodeBlock", "AADDecoder_DecodeWhole",
    )

    def __init__(self, path):
        self.path = str(path)
        lib = self.lib = C.CDLL(self.path)
        u8p, u16p, u32p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint16), C.POINTER(C.c_uint32)
        i32pp = C.POINTER(C.POINTER(C.c_int32))
        hp, pp = C.POINTER(HeaderInfo), C.POINTER(EncodeParameter)
        sig = {
            "AADEncoder_CalculateBlockSize": (C.c_int, [C.c_uint16, C.c_uint16, C.c_uint32, u16p, u32p]),
            "AADEncoder_EncodeHeader": (C.c_int, [hp, u8p, C.c_uint32]),
            "AADEncoder_CalculateWorkSize": (C.c_int32, [C.c_uint16]),
            "AADEncoder_Create": (C.c_void_p, [C.c_uint16, C.c_void_p, C.c_int32]),
            "AADEncoder_Destroy": (None, [C.c_void_p]),
            "AADEncoder_SetEncodeParameter": (C.c_int, [C.c_void_p, pp]),
            "AADEncoder_EncodeWhole": (C.c_int, [C.c_void_p, i32pp, C.c_uint32, u8p, C.c_uint32, u32p]),
            "AADDecoder_DecodeHeader": (C.c_int, [u8p, C.c_uint32, hp]),
            "AADDecoder_CalculateWorkSize": (C.c_int32, []),
            "AADDecoder_Create": (C.c_void_p, [C.c_void_p, C.c_int32]),
            "AADDecoder_Destroy": (None, [C.c_void_p]),
            "AADDecoder_SetHeader": (C.c_int, [C.c_void_p, hp]),
            "AADDecoder_DecodeBlock": (C.c_int, [C.c_void_p, u8p, C.c_uint32, i32pp, C.c_uint32, C.c_uint32, u32p]),
            "AADDecoder_DecodeWhole": (C.c_int, [C.c_void_p, u8p, C.c_uint32, i32pp, C.c_uint32, C.c_uint32]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args

    # ---- convenience wrappers used by tests, bench and the CLI shim -------------------------
    def calculate_block_size(self, max_block, channels, bits):
        bs, spb = C.c_uint16(0), C.c_uint32(0)
        rc = self.lib.AADEncoder_CalculateBlockSize(max_block, channels, bits, C.byref(bs), C.byref(spb))
        return rc, bs.value, spb.value

    def encode_whole(self, pcm, rate, bits, max_block=1024, ms=False, trials=2, handle=None, capacity=None):
        """pcm: int array [channels, samples] with int16-range values.  Returns (rc, bytes)."""
        pcm = np.ascontiguousarray(np.asarray(pcm), dtype=np.int32)
        channels, n = pcm.shape
        own = handle is None
        if own:
            handle = self.lib.AADEncoder_Create(max_block, None, 0)
            assert handle, "AADEncoder_Create failed"
            rc = self.lib.AADEncoder_SetEncodeParameter(handle, C.byref(make_param(channels, rate, bits, max_block, ms, trials)))
            if rc != OK:
                self.lib.AADEncoder_Destroy(handle)
                return rc, b""
        cap = capacity if capacity is not None else max(4 * channels * n, 64) + 4096
        out = np.zeros(cap, dtype=np.uint8)
        size = C.c_uint32(0)
        rows = [pcm[c] for c in range(channels)]
        rc = self.lib.AADEncoder_EncodeWhole(handle, _planar_pointers(rows), n, out.ctypes.data_as(C.POINTER(C.c_uint8)),
                                             cap, C.byref(size))
        if own:
            self.lib.AADEncoder_Destroy(handle)
        return rc, (out[:size.value].tobytes() if rc == OK else b"")

    def decode_header(self, data):
        buf = np.frombuffer(bytes(data), dtype=np.uint8)
        h = HeaderInfo()
        rc = self.lib.AADDecoder_DecodeHeader(buf.ctypes.data_as(C.POINTER(C.c_uint8)), len(buf), C.byref(h))
        return rc, h

    def decode_whole(self, data, buf_channels=None, buf_samples=None, fill=0):
        """Returns (rc, int32 array [channels, buf_samples], HeaderInfo)."""
        rc, h = self.decode_header(data)
        if rc != OK:
            return rc, None, h
        ch = buf_channels if buf_channels is not None else h.num_channels
        n = buf_samples if buf_samples is not None else h.num_samples
        out = np.full((max(ch, 1), max(n, 1)), fill, dtype=np.int32)
        buf = np.frombuffer(bytes(data), dtype=np.uint8)
        dec = self.lib.AADDecoder_Create(None, 0)
        rows = [out[c] for c in range(out.shape[0])]
        rc = self.lib.AADDecoder_DecodeWhole(dec, buf.ctypes.data_as(C.POINTER(C.c_uint8)), len(buf),
                                             _planar_pointers(rows), ch, n)
        self.lib.AADDecoder_Destroy(dec)
        return rc, out[:ch, :n], h
