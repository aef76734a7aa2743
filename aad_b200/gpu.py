"""ctypes binding of the extended C ABI (include/aad_b200.h): batches of streams, int16 PCM,
host-pipelined and device-resident entry points.  Device pointers are plain integers (e.g.
``torch.Tensor.data_ptr()``); streams are ``cudaStream_t`` handles as integers (e.g.
``torch.cuda.current_stream().cuda_stream``)."""
import ctypes as C

import numpy as np

from .capi import OK, EncodeParameter, make_param


class GpuBatch(C.Structure):
    """struct AADGpuBatch, include/aad_b200.h"""
    _fields_ = [
        ("num_streams", C.c_uint32),
        ("num_samples", C.c_uint32),
        ("param", EncodeParameter),
        ("pcm_stream_stride", C.c_uint64),
        ("pcm_channel_stride", C.c_uint64),
        ("aad_stream_stride", C.c_uint64),
    ]


class AADError(RuntimeError):
    def __init__(self, where, rc, detail=""):
        super().__init__(f"{where} failed: AADApiResult={rc} {detail}".strip())
        self.rc = rc


class GpuApi:
    SYMBOLS = (
        "AADGpu_DeviceCount", "AADGpu_Create", "AADGpu_Destroy", "AADGpu_LastError", "AADGpu_KernelLaunchCount", "AADGpu_TmaLaunchCount",
        "AADGpu_SetMaxChannels", "AADGpu_GetMaxChannels", "AADGpu_HostAlloc", "AADGpu_HostFree", "AADGpu_BindHostThread",
        "AADGpu_StreamBytesBound", "AADGpu_StreamBytes", "AADGpu_EncodeBatchDevice", "AADGpu_DecodeBatchDevice",
        "AADGpu_EncodeBatch", "AADGpu_DecodeBatch", "AADGpu_ReconstructBatch", "AADGpu_SynthBatchDevice", "AADGpu_Deinterleave16Device",
        "AADGpu_Interleave16Device", "AADGpu_SynthLut", "AADGpu_SetKernelPath", "AADGpu_SetEncoderPairing", "AADGpu_SetEncoderSchedule",
        "AADGpu_SetEncodeSegmentBlocks", "AADGpu_GetEncodeSegmentBlocks",
        "AADGpu_EncodeInterleaved16", "AADGpu_DecodeInterleaved16", "AADGpu_ReconstructInterleaved16", "AADGpu_AnalyzeWav", "AADGpu_EncodeWav",
        "AADGpuGroup_Create", "AADGpuGroup_Destroy", "AADGpuGroup_Size", "AADGpuGroup_Device",
        "AADGpuGroup_EncodeBatch", "AADGpuGroup_DecodeBatch", "AADGpuGroup_DecodeInterleaved16",
        "AADGpuGroup_EncodeInterleaved16", "AADGpu_LinkProbe", "AADGpu_LinkProbeRows", "AADGpu_CopyProbeBatch",
    )

    def __init__(self, lib):
        self.lib = lib
        vp, u64, u32 = C.c_void_p, C.c_uint64, C.c_uint32
        bp, pp = C.POINTER(GpuBatch), C.POINTER(EncodeParameter)
        sig = {
            "AADGpu_DeviceCount": (C.c_int, []),
            "AADGpu_Create": (vp, [C.c_int]),
            "AADGpu_Destroy": (None, [vp]),
            "AADGpu_LastError": (C.c_char_p, []),
            "AADGpu_KernelLaunchCount": (u64, []),
            "AADGpu_TmaLaunchCount": (u64, []),
            "AADGpu_SetMaxChannels": (None, [u32]),
            "AADGpu_GetMaxChannels": (u32, []),
            "AADGpu_HostAlloc": (vp, [C.c_size_t]),
            "AADGpu_HostFree": (None, [vp]),
            "AADGpu_BindHostThread": (C.c_int, [vp]),
            "AADGpu_StreamBytesBound": (u64, [pp, u32]),
            "AADGpu_StreamBytes": (u64, [pp, u32]),
            "AADGpu_EncodeBatchDevice": (C.c_int, [vp, bp, vp, vp, vp, vp, vp]),
            "AADGpu_DecodeBatchDevice": (C.c_int, [vp, bp, vp, vp, vp, vp]),
            "AADGpu_EncodeBatch": (C.c_int, [vp, bp, vp, vp, vp, vp]),
            "AADGpu_DecodeBatch": (C.c_int, [vp, bp, vp, vp, vp]),
            "AADGpu_ReconstructBatch": (C.c_int, [vp, bp, vp, vp, vp, vp, vp]),
            "AADGpu_SynthBatchDevice": (C.c_int, [vp, bp, u32, vp, vp]),
            "AADGpu_Deinterleave16Device": (C.c_int, [vp, vp, vp, u64, u32, u32, vp]),
            "AADGpu_Interleave16Device": (C.c_int, [vp, vp, u64, vp, u32, u32, vp]),
            "AADGpu_SynthLut": (None, [vp]),
            "AADGpu_SetKernelPath": (None, [C.c_int]),
            "AADGpu_SetEncoderPairing": (None, [C.c_int]),
            "AADGpu_SetEncoderSchedule": (None, [C.c_int]),
            "AADGpu_SetEncodeSegmentBlocks": (C.c_int, [C.c_void_p, C.c_uint32]),
            "AADGpu_GetEncodeSegmentBlocks": (C.c_uint32, [C.c_void_p]),
            "AADGpu_EncodeInterleaved16": (C.c_int, [vp, pp, vp, u32, vp, u32, C.POINTER(u32)]),
            "AADGpu_DecodeInterleaved16": (C.c_int, [vp, vp, u32, vp, u32]),
            "AADGpu_ReconstructInterleaved16": (C.c_int, [vp, pp, vp, u32, vp, C.POINTER(u32)]),
            "AADGpu_EncodeWav": (C.c_int, [vp, pp, vp, u32, u32, vp, u32, C.POINTER(u32)]),
            "AADGpu_AnalyzeWav": (C.c_int, [vp, pp, vp, u32, u32, C.c_int, vp, C.POINTER(C.c_double), C.POINTER(u32)]),
            "AADGpuGroup_Create": (vp, [C.POINTER(C.c_int), C.c_int]),
            "AADGpuGroup_Destroy": (None, [vp]),
            "AADGpuGroup_Size": (C.c_int, [vp]),
            "AADGpuGroup_Device": (vp, [vp, C.c_int]),
            "AADGpuGroup_EncodeBatch": (C.c_int, [vp, bp, vp, vp, vp, vp]),
            "AADGpuGroup_DecodeBatch": (C.c_int, [vp, bp, vp, vp, vp]),
            "AADGpuGroup_DecodeInterleaved16": (C.c_int, [vp, vp, u32, vp, u32]),
            "AADGpuGroup_EncodeInterleaved16": (C.c_int, [vp, pp, u32, vp, u32, vp, u32, C.POINTER(u32)]),
            "AADGpu_LinkProbe": (C.c_int, [vp, C.c_size_t, C.c_int, C.POINTER(C.c_double)]),
            "AADGpu_CopyProbeBatch": (C.c_int, [vp, bp, vp, vp, vp]),
            "AADGpu_LinkProbeRows": (C.c_int, [vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.POINTER(C.c_double)]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args

    # ---- helpers ----------------------------------------------------------------------------
    def last_error(self):
        return (self.lib.AADGpu_LastError() or b"").decode()

    def device_count(self):
        return self.lib.AADGpu_DeviceCount()

    def create(self, device=0):
        h = self.lib.AADGpu_Create(device)
        if not h:
            raise AADError("AADGpu_Create", -1, self.last_error())
        return h

    def destroy(self, h):
        self.lib.AADGpu_Destroy(h)

    def launch_count(self):
        return int(self.lib.AADGpu_KernelLaunchCount())

    def stream_bytes_bound(self, param, num_samples):
        return int(self.lib.AADGpu_StreamBytesBound(C.byref(param), num_samples))

    def stream_bytes(self, param, num_samples):
        return int(self.lib.AADGpu_StreamBytes(C.byref(param), num_samples))

    def synth_lut(self):
        lut = np.zeros(1024, dtype=np.int16)
        self.lib.AADGpu_SynthLut(lut.ctypes.data)
        return lut

    def batch(self, num_streams, num_samples, param, pcm_channel_stride=None, pcm_stream_stride=None,
              aad_stream_stride=None):
        ch_stride = pcm_channel_stride if pcm_channel_stride is not None else num_samples
        st_stride = pcm_stream_stride if pcm_stream_stride is not None else ch_stride * param.num_channels
        a_stride = aad_stream_stride if aad_stream_stride is not None else self.stream_bytes_bound(param, num_samples)
        return GpuBatch(num_streams, num_samples, param, st_stride, ch_stride, a_stride)

    def pinned(self, shape, dtype):
        """numpy array over cudaMallocHost memory; release with free_pinned(arr)."""
        dtype = np.dtype(dtype)
        count = int(np.prod(shape))
        nbytes = max(count * dtype.itemsize, 1)
        p = self.lib.AADGpu_HostAlloc(nbytes)
        if not p:
            raise AADError("AADGpu_HostAlloc", -1, self.last_error())
        buf = (C.c_uint8 * nbytes).from_address(p)
        arr = np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)
        self._pinned = getattr(self, "_pinned", {})
        self._pinned[arr.ctypes.data] = p
        return arr

    def free_pinned(self, arr):
        p = getattr(self, "_pinned", {}).pop(arr.ctypes.data, None)
        if p:
            self.lib.AADGpu_HostFree(p)

    # ---- host batch calls ----------------------------------------------------------------------
    def encode_batch(self, h, pcm, rate, bits, max_block=1024, ms=False, trials=2, num_samples=None, out=None):
        """pcm: int16 array [streams, channels, samples] (C-contiguous).  Returns (aad uint8
        [streams, stride], sizes uint32 [streams])."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        n_streams, channels, n = pcm.shape
        prm = make_param(channels, rate, bits, max_block, ms, trials)
        b = self.batch(n_streams, n, prm)
        if b.aad_stream_stride == 0:
            raise AADError("AADGpu_StreamBytesBound", 2, "rejected parameters")
        aad = out if out is not None else np.zeros((n_streams, b.aad_stream_stride), dtype=np.uint8)
        sizes = np.zeros(n_streams, dtype=np.uint32)
        lens = None if num_samples is None else np.ascontiguousarray(num_samples, dtype=np.uint32)
        rc = self.lib.AADGpu_EncodeBatch(h, C.byref(b), pcm.ctypes.data, None if lens is None else lens.ctypes.data,
                                         aad.ctypes.data, sizes.ctypes.data)
        if rc != OK:
            raise AADError("AADGpu_EncodeBatch", rc, self.last_error())
        return aad, sizes

    def decode_batch(self, h, aad, num_samples, rate, channels, bits, max_block=1024, ms=False, sizes=None, out=None):
        """aad: uint8 [streams, stride].  Returns int16 [streams, channels, num_samples]."""
        aad = np.ascontiguousarray(aad, dtype=np.uint8)
        n_streams, stride = aad.shape
        prm = make_param(channels, rate, bits, max_block, ms, 0)
        b = self.batch(n_streams, num_samples, prm, aad_stream_stride=stride)
        pcm = out if out is not None else np.zeros((n_streams, channels, num_samples), dtype=np.int16)
        sz = None if sizes is None else np.ascontiguousarray(sizes, dtype=np.uint32)
        rc = self.lib.AADGpu_DecodeBatch(h, C.byref(b), aad.ctypes.data, None if sz is None else sz.ctypes.data,
                                         pcm.ctypes.data)
        if rc != OK:
            raise AADError("AADGpu_DecodeBatch", rc, self.last_error())
        return pcm
