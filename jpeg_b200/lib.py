"""ctypes binding of the C-ABI in include/jpeg_sm100.h (libjpeg_sm100.so).  No oracle, no fallback."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# JPEG_SM100_LIB: an alternative build of the same library (A/B measurements of kernel variants)
SO_PATH = os.environ.get("JPEG_SM100_LIB") or os.path.join(_HERE, "libjpeg_sm100.so")

INTERVAL_NONE = (1 << 64) - 1
BITS_MAX = -1
SCAN_EXTEND = 1  # bits of the `extend` argument of the decode_scan entry points
SCAN_FRESH = 2
SCAN_T81 = 4

OK = 0
ERR_TRUNCATED_ECS = -1
ERR_INVALID_COMPOSITE_VALUE = -2
ERR_INVALID_BLOCK_RUN = -3
ERR_UNDEFINED_DC = -4
ERR_UNDEFINED_AC = -5
ERR_PRECONDITION = -7
ERR_INVALID_HUFFMAN = -8
ERR_RESTART_PHASE = -9
ERR_ECS_COUNT = -10
ERR_MISSING_INTERVAL = -11
ERR_INVALID_ARGUMENT = -20
ERR_UNSUPPORTED = -21
ERR_NO_MEMORY = -22
ERR_CUDA = -100


class JpegSm100Error(RuntimeError):
    def __init__(self, code, detail=""):
        name = load().jpeg_sm100_error_string(code).decode()
        super().__init__(f"jpeg_sm100 error {code}: {name}{(' -- ' + detail) if detail else ''}")
        self.code = code


class HuffTable(C.Structure):
    _fields_ = [("present", C.c_int32), ("counts", C.c_uint8 * 16), ("values", C.c_uint8 * 256)]

    @classmethod
    def make(cls, counts, values):
        t = cls()
        t.present = 1
        for i, c in enumerate(counts):
            t.counts[i] = c
        for i, v in enumerate(values):
            t.values[i] = v
        return t

    def as_tuple(self):
        return bytes(self.counts), bytes(self.values[:sum(self.counts)])


class ScanComp(C.Structure):
    _fields_ = [("plane", C.c_int32), ("factor_x", C.c_int32), ("factor_y", C.c_int32), ("dc", C.c_int32),
                ("ac", C.c_int32)]


class ScanDesc(C.Structure):
    _fields_ = [("band_lo", C.c_int32), ("band_hi", C.c_int32), ("bit_lo", C.c_int32), ("bit_hi", C.c_int32),
                ("n_comp", C.c_int32), ("comp", ScanComp * 4), ("blocks_x", C.c_int32), ("blocks_y", C.c_int32)]


class PlaneI16(C.Structure):
    _fields_ = [("coef", C.c_void_p), ("units_x", C.c_int32), ("units_y", C.c_int32)]


class PlaneU16(C.Structure):
    _fields_ = [("samples", C.c_void_p), ("units_x", C.c_int32), ("units_y", C.c_int32), ("factor_x", C.c_int32),
                ("factor_y", C.c_int32)]


class DevSpectralPlane(C.Structure):
    _fields_ = [("coef", C.c_void_p), ("image_stride", C.c_uint64), ("units_x", C.c_int32), ("units_y", C.c_int32),
                ("factor_x", C.c_int32), ("factor_y", C.c_int32)]


class DevSpectral(C.Structure):
    _fields_ = [("n_images", C.c_uint32), ("n_planes", C.c_uint32), ("plane", DevSpectralPlane * 4)]


class DevPlanarPlane(C.Structure):
    _fields_ = [("samples", C.c_void_p), ("image_stride", C.c_uint64), ("units_x", C.c_int32), ("units_y", C.c_int32),
                ("factor_x", C.c_int32), ("factor_y", C.c_int32)]


class DevPlanar(C.Structure):
    _fields_ = [("n_images", C.c_uint32), ("n_planes", C.c_uint32), ("sample_bytes", C.c_int32),
                ("plane", DevPlanarPlane * 4)]


# every symbol include/jpeg_sm100.h declares: name -> (restype, argtypes)
_vp, _u64, _u32, _i = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
_HT = C.POINTER(HuffTable)
_SD = C.POINTER(ScanDesc)
SYMBOLS = {
    "jpeg_sm100_abi_version": (_i, []),
    "jpeg_sm100_create": (_i, [_i, C.POINTER(_vp)]),
    "jpeg_sm100_create_on_stream": (_i, [_i, _vp, C.POINTER(_vp)]),
    "jpeg_sm100_destroy": (None, [_vp]),
    "jpeg_sm100_sync": (_i, [_vp]),
    "jpeg_sm100_error_string": (C.c_char_p, [_i]),
    "jpeg_sm100_last_cuda_error": (C.c_char_p, [_vp]),
    "jpeg_sm100_launch_count": (_u64, [_vp]),
    "jpeg_sm100_malloc": (_i, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "jpeg_sm100_free": (_i, [_vp, _vp]),
    "jpeg_sm100_malloc_host": (_i, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "jpeg_sm100_free_host": (_i, [_vp, _vp]),
    "jpeg_sm100_upload": (_i, [_vp, _vp, _vp, C.c_size_t]),
    "jpeg_sm100_download": (_i, [_vp, _vp, _vp, C.c_size_t]),
    "jpeg_sm100_memset": (_i, [_vp, _vp, _i, C.c_size_t]),
    "jpeg_sm100_decode_scan": (_i, [_vp, _SD, _vp, _vp, _u32, _u64, _i, _HT, _HT, C.POINTER(PlaneI16), _u32]),
    "jpeg_sm100_idct": (_i, [_vp, _vp, _u32, _u32, _vp, _i, _vp]),
    "jpeg_sm100_idct_u8": (_i, [_vp, _vp, _u32, _u32, _vp, _vp]),
    "jpeg_sm100_interleave": (_i, [_vp, C.POINTER(PlaneU16), _u32, _u32, _u32, _i, _vp]),
    "jpeg_sm100_unpack_rgb8": (_i, [_vp, _vp, _u64, _i, _vp]),
    "jpeg_sm100_unpack_ycc8": (_i, [_vp, _vp, _u64, _i, _vp]),
    "jpeg_sm100_spectral_to_rgb8": (_i, [_vp, C.POINTER(PlaneI16), _u32, _vp, _vp, _u32, _u32, _i, _vp]),
    "jpeg_sm100_decode_batch_rgb8": (_i, [_vp, _SD, _u32, _vp, _vp, _u32, _u64, _HT, _i, _vp, _u32, _u32, _i, _vp, _vp]),
    "jpeg_sm100_lex_scan": (_i, [_vp, _vp, _u64, _vp, _u64, _vp, _u32, C.POINTER(_u32)]),
    "jpeg_sm100_decode_scan_raw": (_i, [_vp, _SD, _vp, _u64, _u64, _i, _HT, _HT, C.POINTER(PlaneI16), _u32]),
    "jpeg_sm100_decode_batch_raw_rgb8": (_i, [_vp, _SD, _u32, _vp, _vp, _vp, _u32, _u64, _HT, _i, _vp, _u32, _u32, _i, _vp, _vp]),
    "jpeg_sm100_dev_lex_scan": (_i, [_vp, _vp, _vp, _vp, _u32, _u32, _vp, _vp, _vp]),
    "jpeg_sm100_pack_rgb8": (_i, [_vp, _vp, _u64, _i, _vp]),
    "jpeg_sm100_decompose": (_i, [_vp, _vp, _u32, _u32, C.POINTER(PlaneU16), _u32]),
    "jpeg_sm100_fdct": (_i, [_vp, _vp, _u32, _u32, _vp, _i, _vp]),
    "jpeg_sm100_encode_scan": (_i, [_vp, _SD, C.POINTER(PlaneI16), _u32, _u64, _HT, _HT, _vp, _u64, C.POINTER(_u64)]),
    "jpeg_sm100_rgb8_to_spectral": (_i, [_vp, _vp, _u32, _u32, C.POINTER(PlaneI16), _u32, _vp, _vp]),
    "jpeg_sm100_dev_decode_scan": (_i, [_vp, _SD, _vp, _vp, _u32, _u64, _i, _HT, _i, C.POINTER(DevSpectral), _vp]),
    "jpeg_sm100_dev_idct": (_i, [_vp, C.POINTER(DevSpectral), _vp, _i, C.POINTER(DevPlanar)]),
    "jpeg_sm100_dev_planar_to_rgb8": (_i, [_vp, C.POINTER(DevPlanar), _u32, _u32, _i, _vp]),
    "jpeg_sm100_dev_spectral_to_rgb8": (_i, [_vp, C.POINTER(DevSpectral), _vp, _u32, _u32, _i, _vp]),
    "jpeg_sm100_dev_interleave": (_i, [_vp, C.POINTER(DevPlanar), _u32, _u32, _i, _vp]),
    "jpeg_sm100_dev_unpack_rgb8": (_i, [_vp, _vp, _u64, _i, _vp]),
    "jpeg_sm100_dev_unpack_ycc8": (_i, [_vp, _vp, _u64, _i, _vp]),
    "jpeg_sm100_dev_rgb8_to_planar": (_i, [_vp, _vp, _u32, _u32, C.POINTER(DevPlanar)]),
    "jpeg_sm100_dev_pack_rgb8": (_i, [_vp, _vp, _u64, _i, _vp]),
    "jpeg_sm100_dev_decompose": (_i, [_vp, _vp, _u32, _u32, C.POINTER(DevPlanar)]),
    "jpeg_sm100_dev_fdct": (_i, [_vp, C.POINTER(DevPlanar), _vp, _i, C.POINTER(DevSpectral)]),
    "jpeg_sm100_dev_encode_scan": (_i, [_vp, _SD, C.POINTER(DevSpectral), _u64, _HT, _vp, _u64, _vp]),
    "jpeg_sm100_requantize": (_i, [_vp, _vp, _u32, _u32, _vp, _vp, _vp]),
    "jpeg_sm100_transform_blocks": (_i, [_vp, _vp, _u32, _u32, _vp, _vp, _vp, _vp, _u32, _u32]),
    "jpeg_sm100_dev_requantize": (_i, [_vp, C.POINTER(DevSpectral), _vp, _vp, C.POINTER(DevSpectral)]),
    "jpeg_sm100_dev_transform_blocks": (_i, [_vp, C.POINTER(DevSpectral), _vp, _vp, _vp, C.POINTER(DevSpectral)]),
    "jpeg_sm100_spectral_create": (_i, [_vp, _u32, _vp, C.POINTER(_vp)]),
    "jpeg_sm100_spectral_destroy": (None, [_vp, _vp]),
    "jpeg_sm100_spectral_resize": (_i, [_vp, _vp, _vp]),
    "jpeg_sm100_spectral_upload": (_i, [_vp, _vp, C.POINTER(PlaneI16), _u32]),
    "jpeg_sm100_spectral_download": (_i, [_vp, _vp, C.POINTER(PlaneI16), _u32]),
    "jpeg_sm100_spectral_decode_scan": (_i, [_vp, _vp, _SD, _vp, _vp, _u32, _u64, _i, _HT, _HT]),
    "jpeg_sm100_spectral_decode_scan_raw": (_i, [_vp, _vp, _SD, _vp, _u64, _u64, _i, _HT, _HT]),
    "jpeg_sm100_spectral_encode_scan": (_i, [_vp, _vp, _SD, _u64, _HT, _HT, _vp, _u64, C.POINTER(_u64)]),
    "jpeg_sm100_spectral_idct": (_i, [_vp, _vp, _vp, _i, C.POINTER(PlaneU16), _u32]),
    "jpeg_sm100_spectral_rgb8": (_i, [_vp, _vp, _vp, _vp, _u32, _u32, _i, _vp]),
    "jpeg_sm100_transfer_counts": (None, [_vp, C.POINTER(_u64), C.POINTER(_u64)]),
}

_lib = None


def load():
    """Loads libjpeg_sm100.so; raises loudly if it has not been built (there is no fallback implementation)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). jpeg_b200 has no CPU fallback.")
    L = C.CDLL(SO_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


class Context:
    """One device + one stream (jpeg_sm100_ctx)."""

    def __init__(self, device=0, stream=None):
        L = load()
        h = C.c_void_p()
        if stream is None:
            rc = L.jpeg_sm100_create(device, C.byref(h))
        else:
            rc = L.jpeg_sm100_create_on_stream(device, C.c_void_p(stream), C.byref(h))
        if rc != 0:
            raise JpegSm100Error(rc, "jpeg_sm100_create failed: is an sm_100 (B200) device visible?")
        self.h = h
        self.L = L

    def close(self):
        if getattr(self, "h", None):
            self.L.jpeg_sm100_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != 0:
            detail = self.L.jpeg_sm100_last_cuda_error(self.h).decode() if rc == ERR_CUDA else ""
            raise JpegSm100Error(rc, detail)

    def sync(self):
        self.check(self.L.jpeg_sm100_sync(self.h))

    @property
    def launches(self):
        return int(self.L.jpeg_sm100_launch_count(self.h))

    @property
    def transfers(self):
        """(bytes host -> device, bytes device -> host) moved by the layer-A calls of this context so far"""
        a, b = C.c_uint64(), C.c_uint64()
        self.L.jpeg_sm100_transfer_counts(self.h, C.byref(a), C.byref(b))
        return int(a.value), int(b.value)
