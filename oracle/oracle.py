"""ctypes binding for the CPU ORACLE (oracle/libjpeg_oracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Product code under
jpeg_b200/ never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libjpeg_oracle.so")

INTERVAL_NONE = (1 << 63) - 1

ERRORS = {
    0: "ok", -1: "truncatedEntropyCodedSegment", -2: "invalidCompositeValue", -3: "invalidCompositeBlockRun",
    -4: "undefinedScanHuffmanDCReference", -5: "undefinedScanHuffmanACReference",
    -6: "undefinedScanQuantizationReference", -7: "precondition", -10: "lexing", -11: "parsing",
    -12: "decoding", -13: "unsupported",
}


class OracleError(RuntimeError):
    def __init__(self, code):
        super().__init__(f"oracle error {code} ({ERRORS.get(code, '?')})")
        self.code = code


class HuffSpec(C.Structure):
    _fields_ = [("present", C.c_int), ("counts", C.c_uint8 * 16), ("values", C.c_uint8 * 256)]

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
        n = sum(self.counts)
        return bytes(self.counts), bytes(self.values[:n])


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("jpeg_oracle.c", "jpeg_oracle.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    u8p, u16p, i16p, ip = C.POINTER(C.c_uint8), C.POINTER(C.c_uint16), C.POINTER(C.c_int16), C.POINTER(C.c_int)
    L.orc_zigzag.argtypes = [C.c_int, C.c_int]
    L.orc_extend.argtypes = [C.c_int, C.c_uint]
    L.orc_compact.argtypes = [C.c_int, ip, C.POINTER(C.c_uint)]
    L.orc_huff_lookup.argtypes = [C.POINTER(HuffSpec), C.c_uint, ip, ip]
    L.orc_huff_from_frequencies.argtypes = [C.POINTER(C.c_int64), C.POINTER(HuffSpec)]
    L.orc_huff_encoder.argtypes = [C.POINTER(HuffSpec), u16p, u8p]
    L.orc_quanta.argtypes = [C.c_double, C.c_int, u16p]
    L.orc_decompress.restype = C.c_void_p
    L.orc_decompress.argtypes = [C.c_char_p, C.c_size_t, ip]
    L.orc_decompress_format.restype = C.c_void_p
    L.orc_decompress_format.argtypes = [C.c_char_p, C.c_size_t, ip, C.c_int, C.c_int, ip]
    L.orc_decompress_scans.restype = C.c_void_p
    L.orc_decompress_scans.argtypes = [C.c_char_p, C.c_size_t, C.c_int, ip]
    L.orc_spectral_precision.argtypes = [C.c_void_p]
    L.orc_spectral_set_format.argtypes = [C.c_void_p, ip, C.c_int]
    L.orc_spectral_create.restype = C.c_void_p
    L.orc_spectral_create.argtypes = [C.c_int, C.c_int, C.c_int, ip, C.c_int]
    L.orc_spectral_free.argtypes = [C.c_void_p]
    L.orc_spectral_info.argtypes = [C.c_void_p, ip, ip, ip, ip, ip]
    L.orc_spectral_plane_info.argtypes = [C.c_void_p, C.c_int, ip, ip, ip]
    L.orc_spectral_coefficients.restype = i16p
    L.orc_spectral_coefficients.argtypes = [C.c_void_p, C.c_int]
    L.orc_spectral_quanta.restype = u16p
    L.orc_spectral_quanta.argtypes = [C.c_void_p, C.c_int]
    L.orc_spectral_set_quanta.argtypes = [C.c_void_p, C.c_int, u16p]
    L.orc_decode_scan.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, ip, ip, ip,
                                  C.POINTER(HuffSpec), C.POINTER(HuffSpec), C.c_void_p, C.POINTER(C.c_uint64),
                                  C.c_int, C.c_int64, C.c_int]
    L.orc_idct_plane.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    L.orc_interleave.argtypes = [C.POINTER(C.c_void_p), ip, ip, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.orc_unpack_rgb.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    L.orc_unpack_ycc.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    L.orc_pack_rgb.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    L.orc_decompose_plane.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_void_p]
    L.orc_fdct_plane.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    L.orc_encode_scan.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, ip, ip, ip, C.c_int64,
                                  C.POINTER(HuffSpec), C.POINTER(HuffSpec), C.POINTER(C.c_void_p),
                                  C.POINTER(C.c_size_t)]
    L.orc_free.argtypes = [C.c_void_p]
    _lib = L
    return L


def _ints(xs):
    return (C.c_int * len(xs))(*xs)


def units(size, stride):
    return size // stride + (1 if size % stride else 0)


def zigzag_table():
    L = lib()
    return np.array([[L.orc_zigzag(k, h) for k in range(8)] for h in range(8)], dtype=np.int32)


def quanta(level, chrominance):
    out = np.zeros(64, dtype=np.uint16)
    lib().orc_quanta(float(level), int(bool(chrominance)), out.ctypes.data_as(C.POINTER(C.c_uint16)))
    return out


def huff_from_frequencies(freq):
    f = (C.c_int64 * 256)(*[int(x) for x in freq])
    t = HuffSpec()
    lib().orc_huff_from_frequencies(f, C.byref(t))
    return t


def huff_encoder(spec):
    code = np.zeros(256, dtype=np.uint16)
    ln = np.zeros(256, dtype=np.uint8)
    lib().orc_huff_encoder(C.byref(spec), code.ctypes.data_as(C.POINTER(C.c_uint16)),
                           ln.ctypes.data_as(C.POINTER(C.c_uint8)))
    return code, ln


def huff_lookup(spec, codeword):
    s, l = C.c_int(), C.c_int()
    e = lib().orc_huff_lookup(C.byref(spec), codeword, C.byref(s), C.byref(l))
    if e:
        raise OracleError(e)
    return s.value, l.value


class Spectral:
    """Mirror of JPEG.Data.Spectral<JPEG.Common> held by the oracle."""

    def __init__(self, handle):
        self._h = handle
        L = lib()
        sz, bl, sc = _ints([0, 0]), _ints([0, 0]), _ints([0, 0])
        nc, pr = C.c_int(), C.c_int()
        L.orc_spectral_info(handle, sz, bl, C.byref(nc), sc, C.byref(pr))
        self.size = (sz[0], sz[1])
        self.blocks = (bl[0], bl[1])
        self.scale = (sc[0], sc[1])
        self.ncomp = nc.value
        self.process = pr.value

    @classmethod
    def decompress(cls, data: bytes, format=None):
        """format: None = JPEG.Common; (component keys in plane order, precision) = a user-defined JPEG.Format in the style of
        examples/custom-color/main.swift:41-63."""
        err = C.c_int()
        if format is None:
            h = lib().orc_decompress(data, len(data), C.byref(err))
        else:
            ids, precision = format
            h = lib().orc_decompress_format(data, len(data), _ints(list(ids)), len(ids), precision, C.byref(err))
        if not h:
            raise OracleError(err.value)
        return cls(h)

    @classmethod
    def decompress_scans(cls, data: bytes, scans: int):
        """the image as JPEG.Context holds it after `scans` scans (examples/decode-online/main.swift)"""
        err = C.c_int()
        h = lib().orc_decompress_scans(data, len(data), scans, C.byref(err))
        if not h:
            raise OracleError(err.value)
        return cls(h)

    @classmethod
    def create(cls, size, factors, progressive=False, format=None):
        flat = [v for f in factors for v in f]
        h = lib().orc_spectral_create(size[0], size[1], len(factors), _ints(flat), int(progressive))
        if format is not None:
            ids, precision = format
            assert len(ids) == len(factors)
            lib().orc_spectral_set_format(h, _ints(list(ids)), precision)
        return cls(h)

    @property
    def precision(self):
        return lib().orc_spectral_precision(self._h)

    def __del__(self):
        try:
            if self._h:
                lib().orc_spectral_free(self._h)
                self._h = None
        except Exception:
            pass

    def plane_info(self, p):
        u, f = _ints([0, 0]), _ints([0, 0])
        cid = C.c_int()
        lib().orc_spectral_plane_info(self._h, p, u, f, C.byref(cid))
        return (u[0], u[1]), (f[0], f[1]), cid.value

    def units(self, p):
        return self.plane_info(p)[0]

    def factor(self, p):
        return self.plane_info(p)[1]

    def coefficients(self, p):
        """numpy view (no copy) of the plane's Int16 buffer, shape (uy, ux, 64)."""
        (ux, uy), _, _ = self.plane_info(p)
        ptr = lib().orc_spectral_coefficients(self._h, p)
        n = 64 * ux * uy
        if n == 0:
            return np.zeros((uy, ux, 64), dtype=np.int16)
        arr = np.ctypeslib.as_array(ptr, shape=(n,))
        return arr.reshape(uy, ux, 64)

    def quanta(self, p):
        ptr = lib().orc_spectral_quanta(self._h, p)
        return np.ctypeslib.as_array(ptr, shape=(64,)).copy()

    def set_quanta(self, p, q):
        q = np.ascontiguousarray(q, dtype=np.uint16)
        lib().orc_spectral_set_quanta(self._h, p, q.ctypes.data_as(C.POINTER(C.c_uint16)))

    def decode_scan(self, band, bits, comps, dcsel, acsel, dc, ac, ecss, interval=INTERVAL_NONE, extend=False):
        """Spectral.decode(ecss:interval:scan:tables:extend:). ecss = list of unstuffed bytes."""
        cat = b"".join(ecss)
        offs = [0]
        for e in ecss:
            offs.append(offs[-1] + len(e))
        dcs = (HuffSpec * 4)(*dc)
        acs = (HuffSpec * 4)(*ac)
        buf = C.create_string_buffer(cat, len(cat) + 1)
        e = lib().orc_decode_scan(self._h, band[0], band[1], bits[0], -1 if bits[1] is None else bits[1],
                                  len(comps), _ints(comps), _ints(dcsel), _ints(acsel), dcs, acs,
                                  C.cast(buf, C.c_void_p), (C.c_uint64 * len(offs))(*offs), len(ecss),
                                  interval, int(extend))
        if e:
            raise OracleError(e)

    def encode_scan(self, band, bits, comps, dcsel, acsel, interval_mcus=0):
        """Returns (stuffed ECS bytes incl. RSTn, dc tables[4], ac tables[4])."""
        dcs, acs = (HuffSpec * 4)(), (HuffSpec * 4)()
        out, n = C.c_void_p(), C.c_size_t()
        e = lib().orc_encode_scan(self._h, band[0], band[1], bits[0], -1 if bits[1] is None else bits[1],
                                  len(comps), _ints(comps), _ints(dcsel), _ints(acsel), interval_mcus, dcs, acs,
                                  C.byref(out), C.byref(n))
        if e:
            raise OracleError(e)
        data = C.string_at(out, n.value)
        lib().orc_free(out)
        return data, list(dcs), list(acs)

    # --- transform stages -------------------------------------------------
    def idct(self):
        """Spectral.idct() -> list of uint16 planes, each (8uy, 8ux)."""
        planes = []
        for p in range(self.ncomp):
            (ux, uy), _, _ = self.plane_info(p)
            planes.append(idct_plane(self.coefficients(p), self.quanta(p), self.precision))
        return planes

    def to_rectangular(self, cosited=False):
        planes = self.idct()
        return interleave(planes, [self.units(p) for p in range(self.ncomp)],
                          [self.factor(p) for p in range(self.ncomp)], self.size, cosited)


def idct_plane(coef, q, precision=8):
    coef = np.ascontiguousarray(coef, dtype=np.int16)
    uy, ux = coef.shape[0], coef.shape[1]
    q = np.ascontiguousarray(q, dtype=np.uint16)
    out = np.zeros((8 * uy, 8 * ux), dtype=np.uint16)
    lib().orc_idct_plane(coef.ctypes.data, ux, uy, q.ctypes.data, precision, out.ctypes.data)
    return out


def interleave(planes, units_list, factors, size, cosited=False):
    n = len(planes)
    planes = [np.ascontiguousarray(p, dtype=np.uint16) for p in planes]
    ptrs = (C.c_void_p * n)(*[p.ctypes.data for p in planes])
    out = np.zeros((size[1], size[0], n), dtype=np.uint16)
    lib().orc_interleave(ptrs, _ints([v for u in units_list for v in u]), _ints([v for f in factors for v in f]),
                         n, size[0], size[1], int(cosited), out.ctypes.data)
    return out


def unpack_rgb(interleaved):
    il = np.ascontiguousarray(interleaved, dtype=np.uint16)
    h, w, n = il.shape
    out = np.zeros((h, w, 3), dtype=np.uint8)
    lib().orc_unpack_rgb(il.ctypes.data, h * w, n, out.ctypes.data)
    return out


def unpack_ycc(interleaved):
    il = np.ascontiguousarray(interleaved, dtype=np.uint16)
    h, w, n = il.shape
    out = np.zeros((h, w, 3), dtype=np.uint8)
    lib().orc_unpack_ycc(il.ctypes.data, h * w, n, out.ctypes.data)
    return out


def pack_rgb(rgb, ncomp=3):
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    h, w, _ = rgb.shape
    out = np.zeros((h, w, ncomp), dtype=np.uint16)
    lib().orc_pack_rgb(rgb.ctypes.data, h * w, ncomp, out.ctypes.data)
    return out


def decompose(interleaved, factors):
    il = np.ascontiguousarray(interleaved, dtype=np.uint16)
    h, w, n = il.shape
    scx = max(f[0] for f in factors)
    scy = max(f[1] for f in factors)
    planes = []
    for p, (fx, fy) in enumerate(factors):
        ux, uy = units(w * fx, 8 * scx), units(h * fy, 8 * scy)
        out = np.zeros((8 * uy, 8 * ux), dtype=np.uint16)
        lib().orc_decompose_plane(il.ctypes.data, w, h, n, p, fx, fy, scx, scy, out.ctypes.data)
        planes.append(out)
    return planes


def fdct_plane(samples, q, precision=8):
    s = np.ascontiguousarray(samples, dtype=np.uint16)
    uy, ux = s.shape[0] // 8, s.shape[1] // 8
    q = np.ascontiguousarray(q, dtype=np.uint16)
    out = np.zeros((uy, ux, 64), dtype=np.int16)
    lib().orc_fdct_plane(s.ctypes.data, ux, uy, q.ctypes.data, precision, out.ctypes.data)
    return out


def decode_rgb(data: bytes):
    """Rectangular<Common>.decompress + unpack(as: RGB) and unpack(as: YCbCr)."""
    s = Spectral.decompress(data)
    rect = s.to_rectangular()
    return unpack_rgb(rect), unpack_ycc(rect), s


# ----------------------------------------------------------------------------------------------------------------------
# N3: spectral-domain operations.  The reference has no library function for these; its examples spell them out as loops
# over Spectral.Plane subscripts, and the files those examples committed pin the arithmetic.
# ----------------------------------------------------------------------------------------------------------------------
def requantize(coef, q_old, q_new):
    """examples/recompress/main.swift:46-52: Int16(q_old[z]) * c, as Double / Double(q_new[z]), + 0.3 * sign, truncated."""
    c = coef.astype(np.int64) * np.asarray(q_old, dtype=np.int64)
    c = c.astype(np.int16).astype(np.float64)                      # `let coefficient:Int16 = ...` (no overflow in valid files)
    r = c / np.asarray(q_new, dtype=np.float64)
    return np.trunc(r + 0.3 * np.where(r < 0, -1.0, 1.0)).astype(np.int16)


def block_mapping(kind):
    """examples/rotate/main.swift:13-99 (Block.transform) and 113-152: per output zig-zag index z the source index and the
    sign, plus the block matrix ((xx, xy), (yx, yy)).  kind: 'ii' | 'iii' | 'iv' (quadrant the x axis is rotated into)."""
    zz = zigzag_table()                                            # zz[h][k] = z(k: k, h: h)
    blank = [(int(zz[y][x]), 1) for y in range(8) for x in range(8)]
    transpose = lambda a: [a[8 * x + y] for y in range(8) for x in range(8)]
    reflect_v = lambda a: [(a[8 * y + x][0], a[8 * y + x][1] * (1 - 2 * (y & 1))) for y in range(8) for x in range(8)]
    reflect_h = lambda a: [(a[8 * y + x][0], a[8 * y + x][1] * (1 - 2 * (x & 1))) for y in range(8) for x in range(8)]
    if kind == "ii":
        result, matrix = reflect_v(transpose(blank)), ((0, 1), (-1, 0))
    elif kind == "iii":
        result, matrix = reflect_v(reflect_h(blank)), ((-1, 0), (0, -1))
    elif kind == "iv":
        result, matrix = reflect_h(transpose(blank)), ((0, -1), (1, 0))
    else:
        raise ValueError(kind)
    zmap, mul = np.zeros(64, dtype=np.uint8), np.zeros(64, dtype=np.int8)
    for h in range(8):
        for k in range(8):
            zmap[zz[h][k]], mul[zz[h][k]] = result[8 * h + k]
    return zmap, mul, matrix


def transform_blocks(src, matrix, zmap, mul, dst_units):
    """examples/rotate/main.swift:164-190: block s of the source lands at offset + M s; coefficient z of the destination block is
    source coefficient zmap[z] times mul[z].  src: (uy, ux, 64); destination blocks outside dst_units are dropped
    (decode.swift:1470-1475), blocks nothing lands on stay zero."""
    uy, ux = src.shape[:2]
    (xx, xy), (yx, yy) = matrix
    ox = (ux - 1 if xx < 0 else 0) + (uy - 1 if xy < 0 else 0)
    oy = (ux - 1 if yx < 0 else 0) + (uy - 1 if yy < 0 else 0)
    dst = np.zeros((dst_units[1], dst_units[0], 64), dtype=np.int16)
    sy, sx = np.mgrid[0:uy, 0:ux]
    dx, dy = ox + xx * sx + xy * sy, oy + yx * sx + yy * sy
    ok = (dx >= 0) & (dx < dst_units[0]) & (dy >= 0) & (dy < dst_units[1])
    vals = (src[:, :, zmap].astype(np.int32) * mul.astype(np.int32)).astype(np.int16)
    dst[dy[ok], dx[ok]] = vals[ok]
    return dst


def rotated(s, kind):
    """examples/rotate/main.swift:101-199 on an oracle Spectral: crop to whole MCUs along the axes that get mirrored
    (Spectral.set(width:/height:), decode.swift:2456-2495), transform every plane, permute the quantisation tables."""
    zmap, mul, matrix = block_mapping(kind)
    w, h = s.size
    if kind in ("ii", "iii"):
        w -= w % (8 * s.scale[0])
    if kind in ("iii", "iv"):
        h -= h % (8 * s.scale[1])
    size = (w, h) if kind == "iii" else (h, w)
    factors = [s.factor(p) for p in range(s.ncomp)]
    out = Spectral.create(size, factors, progressive=bool(s.process))
    for p in range(s.ncomp):
        fx, fy = factors[p]
        cux, cuy = units(w * fx, 8 * s.scale[0]), units(h * fy, 8 * s.scale[1])   # plane units after the crop
        src = s.coefficients(p)[:cuy, :cux]
        out.coefficients(p)[...] = transform_blocks(src, matrix, zmap, mul, out.units(p))
        out.set_quanta(p, s.quanta(p)[zmap])
    return out
