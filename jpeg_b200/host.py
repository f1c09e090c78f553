"""Host-side mirror of the reference's staged interface (tayloraswift/jpeg), with the hot-path bodies replaced by
calls through the C-ABI of libjpeg_sm100.so -- exactly the substitution swift/JPEGSM100Shim.swift makes inside the
Swift module.

    Spectral.decompress(bytes)            JPEG.Data.Spectral.decompress(stream:)      decode.swift:4315, 3728
      .idct()              -> Planar      Spectral.idct()                             decode.swift:4154
      .interleaved(cosite) -> Rectangular Planar.interleaved(cosite:)                 decode.swift:4182
      .unpack_rgb()/.unpack_ycc()         Rectangular.unpack(as:)                     decode.swift:4294
    Rectangular.pack(rgb, ...)            Rectangular.pack(size:layout:metadata:pixels:)  encode.swift:456
      .decomposed()        -> Planar      Rectangular.decomposed()                    encode.swift:389
      .fdct(quanta)        -> Spectral    Planar.fdct(quanta:)                        encode.swift:353
      .compress()          -> bytes       Spectral.compress(stream:)                  encode.swift:1918

Container lexing / parsing / serialisation (decode.swift:53-1005, 3554-3961; encode.swift:1623-1972) is host
bookkeeping that the reference keeps in Swift; it is restated here only so the tests can drive whole files through
the GPU path.  No oracle, no CPU fallback: every stage below calls the CUDA library.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import lib as L


class LexingError(ValueError):
    pass


class ParsingError(ValueError):
    pass


class DecodingError(ValueError):
    pass


def units(size, stride):
    return size // stride + (1 if size % stride else 0)


_ctx = None


def transfer_stats(s):
    """host <-> device traffic of a device-resident Spectral's context since the Spectral was created (bytes)"""
    a, b = s.ctx.transfers
    return {"h2d_bytes": a - s._transfers0[0], "d2h_bytes": b - s._transfers0[1]}


def default_context():
    global _ctx
    if _ctx is None:
        _ctx = L.Context(0)
    return _ctx


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


def _zigzag(k, h):
    """JPEG.Table.Quantization.z(k:h:) decode.swift:1289-1298"""
    p = 1 if k + h < 8 else 0
    q = (k + h) & 1
    a, b = 72 * (p ^ 1), 2 * p - 1
    n = b * (k + h) - 14 * p + 15
    return a + b * ((n * (n + 1)) >> 1) - q * k - (q ^ 1) * h - 1


# ---------------------------------------------------------------------------------------------------------------
# data types
# ---------------------------------------------------------------------------------------------------------------
class CompressionLevel:
    """JPEG.CompressionLevel (encode.swift:260-333): quantum values for a quality parameter, 0.0 = all ones, 1.0 = the
    keyframe table; `.quanta` is in zig-zag order.  Host-side arithmetic (Double), kept next to the types that consume it."""
    _KEYFRAMES = {
        "luminance": (16, 11, 10, 16, 124, 140, 151, 161, 12, 12, 14, 19, 126, 158, 160, 155,
                      14, 13, 16, 24, 140, 157, 169, 156, 14, 17, 22, 29, 151, 187, 180, 162,
                      18, 22, 37, 56, 168, 109, 103, 177, 24, 35, 55, 64, 181, 104, 113, 192,
                      49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 199),
        "chrominance": (17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
                        24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99) + (99,) * 32,
    }

    def __init__(self, kind, level):
        if kind not in self._KEYFRAMES:
            raise ValueError(kind)
        self.kind, self.level = kind, float(level)

    @classmethod
    def luminance(cls, level):
        return cls("luminance", level)

    @classmethod
    def chrominance(cls, level):
        return cls("chrominance", level)

    @property
    def quanta(self):
        t = self.level
        key = np.asarray(self._KEYFRAMES[self.kind], dtype=np.float64)
        v = 1.0 * (1 - t) + key * t
        v = np.copysign(np.floor(np.abs(v) + 0.5), v)  # Double.rounded(): to nearest, ties away from zero
        v = np.maximum(1.0, np.minimum(v, 255.0)).astype(np.uint16)
        out = np.zeros(64, dtype=np.uint16)
        for h in range(8):
            for k in range(8):
                out[_zigzag(k, h)] = v[8 * h + k]
        return out


@dataclass(frozen=True)
class Format:
    """A user-defined JPEG.Format (jpeg.swift:300-340) in the style of examples/custom-color/main.swift:41-63: recognised iff
    the frame's component keys are exactly `components` and its precision is `precision`; `components` orders the planes.
    `None` wherever a format is expected means JPEG.Common (jpeg.swift:370-397)."""
    components: tuple
    precision: int

    def recognize(self, keys, precision):
        return sorted(keys) == sorted(self.components) and precision == self.precision


class SpectralPlane:
    """JPEG.Data.Spectral.Plane: units, sampling factor, coefficients int16 (uy, ux, 64), index into Spectral.quanta.
    In a device-resident Spectral the coefficients live in HBM; `coef` materialises the host copy on first use."""

    def __init__(self, units, factor, coef, q=0, comp_id=0, owner=None):
        self.units, self.factor, self._coef, self.q, self.comp_id, self._owner = units, factor, coef, q, comp_id, owner

    @property
    def coef(self):
        if self._owner is not None:
            self._owner._materialize()
        return self._coef

    @coef.setter
    def coef(self, value):
        if self._owner is not None and self._owner.resident:
            raise ValueError("the coefficients of a device-resident Spectral are read-only on the host")
        self._coef = value


@dataclass
class Scan:
    band: tuple
    bits: tuple            # (lo, hi) with hi None == .max
    comps: list            # [(plane index, dc slot, ac slot)]


class Spectral:
    """JPEG.Data.Spectral<JPEG.Common> (decode.swift:1397-1519)."""

    def __init__(self, size, factors, comp_ids=None, process=0, ctx=None, precision=8, resident=False):
        """resident: keep the coefficient planes in HBM (a jpeg_sm100_spectral handle) across scans and stages -- what
        JPEG.Context does with its one Spectral per file (decode.swift:3565-3587); host planes are materialised on demand."""
        self.ctx = ctx or default_context()
        self.process = process
        self.precision = precision  # Format.precision: 8 for JPEG.Common, up to 16 for user-defined formats
        self.scale = (max(f[0] for f in factors), max(f[1] for f in factors))
        self.quanta = [np.zeros(64, dtype=np.uint16)]
        self.resident, self._handle, self._host_valid = bool(resident), None, True
        self.planes = [SpectralPlane((0, 0), tuple(f), np.zeros((0, 0, 64), np.int16), 0,
                                     (comp_ids[i] if comp_ids else i + 1), self if resident else None) for i, f in enumerate(factors)]
        self.size = (0, 0)
        self.blocks = (0, 0)
        self.scans: list[Scan] = []
        self._transfers0 = self.ctx.transfers if resident else (0, 0)
        self.set_size(size)

    def __del__(self):
        try:
            if self._handle is not None:
                self.ctx.L.jpeg_sm100_spectral_destroy(self.ctx.h, self._handle)
                self._handle = None
        except Exception:
            pass

    def _units_array(self):
        return np.array([v for p in self.planes for v in p.units], dtype=np.int32)

    def _materialize(self):
        """device-resident: bring the coefficient planes to the host (once per change on the device)"""
        if not self.resident or self._host_valid:
            return
        for p in self.planes:
            p._coef = np.zeros((p.units[1], p.units[0], 64), dtype=np.int16)
        arr = (L.PlaneI16 * len(self.planes))()
        for i, p in enumerate(self.planes):
            arr[i].coef = p._coef.ctypes.data if p._coef.size else None
            arr[i].units_x, arr[i].units_y = p.units
        self.ctx.check(self.ctx.L.jpeg_sm100_spectral_download(self.ctx.h, self._handle, arr, len(self.planes)))
        for p in self.planes:
            p._coef.setflags(write=False)
        self._host_valid = True

    # decode.swift:2456-2495 set(width:) / set(height:)
    def set_size(self, size):
        w, h = size
        self.size = (w, h)
        self.blocks = (units(w, 8 * self.scale[0]), units(h, 8 * self.scale[1]))
        if self.resident:
            for p in self.planes:
                p.units = (units(w * p.factor[0], 8 * self.scale[0]), units(h * p.factor[1], 8 * self.scale[1]))
            u = self._units_array()
            if self._handle is None:
                h_ = C.c_void_p()
                self.ctx.check(self.ctx.L.jpeg_sm100_spectral_create(self.ctx.h, len(self.planes), _ptr(u), C.byref(h_)))
                self._handle = h_
            else:
                self.ctx.check(self.ctx.L.jpeg_sm100_spectral_resize(self.ctx.h, self._handle, _ptr(u)))
            self._host_valid = False
            return
        for p in self.planes:
            ux, uy = units(w * p.factor[0], 8 * self.scale[0]), units(h * p.factor[1], 8 * self.scale[1])
            new = np.zeros((uy, ux, 64), dtype=np.int16)
            oy, ox = min(uy, p.coef.shape[0]), min(ux, p.coef.shape[1])
            new[:oy, :ox] = p.coef[:oy, :ox]
            p.coef, p.units = new, (ux, uy)

    @property
    def ncomp(self):
        return len(self.planes)

    # ---- hot path: one scan -------------------------------------------------------------------------------
    def scan_desc(self, band, bits, comps):
        d = L.ScanDesc()
        d.band_lo, d.band_hi = band
        d.bit_lo = bits[0]
        d.bit_hi = L.BITS_MAX if bits[1] is None else bits[1]
        d.n_comp = len(comps)
        for i, (p, dc, ac) in enumerate(comps):
            d.comp[i].plane = p
            d.comp[i].factor_x, d.comp[i].factor_y = self.planes[p].factor
            d.comp[i].dc, d.comp[i].ac = dc, ac
        d.blocks_x, d.blocks_y = self.blocks
        return d

    def _plane_structs(self):
        arr = (L.PlaneI16 * len(self.planes))()
        for i, p in enumerate(self.planes):
            p.coef = np.ascontiguousarray(p.coef)
            arr[i].coef = p.coef.ctypes.data if p.coef.size else None
            arr[i].units_x, arr[i].units_y = p.units
        return arr

    def decode_scan(self, band, bits, comps, dc_tables, ac_tables, ecss, interval, extend=False):
        """Spectral.decode(ecss:interval:scan:tables:extend:) decode.swift:3476 -> jpeg_sm100_decode_scan."""
        cat = b"".join(ecss)
        offs = np.zeros(len(ecss) + 1, dtype=np.uint64)
        np.cumsum([len(e) for e in ecss], out=offs[1:])
        buf = np.frombuffer(cat + b"\0" * 8, dtype=np.uint8)
        dcs = (L.HuffTable * 4)(*[t if t is not None else L.HuffTable() for t in dc_tables])
        acs = (L.HuffTable * 4)(*[t if t is not None else L.HuffTable() for t in ac_tables])
        desc = self.scan_desc(band, bits, comps)
        if self.resident:
            rc = self.ctx.L.jpeg_sm100_spectral_decode_scan(self.ctx.h, self._handle, C.byref(desc), _ptr(buf), _ptr(offs), len(ecss),
                                                            L.INTERVAL_NONE if interval is None else interval, int(extend), dcs, acs)
            self._host_valid = False
            self.ctx.check(rc)
            return
        planes = self._plane_structs()
        rc = self.ctx.L.jpeg_sm100_decode_scan(self.ctx.h, C.byref(desc), _ptr(buf), _ptr(offs), len(ecss),
                                               L.INTERVAL_NONE if interval is None else interval, int(extend),
                                               dcs, acs, planes, len(self.planes))
        self.ctx.check(rc)

    def decode_scan_raw(self, band, bits, comps, dc_tables, ac_tables, raw, interval, extend=False):
        """The same scan from its raw bytes (stuffed, RSTn-delimited): lexing (decode.swift:130-190, 3895-3933) and
        decoding both on the GPU -> jpeg_sm100_decode_scan_raw."""
        buf = np.frombuffer(bytes(raw) + b"\0" * 32, dtype=np.uint8)
        dcs = (L.HuffTable * 4)(*[t if t is not None else L.HuffTable() for t in dc_tables])
        acs = (L.HuffTable * 4)(*[t if t is not None else L.HuffTable() for t in ac_tables])
        desc = self.scan_desc(band, bits, comps)
        if self.resident:
            rc = self.ctx.L.jpeg_sm100_spectral_decode_scan_raw(self.ctx.h, self._handle, C.byref(desc), _ptr(buf), len(raw),
                                                                L.INTERVAL_NONE if interval is None else interval, int(extend), dcs, acs)
            self._host_valid = False
            self.ctx.check(rc)
            return
        planes = self._plane_structs()
        rc = self.ctx.L.jpeg_sm100_decode_scan_raw(self.ctx.h, C.byref(desc), _ptr(buf), len(raw),
                                                   L.INTERVAL_NONE if interval is None else interval, int(extend),
                                                   dcs, acs, planes, len(self.planes))
        self.ctx.check(rc)

    def encode_scan(self, band, bits, comps, interval_mcus=0):
        """Spectral.encode(scan:) encode.swift:1559 -> jpeg_sm100_encode_scan.  Returns (ecs, dc[4], ac[4])."""
        desc = self.scan_desc(band, bits, comps)
        dcs, acs = (L.HuffTable * 4)(), (L.HuffTable * 4)()
        cap = sum(64 * p.units[0] * p.units[1] for p in self.planes) * 4 + 4096
        out = np.zeros(cap, dtype=np.uint8)
        n = C.c_uint64()
        if self.resident:
            rc = self.ctx.L.jpeg_sm100_spectral_encode_scan(self.ctx.h, self._handle, C.byref(desc), interval_mcus, dcs, acs,
                                                            _ptr(out), cap, C.byref(n))
        else:
            planes = self._plane_structs()
            rc = self.ctx.L.jpeg_sm100_encode_scan(self.ctx.h, C.byref(desc), planes, len(self.planes), interval_mcus,
                                                   dcs, acs, _ptr(out), cap, C.byref(n))
        self.ctx.check(rc)
        return out[:n.value].tobytes(), list(dcs), list(acs)

    # ---- transform stages ----------------------------------------------------------------------------------
    def idct(self):
        """Spectral.idct() decode.swift:4154 -> jpeg_sm100_idct per plane (one call on the resident image)."""
        out = []
        if self.resident:
            arr = (L.PlaneU16 * len(self.planes))()
            for i, p in enumerate(self.planes):
                out.append(np.zeros((8 * p.units[1], 8 * p.units[0]), dtype=np.uint16))
                arr[i].samples = out[i].ctypes.data if out[i].size else None
                arr[i].units_x, arr[i].units_y = p.units
                arr[i].factor_x, arr[i].factor_y = p.factor
            q = np.ascontiguousarray(np.stack([self.quanta[p.q] for p in self.planes]), dtype=np.uint16)
            self.ctx.check(self.ctx.L.jpeg_sm100_spectral_idct(self.ctx.h, self._handle, _ptr(q), self.precision, arr, len(self.planes)))
            return Planar(self.size, [p.units for p in self.planes], [p.factor for p in self.planes], out, self.ctx, self.precision)
        for p in self.planes:
            ux, uy = p.units
            s = np.zeros((8 * uy, 8 * ux), dtype=np.uint16)
            coef = np.ascontiguousarray(p.coef)
            q = np.ascontiguousarray(self.quanta[p.q], dtype=np.uint16)
            self.ctx.check(self.ctx.L.jpeg_sm100_idct(self.ctx.h, _ptr(coef), ux, uy, _ptr(q), self.precision, _ptr(s)))
            out.append(s)
        return Planar(self.size, [p.units for p in self.planes], [p.factor for p in self.planes], out, self.ctx,
                      self.precision)

    def idct_u8(self):
        if self.precision != 8:
            raise ValueError("8-bit sample planes exist for 8-bit formats only")
        out = []
        for p in self.planes:
            ux, uy = p.units
            s = np.zeros((8 * uy, 8 * ux), dtype=np.uint8)
            coef = np.ascontiguousarray(p.coef)
            q = np.ascontiguousarray(self.quanta[p.q], dtype=np.uint16)
            self.ctx.check(self.ctx.L.jpeg_sm100_idct_u8(self.ctx.h, _ptr(coef), ux, uy, _ptr(q), _ptr(s)))
            out.append(s)
        return out

    def to_rgb8(self, cosited=False):
        """Fused Spectral -> RGB8 (jpeg_sm100_spectral_to_rgb8; jpeg_sm100_spectral_rgb8 on the resident image)."""
        q = np.ascontiguousarray(np.stack([self.quanta[p.q] for p in self.planes]), dtype=np.uint16)
        f = np.array([v for p in self.planes for v in p.factor], dtype=np.int32)
        rgb = np.zeros((self.size[1], self.size[0], 3), dtype=np.uint8)
        if self.resident:
            self.ctx.check(self.ctx.L.jpeg_sm100_spectral_rgb8(self.ctx.h, self._handle, _ptr(q), _ptr(f), self.size[0], self.size[1],
                                                               int(cosited), _ptr(rgb)))
            return rgb
        planes = self._plane_structs()
        self.ctx.check(self.ctx.L.jpeg_sm100_spectral_to_rgb8(self.ctx.h, planes, len(self.planes), _ptr(q), _ptr(f),
                                                              self.size[0], self.size[1], int(cosited), _ptr(rgb)))
        return rgb

    # ---- N3: spectral-domain operations (the loops of the reference's examples, as kernels) -------------------------
    def requantized(self, new_quanta):
        """examples/recompress/main.swift:35-58: the same image with other quantisation tables, without leaving the
        coefficient domain.  new_quanta: one 64-entry table (zig-zag) per entry of self.quanta."""
        out = Spectral(self.size, [p.factor for p in self.planes], [p.comp_id for p in self.planes], self.process, self.ctx,
                       self.precision)
        out.quanta = [np.ascontiguousarray(q, dtype=np.uint16) for q in new_quanta]
        for src, dst in zip(self.planes, out.planes):
            dst.q = src.q
            coef = np.ascontiguousarray(src.coef)
            res = np.zeros_like(coef)
            qo = np.ascontiguousarray(self.quanta[src.q], dtype=np.uint16)
            self.ctx.check(self.ctx.L.jpeg_sm100_requantize(self.ctx.h, _ptr(coef), src.units[0], src.units[1], _ptr(qo),
                                                            _ptr(out.quanta[src.q]), _ptr(res)))
            dst.coef = res
        return out

    @staticmethod
    def block_mapping(kind):
        """examples/rotate/main.swift:13-99, 113-152: (zmap, mul, matrix) of the rotation into quadrant 'ii' | 'iii' | 'iv'."""
        zz = [[_zigzag(k, h) for k in range(8)] for h in range(8)]
        blank = [(zz[y][x], 1) for y in range(8) for x in range(8)]
        transpose = lambda a: [a[8 * x + y] for y in range(8) for x in range(8)]
        reflect_v = lambda a: [(a[8 * y + x][0], a[8 * y + x][1] * (1 - 2 * (y & 1))) for y in range(8) for x in range(8)]
        reflect_h = lambda a: [(a[8 * y + x][0], a[8 * y + x][1] * (1 - 2 * (x & 1))) for y in range(8) for x in range(8)]
        result, matrix = {"ii": (reflect_v(transpose(blank)), (0, 1, -1, 0)),
                          "iii": (reflect_v(reflect_h(blank)), (-1, 0, 0, -1)),
                          "iv": (reflect_h(transpose(blank)), (0, -1, 1, 0))}[kind]
        zmap, mul = np.zeros(64, dtype=np.uint8), np.zeros(64, dtype=np.int8)
        for h in range(8):
            for k in range(8):
                zmap[zz[h][k]], mul[zz[h][k]] = result[8 * h + k]
        return zmap, mul, np.array(matrix, dtype=np.int32)

    def rotated(self, kind):
        """examples/rotate/main.swift:101-199: lossless rotation; the axes that get mirrored are first cropped to whole MCUs."""
        zmap, mul, matrix = self.block_mapping(kind)
        w, h = self.size
        if kind in ("ii", "iii"):
            w -= w % (8 * self.scale[0])
        if kind in ("iii", "iv"):
            h -= h % (8 * self.scale[1])
        src = Spectral(self.size, [p.factor for p in self.planes], [p.comp_id for p in self.planes], self.process, self.ctx,
                       self.precision)
        for a, b in zip(self.planes, src.planes):
            b.coef, b.q = a.coef, a.q
        src.set_size((w, h))                                       # Spectral.set(width:) / set(height:)
        out = Spectral((w, h) if kind == "iii" else (h, w), [p.factor for p in self.planes],
                       [p.comp_id for p in self.planes], self.process, self.ctx, self.precision)
        out.quanta = [np.ascontiguousarray(q, dtype=np.uint16)[zmap] for q in self.quanta]
        for a, b in zip(src.planes, out.planes):
            b.q = a.q
            coef = np.ascontiguousarray(a.coef)
            res = np.zeros((b.units[1], b.units[0], 64), dtype=np.int16)
            self.ctx.check(self.ctx.L.jpeg_sm100_transform_blocks(self.ctx.h, _ptr(coef), a.units[0], a.units[1], _ptr(matrix),
                                                                  _ptr(zmap), _ptr(mul), _ptr(res), b.units[0], b.units[1]))
            b.coef = res
        return out

    # ---- container: JPEG.Context.decompress (decode.swift:3728-3960) ----------------------------------------
    @classmethod
    def decompress(cls, data: bytes, ctx=None, gpu_lexer=False, format=None, on_scan=None, resident=False, t81_intervals=False):
        """on_scan(spectral, scan): called after every scan with the image as JPEG.Context holds it at that point -- the
        capture closure of examples/decode-online/main.swift:252-282 (online / progressive display).
        resident: the image stays in HBM across its scans and stages (see __init__).
        t81_intervals: restart intervals are placed as ITU-T T.81 defines them (interval e starts at MCU e * Ri; JPEG_SM100_SCAN_T81)
        instead of the reference's placement by rows (decode.swift:3205-3207), which mis-decodes files whose DRI is not a whole
        number of MCU rows.  An extension; the default is the reference's behaviour."""
        return _decompress(data, ctx or default_context(), gpu_lexer, format, on_scan, resident, t81_intervals)

    def compress(self, scans=None, quanta_slots=None, interval_mcus=0, jfif=True):
        return _compress(self, scans, quanta_slots, interval_mcus, jfif)


class Planar:
    """JPEG.Data.Planar<JPEG.Common> (decode.swift:1543-1632); planes are uint16 (8uy, 8ux)."""

    def __init__(self, size, units_list, factors, planes, ctx=None, precision=8):
        self.ctx = ctx or default_context()
        self.size, self.units, self.factors, self.planes = size, list(units_list), list(factors), planes
        self.precision = precision

    def _structs(self):
        arr = (L.PlaneU16 * len(self.planes))()
        for i, pl in enumerate(self.planes):
            self.planes[i] = np.ascontiguousarray(pl, dtype=np.uint16)
            arr[i].samples = self.planes[i].ctypes.data if self.planes[i].size else None
            arr[i].units_x, arr[i].units_y = self.units[i]
            arr[i].factor_x, arr[i].factor_y = self.factors[i]
        return arr

    def interleaved(self, cosite=False):
        """Planar.interleaved(cosite:) decode.swift:4182 -> jpeg_sm100_interleave."""
        n = len(self.planes)
        out = np.zeros((self.size[1], self.size[0], n), dtype=np.uint16)
        self.ctx.check(self.ctx.L.jpeg_sm100_interleave(self.ctx.h, self._structs(), n, self.size[0], self.size[1],
                                                        int(cosite), _ptr(out)))
        return Rectangular(self.size, self.factors, out, self.ctx, self.precision)

    def fdct(self, quanta, comp_ids=None, process=0):
        """Planar.fdct(quanta:) encode.swift:353 -> jpeg_sm100_fdct per plane.  quanta: one 64-array per plane."""
        sp = Spectral(self.size, self.factors, comp_ids, process, ctx=self.ctx, precision=self.precision)
        for i, pl in enumerate(self.planes):
            ux, uy = self.units[i]
            pl = np.ascontiguousarray(pl, dtype=np.uint16)
            q = np.ascontiguousarray(quanta[i], dtype=np.uint16)
            coef = np.zeros((uy, ux, 64), dtype=np.int16)
            self.ctx.check(self.ctx.L.jpeg_sm100_fdct(self.ctx.h, _ptr(pl), ux, uy, _ptr(q), self.precision, _ptr(coef)))
            sp.planes[i].coef = coef
            sp.quanta.append(q.copy())
            sp.planes[i].q = len(sp.quanta) - 1
        return sp


class Rectangular:
    """JPEG.Data.Rectangular<JPEG.Common> (decode.swift:1650-1718); values uint16 (h, w, n)."""

    def __init__(self, size, factors, values, ctx=None, precision=8):
        self.ctx = ctx or default_context()
        self.size, self.factors, self.values = size, list(factors), values
        self.precision = precision

    def _unpack(self, fn):
        v = np.ascontiguousarray(self.values, dtype=np.uint16)
        h, w, n = v.shape
        out = np.zeros((h, w, 3), dtype=np.uint8)
        self.ctx.check(fn(self.ctx.h, _ptr(v), h * w, n, _ptr(out)))
        return out

    def unpack_rgb(self):
        return self._unpack(self.ctx.L.jpeg_sm100_unpack_rgb8)

    def unpack_ycc(self):
        return self._unpack(self.ctx.L.jpeg_sm100_unpack_ycc8)

    @classmethod
    def pack(cls, rgb, factors, ctx=None):
        """Rectangular.pack(size:layout:metadata:pixels:) with RGB pixels -> jpeg_sm100_pack_rgb8."""
        ctx = ctx or default_context()
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        h, w, _ = rgb.shape
        n = len(factors)
        out = np.zeros((h, w, n), dtype=np.uint16)
        ctx.check(ctx.L.jpeg_sm100_pack_rgb8(ctx.h, _ptr(rgb), h * w, n, _ptr(out)))
        return cls((w, h), factors, out, ctx)

    def decomposed(self):
        """Rectangular.decomposed() encode.swift:389 -> jpeg_sm100_decompose."""
        w, h = self.size
        scx, scy = max(f[0] for f in self.factors), max(f[1] for f in self.factors)
        us = [(units(w * fx, 8 * scx), units(h * fy, 8 * scy)) for fx, fy in self.factors]
        planes = [np.zeros((8 * uy, 8 * ux), dtype=np.uint16) for ux, uy in us]
        pl = Planar(self.size, us, self.factors, planes, self.ctx, self.precision)
        v = np.ascontiguousarray(self.values, dtype=np.uint16)
        self.ctx.check(self.ctx.L.jpeg_sm100_decompose(self.ctx.h, _ptr(v), w, h, pl._structs(), len(planes)))
        return pl


# ---------------------------------------------------------------------------------------------------------------
# container (host bookkeeping; restated from decode.swift:130-190, 475-1005, 3728-3960)
# ---------------------------------------------------------------------------------------------------------------
def _marker_valid(c):
    return (0xC0 <= c <= 0xCF and c != 0xC8) or 0xD0 <= c <= 0xEF or c == 0xFE


class _Lexer:
    def __init__(self, data):
        self.d, self.pos, self.n = data, 0, len(data)

    def segment(self, prefix=False):
        """decode.swift:130-190 -> (ecs, marker, body)."""
        d, n = self.d, self.n
        ecs = bytearray()
        while self.pos < n:
            k = d.find(b"\xff", self.pos)
            if k < 0:
                break
            if k > self.pos:
                if not prefix:
                    raise LexingError("invalidMarkerSegmentPrefix")
                ecs += d[self.pos:k]
            self.pos = k + 1
            stuffed = False
            while True:
                if self.pos >= n:
                    raise LexingError("truncatedMarkerSegmentType")
                b = d[self.pos]
                self.pos += 1
                if b == 0x00:
                    if not prefix:
                        raise LexingError("invalidMarkerSegmentPrefix")
                    ecs.append(0xFF)
                    stuffed = True
                    break
                if b != 0xFF:
                    break
            if stuffed:
                continue
            if not _marker_valid(b):
                raise LexingError("invalidMarkerSegmentType")
            if b in (0xD8, 0xD9) or 0xD0 <= b <= 0xD7:
                return bytes(ecs), b, b""
            if self.pos + 2 > n:
                raise LexingError("truncatedMarkerSegmentHeader")
            ln = (d[self.pos] << 8) | d[self.pos + 1]
            self.pos += 2
            if ln < 2:
                raise LexingError("invalidMarkerSegmentLength")
            if self.pos + ln - 2 > n:
                raise LexingError("truncatedMarkerSegmentBody")
            body = d[self.pos:self.pos + ln - 2]
            self.pos += ln - 2
            return bytes(ecs), b, body
        raise LexingError("truncatedEntropyCodedSegment")


def _parse_dht(body):
    out, base = [], 0
    while base < len(body):
        if len(body) < base + 17:
            raise ParsingError("mismatched huffman segment")
        counts = body[base + 1:base + 17]
        total = sum(counts)
        if len(body) < base + 17 + total:
            raise ParsingError("mismatched huffman segment")
        cls, tgt = body[base] >> 4, body[base] & 15
        if cls > 1:
            raise ParsingError("invalidHuffmanTypeCode")
        if tgt > 3:
            raise ParsingError("invalidHuffmanTargetCode")
        if total > 256:
            raise ParsingError("invalidHuffmanTable")
        out.append((cls, tgt, L.HuffTable.make(counts, body[base + 17:base + 17 + total])))
        base += 17 + total
    return out


def _parse_dqt(body):
    out, base = [], 0
    while base < len(body):
        tgt, prec = body[base] & 15, body[base] >> 4
        if tgt > 3:
            raise ParsingError("invalidQuantizationTargetCode")
        if prec == 0:
            if len(body) < base + 65:
                raise ParsingError("mismatched quantization segment")
            out.append((tgt, np.frombuffer(body[base + 1:base + 65], dtype=np.uint8).astype(np.uint16), 0))
            base += 65
        elif prec == 1:  # decode.swift:431-438: 64 big-endian UInt16
            if len(body) < base + 129:
                raise ParsingError("mismatched quantization segment")
            out.append((tgt, np.frombuffer(body[base + 1:base + 129], dtype=">u2").astype(np.uint16), 1))
            base += 129
        else:
            raise ParsingError("invalidQuantizationPrecisionCode")
    return out


def _is_frame(m):
    return 0xC0 <= m <= 0xCF and m not in (0xC4, 0xC8, 0xCC)


def _push_quanta(s, qslot, tables):
    """Spectral.push(qi:quanta:) decode.swift:2546-2557: an 8-bit image shall not use a 16-bit quantisation table"""
    for tgt, q, prec in tables:
        if prec == 1 and s.precision <= 8:
            raise DecodingError("invalidScanQuantizationPrecision")
        s.quanta.append(q)
        qslot[tgt] = len(s.quanta) - 1


def _decompress(data, ctx, gpu_lexer=False, format=None, on_scan=None, resident=False, t81_intervals=False):
    lx = _Lexer(bytes(data))
    _, m, body = lx.segment()
    if m != 0xD8:
        raise DecodingError("missingStartOfImage")
    _, m, body = lx.segment()
    while 0xE0 <= m <= 0xEF or m == 0xFE:
        _, m, body = lx.segment()
    dc, ac, qslot = [None] * 4, [None] * 4, [None] * 4
    pend_q, interval = [], None
    while True:
        if _is_frame(m):
            if len(body) < 6:
                raise ParsingError("mismatched frame")
            precision, fh, fw, count = body[0], (body[1] << 8) | body[2], (body[3] << 8) | body[4], body[5]
            if len(body) != 3 * count + 6:
                raise ParsingError("mismatched frame")
            process = {0xC0: 0, 0xC1: 1, 0xC2: 2}.get(m)
            comps = []
            for i in range(count):
                cid, hv, tq = body[6 + 3 * i:9 + 3 * i]
                if tq & 15 > 3:
                    raise ParsingError("invalidFrameQuantizationSelectorCode")
                if any(c[0] == cid for c in comps):
                    raise ParsingError("duplicateFrameComponentIndex")
                comps.append((cid, hv >> 4, hv & 15, tq & 15))
            if fw <= 0:
                raise ParsingError("invalidFrameWidth")
            for cid, fx, fy, tq in comps:
                if not (1 <= fx <= 4 and 1 <= fy <= 4):
                    raise ParsingError("invalidFrameComponentSamplingFactor")
                if process == 0 and tq > 1:
                    raise ParsingError("invalidFrameQuantizationSelector")
            if process is None:
                raise DecodingError("unsupportedFrameCodingProcess")
            if (process == 0 and precision != 8) or precision not in (8, 12):
                raise ParsingError("invalidFramePrecision")  # decode.swift:700-768
            _, m, body = lx.segment()
            break
        if m == 0xDB:
            pend_q += _parse_dqt(body)
        elif m == 0xC4:
            for cls, tgt, t in _parse_dht(body):
                (dc if cls == 0 else ac)[tgt] = t
        elif m == 0xDD:
            if len(body) != 2:
                raise ParsingError("mismatched interval")
            interval = ((body[0] << 8) | body[1]) or None
        elif m in (0xDA, 0xDC, 0xD9, 0xD8) or 0xD0 <= m <= 0xD7:
            raise DecodingError("premature / unexpected segment")
        _, m, body = lx.segment()

    if format is not None:  # Format.recognize; planes in the order format.components lists them (jpeg.swift:1286-1329)
        if not format.recognize([c[0] for c in comps], precision):
            raise DecodingError("unrecognizedColorFormat")
        comps.sort(key=lambda c: list(format.components).index(c[0]))
    else:  # JPEG.Common.recognize (jpeg.swift:370-397)
        comps.sort(key=lambda c: c[0])
        if precision != 8 or not (len(comps) == 1 or (len(comps) == 3 and comps[1][0] == comps[0][0] + 1
                                                      and comps[2][0] == comps[0][0] + 2)):
            raise DecodingError("unrecognizedColorFormat")
    s = Spectral((fw, fh), [(c[1], c[2]) for c in comps], [c[0] for c in comps], process, ctx, precision, resident)
    qsel = {c[0]: c[3] for c in comps}
    _push_quanta(s, qslot, pend_q)
    approx = [[None] * 64 for _ in comps]  # Progression, jpeg.swift:1581-1634

    first = True
    while True:
        if _is_frame(m):
            raise DecodingError("duplicateFrameHeaderSegment")
        if m == 0xDB:
            _push_quanta(s, qslot, _parse_dqt(body))
        elif m == 0xC4:
            for cls, tgt, t in _parse_dht(body):
                (dc if cls == 0 else ac)[tgt] = t
        elif m == 0xDA:
            if len(body) < 4 or len(body) != 2 * body[0] + 4 or body[0] > 4:
                raise ParsingError("mismatched scan")
            count = body[0]
            hdr = [(body[1 + 2 * i], body[2 + 2 * i] >> 4, body[2 + 2 * i] & 15) for i in range(count)]
            for _, d_, a_ in hdr:
                if d_ > 3 or a_ > 3 or (process == 0 and (d_ > 1 or a_ > 1)):
                    raise ParsingError("invalidScanHuffmanSelector")
            band = (body[2 * count + 1], body[2 * count + 2] + 1)
            lo, hi = body[2 * count + 3] & 15, body[2 * count + 3] >> 4
            bits = (lo, None if hi == 0 else hi)
            if not (band[0] < band[1] and (bits[1] is None or lo < bits[1])):
                raise ParsingError("invalidScanProgressiveSubset")
            if process != 2:
                ok = band == (0, 64) and bits == (0, None) and count >= 1
            elif band == (0, 1):
                ok = (bits[1] is None or bits[1] == lo + 1) and count >= 1
            else:
                ok = band[0] >= 1 and 2 <= band[1] <= 64 and (bits[1] is None or bits[1] == lo + 1) and count == 1
            if not ok:
                raise ParsingError("invalidScanProgressiveSubset / component count")
            ecss, raw = [], None
            if gpu_lexer:
                # the host only finds the FF of the next non-RSTn marker; the GPU unstuffs, splits and checks the phases
                d, p = lx.d, lx.pos
                while True:
                    k = d.find(b"\xff", p)
                    if k < 0:
                        raise LexingError("truncatedEntropyCodedSegment")
                    q = k + 1
                    while q < lx.n and d[q] == 0xFF:
                        q += 1
                    if q >= lx.n:
                        raise LexingError("truncatedMarkerSegmentType")
                    if d[q] == 0 or 0xD0 <= d[q] <= 0xD7:
                        p = q + 1
                        continue
                    break
                raw = d[lx.pos:k]
                lx.pos = k
                _, m, body = lx.segment()
                ival = interval
            else:
                index = 0
                while True:
                    ecs, m, body = lx.segment(prefix=True)
                    ecss.append(ecs)
                    if not (0xD0 <= m <= 0xD7):
                        break
                    if (m & 15) != index % 8:
                        raise DecodingError("invalidRestartPhase")
                    index += 1
                if interval is not None:
                    ival = interval
                elif len(ecss) == 1:
                    ival = None
                else:
                    raise DecodingError("missingRestartIntervalSegment")
            ids = [p.comp_id for p in s.planes]
            for cid, _, _ in hdr:  # Progression.update (jpeg.swift:1597-1634)
                if cid not in ids:
                    continue
                ap = approx[ids.index(cid)]
                if not (ap[0] is not None or band[0] == 0):
                    raise DecodingError("invalidSpectralSelectionProgression")
                for z in range(band[0], band[1]):
                    if not (bits[1] == ap[z] and (ap[z] is None or lo < ap[z])):
                        raise DecodingError("invalidSuccessiveApproximationProgression")
                    ap[z] = lo
            comps_ = []
            volume = 0
            for cid, d_, a_ in hdr:
                if cid not in ids:
                    raise DecodingError("undefinedScanComponentReference")
                p = ids.index(cid)
                volume += s.planes[p].factor[0] * s.planes[p].factor[1]
                comps_.append((p, d_, a_))
            if not (volume <= 10 or count == 1):
                raise DecodingError("invalidScanSamplingVolume")
            if bits[1] is None and band[0] == 0:  # dequantize: decode.swift:3451-3498
                for cid, _, _ in hdr:
                    sel = qsel[cid]
                    if qslot[sel] is None:
                        raise DecodingError("undefinedScanQuantizationReference")
                    s.planes[ids.index(cid)].q = qslot[sel]
            if first:
                # decode.swift:3905-3924: the segment that follows the first scan may redefine the height (DNL); the library
                # decodes into planes of the final size (its `extend` never grows them), so the host resolves the height
                # BEFORE the call -- rows beyond it would be cropped by push(height:) anyway, rows short of it stay zero.
                if m == 0xDC:
                    if len(body) != 2:
                        raise ParsingError("mismatched height")
                    s.set_size((fw, (body[0] << 8) | body[1]))
                elif fh == 0:
                    raise DecodingError("missingHeightRedefinitionSegment")
            flags = (L.SCAN_EXTEND if first else 0) | (L.SCAN_T81 if t81_intervals else 0)
            if gpu_lexer:
                s.decode_scan_raw(band, bits, comps_, dc, ac, raw, ival, extend=flags)
            else:
                s.decode_scan(band, bits, comps_, dc, ac, ecss, ival, extend=flags)
            s.scans.append(Scan(band, bits, comps_))
            if first:
                if m == 0xDC:
                    _, m, body = lx.segment()
                first = False
            if on_scan is not None:
                on_scan(s, s.scans[-1])
            continue
        elif m == 0xDD:
            if len(body) != 2:
                raise ParsingError("mismatched interval")
            interval = ((body[0] << 8) | body[1]) or None
        elif m == 0xD9:
            return s
        elif m in (0xD8, 0xDC) or 0xD0 <= m <= 0xD7:
            raise DecodingError("unexpected segment")
        _, m, body = lx.segment()


# Spectral.compress(stream:) encode.swift:1918-1972 with explicit slot assignment (the reference derives it from
# scan lifetimes through a Dictionary whose order is per-process random, jpeg.swift:1388-1441)
def _seg(marker, tail=b""):
    if marker in (0xD8, 0xD9):
        return bytes([0xFF, marker])
    ln = len(tail) + 2
    return bytes([0xFF, marker, ln >> 8, ln & 255]) + tail


def _compress(s: Spectral, scans, quanta_slots, interval_mcus, jfif):
    scans = scans or s.scans
    out = bytearray(_seg(0xD8))
    if jfif:
        out += _seg(0xE0, b"JFIF\0" + bytes([1, 2, 2, 0, 1, 0, 1, 0, 0]))
    quanta_slots = quanta_slots or {p.q: min(i, 1) for i, p in enumerate(s.planes)}
    sof = bytes([s.precision, s.size[1] >> 8, s.size[1] & 255, s.size[0] >> 8, s.size[0] & 255, len(s.planes)])
    for p in s.planes:
        sof += bytes([p.comp_id, (p.factor[0] << 4) | p.factor[1], quanta_slots[p.q]])
    out += _seg(0xC2 if s.process == 2 else (0xC1 if s.process == 1 else 0xC0), sof)
    dqt = b""
    for qi, slot in sorted(quanta_slots.items(), key=lambda kv: kv[1]):
        if s.precision > 8:  # decode.swift:2528-2532: formats deeper than 8 bits write 16-bit tables
            dqt += bytes([0x10 | slot]) + np.asarray(s.quanta[qi], dtype=">u2").tobytes()
        else:
            dqt += bytes([slot]) + bytes(int(v) for v in s.quanta[qi])
    out += _seg(0xDB, dqt)
    if interval_mcus:
        out += _seg(0xDD, bytes([interval_mcus >> 8, interval_mcus & 255]))
    for sc in scans:
        ecs, dct, act = s.encode_scan(sc.band, sc.bits, sc.comps, interval_mcus)
        dht = b""
        for cls, tabs in ((0, dct), (1, act)):
            for slot, t in enumerate(tabs):
                if t.present:
                    counts, values = t.as_tuple()
                    dht += bytes([(cls << 4) | slot]) + counts + values
        if dht:
            out += _seg(0xC4, dht)
        sos = bytes([len(sc.comps)])
        for p, d_, a_ in sc.comps:
            sos += bytes([s.planes[p].comp_id, (d_ << 4) | a_])
        sos += bytes([sc.band[0], sc.band[1] - 1, ((0 if sc.bits[1] is None else sc.bits[1]) << 4) | sc.bits[0]])
        out += _seg(0xDA, sos) + ecs
    out += _seg(0xD9)
    return bytes(out)
