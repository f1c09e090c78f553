"""Batch plumbing over layer B of the C-ABI (device pointers): torch supplies device memory and the stream, the
library supplies every kernel.  Used by bench.py and the full-size tests."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import lib as L


def units(size, stride):
    return size // stride + (1 if size % stride else 0)


class Geometry:
    """plane geometry of JPEG.Data.Spectral for `size` and sampling `factors` (decode.swift:2456-2495)."""

    def __init__(self, size, factors):
        self.size = tuple(size)
        self.factors = [tuple(f) for f in factors]
        self.scale = (max(f[0] for f in factors), max(f[1] for f in factors))
        self.blocks = (units(size[0], 8 * self.scale[0]), units(size[1], 8 * self.scale[1]))
        self.units = [(units(size[0] * fx, 8 * self.scale[0]), units(size[1] * fy, 8 * self.scale[1]))
                      for fx, fy in factors]

    @property
    def n_planes(self):
        return len(self.factors)

    @property
    def total_blocks(self):
        return sum(ux * uy for ux, uy in self.units)


def sequential_scan(geo: Geometry, dc=(0, 1, 1), ac=(0, 1, 1)):
    d = L.ScanDesc()
    d.band_lo, d.band_hi, d.bit_lo, d.bit_hi = 0, 64, 0, L.BITS_MAX
    d.n_comp = geo.n_planes
    for i in range(geo.n_planes):
        d.comp[i].plane = i
        d.comp[i].factor_x, d.comp[i].factor_y = geo.factors[i]
        d.comp[i].dc, d.comp[i].ac = dc[i], ac[i]
    d.blocks_x, d.blocks_y = geo.blocks
    return d


class DeviceBuffers:
    """coefficient planes, 8-bit sample planes and RGB for n images of one geometry, as torch tensors"""

    def __init__(self, geo: Geometry, n_images: int, device):
        self.geo, self.n = geo, n_images
        self.coef = [torch.zeros((n_images, uy, ux, 64), dtype=torch.int16, device=device) for ux, uy in geo.units]
        self.samples = [torch.zeros((n_images, 8 * uy, 8 * ux), dtype=torch.uint8, device=device) for ux, uy in geo.units]
        self.rgb = torch.zeros((n_images, geo.size[1], geo.size[0], 3), dtype=torch.uint8, device=device)
        self.sp, self.pl = L.DevSpectral(), L.DevPlanar()
        self.sp.n_images = self.pl.n_images = n_images
        self.sp.n_planes = self.pl.n_planes = geo.n_planes
        self.pl.sample_bytes = 1
        for p, (ux, uy) in enumerate(geo.units):
            for dst in (self.sp.plane[p], self.pl.plane[p]):
                dst.image_stride = 64 * ux * uy
                dst.units_x, dst.units_y = ux, uy
                dst.factor_x, dst.factor_y = geo.factors[p]
            self.sp.plane[p].coef = self.coef[p].data_ptr()
            self.pl.plane[p].samples = self.samples[p].data_ptr()


def encode_frames(ctx: L.Context, frames: torch.Tensor, geo: Geometry, quanta, interval_mcus: int):
    """RGB frames (n, H, W, 3) uint8 on the device -> per-image (stuffed ECS bytes, tables[8]) through the GPU encode
    path (K4 colour+downsample, K5 FDCT+quantise, K6/K7 entropy encode).  quanta: n_planes x 64 uint16 (zig-zag)."""
    n = frames.shape[0]
    buf = DeviceBuffers(geo, n, frames.device)
    q = np.ascontiguousarray(quanta, dtype=np.uint16)
    ctx.check(ctx.L.jpeg_sm100_dev_rgb8_to_planar(ctx.h, frames.data_ptr(), geo.size[0], geo.size[1], C.byref(buf.pl)))
    ctx.check(ctx.L.jpeg_sm100_dev_fdct(ctx.h, C.byref(buf.pl), q.ctypes.data, 8, C.byref(buf.sp)))
    desc = sequential_scan(geo)
    stride = (geo.total_blocks * 64 * 3 + 4096 + 255) // 256 * 256  # generous: 3 bytes per coefficient
    out = torch.zeros((n, stride), dtype=torch.uint8, device=frames.device)
    lens = torch.zeros(n, dtype=torch.int64, device=frames.device)
    tables = (L.HuffTable * (8 * n))()
    ctx.check(ctx.L.jpeg_sm100_dev_encode_scan(ctx.h, C.byref(desc), C.byref(buf.sp), interval_mcus, tables,
                                               out.data_ptr(), stride, lens.data_ptr()))
    torch.cuda.synchronize()
    lens_h = lens.cpu().tolist()
    assert max(lens_h) <= stride, "entropy-coded segment overflowed its buffer"
    host = out.cpu().numpy()
    return [host[i, :lens_h[i]] for i in range(n)], tables, buf


def unstuff_split(ecs: np.ndarray):
    """what the reference lexer hands to the decoder (decode.swift:130-190): FF00 -> FF, split at RSTn.
    Returns (unstuffed bytes, interval lengths)."""
    a = np.asarray(ecs, dtype=np.uint8)
    ff = np.flatnonzero(a == 0xFF)
    nxt = a[ff + 1]
    keep = np.ones(a.size, dtype=bool)
    keep[ff[nxt == 0x00] + 1] = False
    rst = ff[(nxt >= 0xD0) & (nxt <= 0xD7)]
    keep[rst] = False
    keep[rst + 1] = False
    # interval boundaries measured in kept bytes
    kept_before = np.cumsum(keep) - keep
    cuts = kept_before[rst] if rst.size else np.zeros(0, dtype=np.int64)
    data = a[keep]
    bounds = np.concatenate([[0], cuts, [data.size]]).astype(np.int64)
    return data, np.diff(bounds)


class DecodeInputs:
    """lexed scan data of a batch: concatenated unstuffed ECS + offsets (image-major) + tables"""

    def __init__(self, ecs_list, tables, n_ecs_expected=None):
        datas, lens = [], []
        for e in ecs_list:
            d, l = unstuff_split(e)
            datas.append(d)
            lens.append(l)
        self.n_images = len(ecs_list)
        self.n_ecs = len(lens[0])
        assert all(len(l) == self.n_ecs for l in lens)
        if n_ecs_expected is not None:
            assert self.n_ecs == n_ecs_expected, (self.n_ecs, n_ecs_expected)
        self.ecs = np.concatenate(datas + [np.zeros(64, np.uint8)])
        self.offsets = np.zeros(self.n_images * self.n_ecs + 1, dtype=np.uint64)
        np.cumsum(np.concatenate(lens), out=self.offsets[1:])
        self.ecs_bytes = int(self.offsets[-1])
        self.tables = tables
