#!/usr/bin/env python3
"""Device-resident timings of BASELINE.json's configs #3-#5 (parity-test configurations, not bench lines): CUDA events on the
launching stream around layer-B calls, results checked against the encoder's / decoder's own round trip.
usage: time_configs.py [3] [4] [5]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import LEVEL, quanta  # noqa: E402
from jpeg_b200 import batch, lib, synth  # noqa: E402

dev = torch.device("cuda:0")
stream = torch.cuda.current_stream()
ctx = lib.Context(0, stream=stream.cuda_stream)
q = np.stack([quanta(LEVEL, 0), quanta(LEVEL, 1), quanta(LEVEL, 1)])
which = [int(a) for a in sys.argv[1:]] or [3, 4, 5]


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def scan_desc(geo, band, bits, comps):
    d = lib.ScanDesc()
    d.band_lo, d.band_hi = band
    d.bit_lo, d.bit_hi = bits[0], (lib.BITS_MAX if bits[1] is None else bits[1])
    d.n_comp = len(comps)
    for i, (p, dc, ac) in enumerate(comps):
        d.comp[i].plane = p
        d.comp[i].factor_x, d.comp[i].factor_y = geo.factors[p]
        d.comp[i].dc, d.comp[i].ac = dc, ac
    d.blocks_x, d.blocks_y = geo.blocks
    return d


def frames_of(n, w, h, base=0):
    return torch.stack([synth.frame(base + i, w, h, dev) for i in range(n)])


if 3 in which:  # 3840x2160 baseline 4:2:0 encode at level 0.25, batch 64
    W, H, N = 3840, 2160, 64
    geo = batch.Geometry((W, H), [(2, 2), (1, 1), (1, 1)])
    buf = batch.DeviceBuffers(geo, N, dev)
    frames = torch.cat([frames_of(8, W, H, b) for b in range(0, N, 8)])
    desc = batch.sequential_scan(geo)
    stride = 8 << 20
    out = torch.zeros((N, stride), dtype=torch.uint8, device=dev)
    lens = torch.zeros(N, dtype=torch.int64, device=dev)
    tables = (lib.HuffTable * (8 * N))()
    res = {}
    for interval in (0, geo.blocks[0]):
        def enc():
            ctx.check(ctx.L.jpeg_sm100_dev_rgb8_to_planar(ctx.h, frames.data_ptr(), W, H, C.byref(buf.pl)))
            ctx.check(ctx.L.jpeg_sm100_dev_fdct(ctx.h, C.byref(buf.pl), q.ctypes.data, 8, C.byref(buf.sp)))
            ctx.check(ctx.L.jpeg_sm100_dev_encode_scan(ctx.h, C.byref(desc), C.byref(buf.sp), interval, tables, out.data_ptr(), stride,
                                                       lens.data_ptr()))
        ms = timed(enc)
        front = timed(lambda: (ctx.check(ctx.L.jpeg_sm100_dev_rgb8_to_planar(ctx.h, frames.data_ptr(), W, H, C.byref(buf.pl))),
                               ctx.check(ctx.L.jpeg_sm100_dev_fdct(ctx.h, C.byref(buf.pl), q.ctypes.data, 8, C.byref(buf.sp)))))
        res[f"interval_{interval}"] = {"ms": round(ms, 3), "ms_colour_fdct": round(front, 3), "Mpixels_per_s": round(N * W * H / ms / 1e3, 1),
                                       "ecs_bytes_per_frame": int(lens.sum().item() // N)}
    print(json.dumps({"config": 3, "what": "RGB8 (device) -> coefficients -> entropy-coded segments, 64 x 4K 4:2:0, level 0.25 "
                      "(the emit pass waits for the host's optimal-table construction: one stream sync inside)", **res}))
    del buf, frames, out

if 4 in which:  # 1920x1080 progressive (4 scans), batch 256
    W, H, N = 1920, 1080, 256
    geo = batch.Geometry((W, H), [(2, 2), (1, 1), (1, 1)])
    src = batch.DeviceBuffers(geo, N, dev)
    for b in range(0, N, 32):
        fr = frames_of(32, W, H, b)
        part = batch.DeviceBuffers(geo, 32, dev)
        ctx.check(ctx.L.jpeg_sm100_dev_rgb8_to_planar(ctx.h, fr.data_ptr(), W, H, C.byref(part.pl)))
        ctx.check(ctx.L.jpeg_sm100_dev_fdct(ctx.h, C.byref(part.pl), q.ctypes.data, 8, C.byref(part.sp)))
        torch.cuda.synchronize()
        for p in range(3):
            src.coef[p][b:b + 32].copy_(part.coef[p])
        del part, fr
    scans = [((0, 1), (0, None), [(0, 0, 0), (1, 1, 0), (2, 1, 0)], geo.blocks[0]),
             ((1, 64), (0, None), [(0, 0, 0)], geo.units[0][0]),
             ((1, 64), (0, None), [(1, 0, 0)], geo.units[1][0]),
             ((1, 64), (0, None), [(2, 0, 0)], geo.units[2][0])]
    stride = 2 << 20
    enc_ms, inputs = [], []
    for band, bits, comps, width in scans:
        d = scan_desc(geo, band, bits, comps)
        out = torch.zeros((N, stride), dtype=torch.uint8, device=dev)
        lens = torch.zeros(N, dtype=torch.int64, device=dev)
        tables = (lib.HuffTable * (8 * N))()
        enc_ms.append(timed(lambda: ctx.check(ctx.L.jpeg_sm100_dev_encode_scan(ctx.h, C.byref(d), C.byref(src.sp), width, tables,
                                                                                out.data_ptr(), stride, lens.data_ptr())), reps=1))
        lh = lens.cpu().tolist()
        host = out.cpu().numpy()
        di = batch.DecodeInputs([host[i, :lh[i]] for i in range(N)], list(tables))
        inputs.append((d, width, di, (lib.HuffTable * (8 * N))(*list(tables)),
                       torch.from_numpy(di.ecs).to(dev), torch.from_numpy(di.offsets.view(np.int64)).to(dev)))
        del out
    dst = batch.DeviceBuffers(geo, N, dev)
    d_st = torch.zeros(N, dtype=torch.int32, device=dev)
    per_scan = []
    for k, (d, width, di, tarr, d_ecs, d_off) in enumerate(inputs):
        flags = lib.SCAN_FRESH if k == 0 else 0
        per_scan.append(timed(lambda: ctx.check(ctx.L.jpeg_sm100_dev_decode_scan(ctx.h, C.byref(d), d_ecs.data_ptr(), d_off.data_ptr(), di.n_ecs,
                                                                                  width, flags, tarr, 0, C.byref(dst.sp), d_st.data_ptr())), reps=1))

    def full():
        for k, (d, width, di, tarr, d_ecs, d_off) in enumerate(inputs):
            ctx.check(ctx.L.jpeg_sm100_dev_decode_scan(ctx.h, C.byref(d), d_ecs.data_ptr(), d_off.data_ptr(), di.n_ecs, width,
                                                       lib.SCAN_FRESH if k == 0 else 0, tarr, 0, C.byref(dst.sp), d_st.data_ptr()))
        ctx.check(ctx.L.jpeg_sm100_dev_idct(ctx.h, C.byref(dst.sp), q.ctypes.data, 8, C.byref(dst.pl)))
        ctx.check(ctx.L.jpeg_sm100_dev_planar_to_rgb8(ctx.h, C.byref(dst.pl), W, H, 0, dst.rgb.data_ptr()))
    ms = timed(full, reps=2)
    assert d_st.cpu().abs().sum().item() == 0
    for p in range(3):
        assert torch.equal(dst.coef[p], src.coef[p]), p
    print(json.dumps({"config": 4, "what": "256 x 1920x1080 4:2:0, 4-scan progression of examples/recompress, DRI = one row: scans -> RGB8",
                      "ms": round(ms, 3), "Mpixels_per_s": round(N * W * H / ms / 1e3, 1), "decode_ms_per_scan": [round(x, 3) for x in per_scan],
                      "encode_ms_per_scan": [round(x, 3) for x in enc_ms], "ecs_bytes_per_frame": [int(i[2].ecs_bytes // N) for i in inputs]}))
    del src, dst, inputs

if 5 in which:  # 4000x3000 4:4:4 decode -> re-encode, 64 frames per GPU
    W, H, N = 4000, 3000, 64
    geo = batch.Geometry((W, H), [(1, 1), (1, 1), (1, 1)])
    ecs_all, tabs_all = [], []
    for b in range(0, N, 4):
        ecs, tabs, enc = batch.encode_frames(ctx, frames_of(4, W, H, b), geo, q, geo.blocks[0])
        ecs_all += [e.copy() for e in ecs]
        tabs_all += list(tabs)
        del enc
    torch.cuda.empty_cache()
    di = batch.DecodeInputs(ecs_all, tabs_all)
    tarr = (lib.HuffTable * (8 * N))(*tabs_all)
    desc = batch.sequential_scan(geo)
    buf = batch.DeviceBuffers(geo, N, dev)
    d_ecs, d_off = torch.from_numpy(di.ecs).to(dev), torch.from_numpy(di.offsets.view(np.int64)).to(dev)
    d_st = torch.zeros(N, dtype=torch.int32, device=dev)
    stride = 16 << 20
    out = torch.zeros((N, stride), dtype=torch.uint8, device=dev)
    lens = torch.zeros(N, dtype=torch.int64, device=dev)
    tabs2 = (lib.HuffTable * (8 * N))()
    dec = lambda: ctx.check(ctx.L.jpeg_sm100_dev_decode_scan(ctx.h, C.byref(desc), d_ecs.data_ptr(), d_off.data_ptr(), di.n_ecs, geo.blocks[0],
                                                              lib.SCAN_FRESH, tarr, 0, C.byref(buf.sp), d_st.data_ptr()))
    enc = lambda: ctx.check(ctx.L.jpeg_sm100_dev_encode_scan(ctx.h, C.byref(desc), C.byref(buf.sp), geo.blocks[0], tabs2, out.data_ptr(), stride,
                                                              lens.data_ptr()))
    ms_dec = timed(dec)
    ms_enc = timed(enc)
    assert d_st.cpu().abs().sum().item() == 0
    lh = lens.cpu().tolist()
    host = out[:2].cpu().numpy()
    for i in range(2):
        assert host[i, :lh[i]].tobytes() == ecs_all[i].tobytes(), i
    print(json.dumps({"config": 5, "what": "64 x 4000x3000 4:4:4 baseline, DRI = 500 MCUs: entropy decode to Spectral, re-encode from the coefficients "
                      "(one GPU's share of the 512-frame configuration)", "decode_ms": round(ms_dec, 3), "encode_ms": round(ms_enc, 3),
                      "roundtrip_Mpixels_per_s": round(N * W * H / (ms_dec + ms_enc) / 1e3, 1), "ecs_bytes_per_frame": int(di.ecs_bytes // N)}))
