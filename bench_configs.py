"""The other configurations of BASELINE.json, timed inside bench.py's run (one rank's share each; bench.py aggregates over ranks):

    #3  3840x2160 baseline 4:2:0 ENCODE at compression level 0.25, batch 64             (RGB8 in HBM -> entropy-coded segments)
    #4  1920x1080 progressive (the 4-scan progression of examples/recompress) DECODE, batch 256      (scans -> RGB8)
    #5  12-Mpixel 4:4:4 decode -> re-encode round trip, 64 frames per GPU (512 sharded across 8 GPUs) (scan -> Spectral -> scan)
    layer A: whole files through the staged host API (Spectral.decompress -> RGB8), next to the CPU oracle on the same file

Device-resident timings with CUDA events on the launching stream, inputs far larger than L2; every configuration checks its
result (round trips on every rank, one frame against the CPU oracle on rank 0) and says so in `parity_checked`.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from jpeg_b200 import batch, lib, synth  # noqa: E402


def _timed(stream, fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def _frames(n, w, h, dev, base):
    return torch.stack([synth.frame(base + i, w, h, dev) for i in range(n)])


def _scan_desc(geo, band, bits, comps):
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


def _oracle_tables(O, tabs, i):
    mk = lambda t: O.HuffSpec.make(bytes(t.counts), bytes(t.values)) if t.present else O.HuffSpec()
    return [mk(t) for t in tabs[8 * i:8 * i + 4]], [mk(t) for t in tabs[8 * i + 4:8 * i + 8]]


def config3(ctx, dev, q, rank, n_frames=64):
    """#3: 64 x 4K 4:2:0 RGB8 -> coefficients -> entropy-coded segments, level 0.25, the reference-faithful single segment per
    image (its encoder never writes DRI) and with DRI = one MCU row."""
    stream = torch.cuda.current_stream()
    W, H, N = 3840, 2160, n_frames
    geo = batch.Geometry((W, H), [(2, 2), (1, 1), (1, 1)])
    buf = batch.DeviceBuffers(geo, N, dev)
    frames = torch.cat([_frames(min(8, N - b), W, H, dev, rank * N + b) for b in range(0, N, 8)])
    desc = batch.sequential_scan(geo)
    stride = 8 << 20
    out = torch.zeros((N, stride), dtype=torch.uint8, device=dev)
    lens = torch.zeros(N, dtype=torch.int64, device=dev)
    tables = (lib.HuffTable * (8 * N))()
    res = {}
    for key, interval in (("no_dri", 0), ("dri_row", geo.blocks[0])):
        def enc():
            ctx.check(ctx.L.jpeg_sm100_dev_rgb8_to_planar(ctx.h, frames.data_ptr(), W, H, C.byref(buf.pl)))
            ctx.check(ctx.L.jpeg_sm100_dev_fdct(ctx.h, C.byref(buf.pl), q.ctypes.data, 8, C.byref(buf.sp)))
            ctx.check(ctx.L.jpeg_sm100_dev_encode_scan(ctx.h, C.byref(desc), C.byref(buf.sp), interval, tables, out.data_ptr(), stride,
                                                       lens.data_ptr()))
        res[key] = round(_timed(stream, enc, 2), 3)
    # parity: the bytes decode back to the coefficients on every rank; on rank 0 one frame equals the CPU oracle's encoder output
    lens_h = lens.cpu().tolist()
    host = out[:2].cpu().numpy()
    ecs = [host[i, :lens_h[i]] for i in range(2)]
    di = batch.DecodeInputs(ecs, list(tables)[:16], n_ecs_expected=geo.blocks[1])
    back = batch.DeviceBuffers(geo, 2, dev)
    d_ecs, d_off = torch.from_numpy(di.ecs).to(dev), torch.from_numpy(di.offsets.view(np.int64)).to(dev)
    d_st = torch.zeros(2, dtype=torch.int32, device=dev)
    tarr = (lib.HuffTable * 16)(*list(tables)[:16])
    ctx.check(ctx.L.jpeg_sm100_dev_decode_scan(ctx.h, C.byref(desc), d_ecs.data_ptr(), d_off.data_ptr(), di.n_ecs, geo.blocks[0],
                                               lib.SCAN_FRESH, tarr, 0, C.byref(back.sp), d_st.data_ptr()))
    torch.cuda.synchronize()
    ok = d_st.cpu().abs().sum().item() == 0 and all(torch.equal(back.coef[p], buf.coef[p][:2]) for p in range(3))
    oracle_checked = False
    if rank == 0:
        from oracle import oracle as O
        rgb = frames[0].cpu().numpy()
        planes = O.decompose(O.pack_rgb(rgb), geo.factors)
        s = O.Spectral.create((W, H), geo.factors)
        for p in range(3):
            s.coefficients(p)[...] = O.fdct_plane(planes[p], q[p])
        want, _, _ = s.encode_scan((0, 64), (0, None), [0, 1, 2], [0, 1, 1], [0, 1, 1], geo.blocks[0])
        ok = ok and ecs[0].tobytes() == want
        oracle_checked = True
    return {"what": "64 x 3840x2160 4:2:0 RGB8 (HBM) -> colour + downsample -> FDCT + quantise -> optimal Huffman tables -> entropy-coded "
                    "segments, level 0.25", "ms": res["no_dri"], "ms_with_dri": res["dri_row"], "pixels_per_gpu": N * W * H,
            "Mpixels_per_s": round(N * W * H / res["no_dri"] / 1e3, 1), "ecs_bytes_per_frame": int(sum(lens_h) // N),
            "parity_checked": bool(ok), "oracle_checked_on_rank0": oracle_checked}


PROGRESSION4 = lambda geo: [((0, 1), (0, None), [(0, 0, 0), (1, 1, 0), (2, 1, 0)], geo.blocks[0]),
                            ((1, 64), (0, None), [(0, 0, 0)], geo.units[0][0]),
                            ((1, 64), (0, None), [(1, 0, 0)], geo.units[1][0]),
                            ((1, 64), (0, None), [(2, 0, 0)], geo.units[2][0])]


def config4(ctx, dev, q, rank, n_frames=64):
    """#4: 256 x 1920x1080 4:2:0, the 4-scan progression of examples/recompress/main.swift:15-26, DRI = one row: scans -> RGB8"""
    stream = torch.cuda.current_stream()
    W, H, N = 1920, 1080, 4 * n_frames
    geo = batch.Geometry((W, H), [(2, 2), (1, 1), (1, 1)])
    src = batch.DeviceBuffers(geo, N, dev)
    for b in range(0, N, 32):
        m = min(32, N - b)
        fr = _frames(m, W, H, dev, rank * N + b)
        part = batch.DeviceBuffers(geo, m, dev)
        ctx.check(ctx.L.jpeg_sm100_dev_rgb8_to_planar(ctx.h, fr.data_ptr(), W, H, C.byref(part.pl)))
        ctx.check(ctx.L.jpeg_sm100_dev_fdct(ctx.h, C.byref(part.pl), q.ctypes.data, 8, C.byref(part.sp)))
        torch.cuda.synchronize()
        for p in range(3):
            src.coef[p][b:b + m].copy_(part.coef[p])
        del part, fr
    stride = 2 << 20
    enc_ms, inputs = [], []
    for band, bits, comps, width in PROGRESSION4(geo):
        d = _scan_desc(geo, band, bits, comps)
        out = torch.zeros((N, stride), dtype=torch.uint8, device=dev)
        lens = torch.zeros(N, dtype=torch.int64, device=dev)
        tables = (lib.HuffTable * (8 * N))()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for rep in range(2):  # (the first call grows the context's scratch buffers: cudaMalloc / cudaFree; the second is timed)
            e0.record(stream)
            ctx.check(ctx.L.jpeg_sm100_dev_encode_scan(ctx.h, C.byref(d), C.byref(src.sp), width, tables, out.data_ptr(), stride, lens.data_ptr()))
            e1.record(stream)
            torch.cuda.synchronize()
        enc_ms.append(e0.elapsed_time(e1))
        lh = lens.cpu().tolist()
        host = out.cpu().numpy()
        di = batch.DecodeInputs([host[i, :lh[i]] for i in range(N)], list(tables))
        inputs.append((d, width, di, (lib.HuffTable * (8 * N))(*list(tables)), torch.from_numpy(di.ecs).to(dev),
                       torch.from_numpy(di.offsets.view(np.int64)).to(dev), [host[0, :lh[0]].copy()], list(tables)[:8]))
        del out
    dst = batch.DeviceBuffers(geo, N, dev)
    d_st = torch.zeros(N, dtype=torch.int32, device=dev)

    def scan(k):
        d, width, di, tarr, d_ecs, d_off = inputs[k][:6]
        ctx.check(ctx.L.jpeg_sm100_dev_decode_scan(ctx.h, C.byref(d), d_ecs.data_ptr(), d_off.data_ptr(), di.n_ecs, width,
                                                   lib.SCAN_FRESH if k == 0 else 0, tarr, 0, C.byref(dst.sp), d_st.data_ptr()))

    def full():
        for k in range(4):
            scan(k)
        ctx.check(ctx.L.jpeg_sm100_dev_idct(ctx.h, C.byref(dst.sp), q.ctypes.data, 8, C.byref(dst.pl)))
        ctx.check(ctx.L.jpeg_sm100_dev_planar_to_rgb8(ctx.h, C.byref(dst.pl), W, H, 0, dst.rgb.data_ptr()))

    ms = _timed(stream, full, 3)
    per_scan = []
    for k in range(4):
        per_scan.append(round(_timed(stream, lambda: scan(k), 1), 3))
    full()
    torch.cuda.synchronize()
    ok = d_st.cpu().abs().sum().item() == 0 and all(torch.equal(dst.coef[p], src.coef[p]) for p in range(3))
    oracle_checked = False
    if rank == 0:  # frame 0 against the CPU oracle: its four scans decoded, inverse-transformed, upsampled and converted
        from oracle import oracle as O
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import jpegfile as J
        ref = O.Spectral.create((W, H), geo.factors, progressive=True)
        for p in range(3):
            ref.set_quanta(p, q[p])
        for (band, bits, comps, width), inp in zip(PROGRESSION4(geo), inputs):
            dct, act = _oracle_tables(O, inp[7], 0)
            ref.decode_scan(band, bits, [c[0] for c in comps], [c[1] for c in comps], [c[2] for c in comps], dct, act,
                            J.unstuff_split(inp[6][0].tobytes()), interval=width)
        ok = ok and np.array_equal(O.unpack_rgb(ref.to_rectangular()), dst.rgb[0].cpu().numpy())
        oracle_checked = True
    return {"what": f"{N} x 1920x1080 4:2:0, 4-scan progression of examples/recompress, DRI = one row: 4 scans -> coefficients -> RGB8",
            "ms": round(ms, 3), "pixels_per_gpu": N * W * H, "Mpixels_per_s": round(N * W * H / ms / 1e3, 1),
            "decode_ms_per_scan": per_scan, "encode_ms_per_scan": [round(x, 3) for x in enc_ms],
            "ecs_bytes_per_frame": [int(i[2].ecs_bytes // N) for i in inputs], "parity_checked": bool(ok),
            "oracle_checked_on_rank0": oracle_checked}


def config5(ctx, dev, q, rank, n_frames=64):
    """#5: 4000x3000 4:4:4 baseline, DRI = 500 MCUs: entropy decode to Spectral, re-encode from the coefficients; this rank's 64
    frames of the 512-frame configuration.  The re-encoded bytes must equal the input bytes (on every rank, every frame)."""
    stream = torch.cuda.current_stream()
    W, H, N = 4000, 3000, n_frames
    geo = batch.Geometry((W, H), [(1, 1), (1, 1), (1, 1)])
    ecs_all, tabs_all = [], []
    for b in range(0, N, 4):
        ecs, tabs, enc = batch.encode_frames(ctx, _frames(min(4, N - b), W, H, dev, rank * N + b), geo, q, geo.blocks[0])
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
    ms_dec = _timed(stream, dec, 2)
    ms_enc = _timed(stream, enc, 2)
    ok = d_st.cpu().abs().sum().item() == 0
    lh = lens.cpu().tolist()
    for i in range(N):  # byte for byte, every frame, on the device
        want = torch.from_numpy(ecs_all[i]).to(dev)
        ok = ok and lh[i] == want.numel() and bool(torch.equal(out[i, :lh[i]], want))
    oracle_checked = False
    if rank == 0:  # frame 0: the decoded coefficients equal the CPU oracle's
        from oracle import oracle as O
        dct, act = _oracle_tables(O, tabs_all, 0)
        data, ln = batch.unstuff_split(ecs_all[0])
        offs = np.concatenate([[0], np.cumsum(ln)])
        ref = O.Spectral.create((W, H), geo.factors)
        ref.decode_scan((0, 64), (0, None), [0, 1, 2], [0, 1, 1], [0, 1, 1], dct, act,
                        [data[offs[k]:offs[k + 1]].tobytes() for k in range(len(ln))], interval=geo.blocks[0])
        ok = ok and all(np.array_equal(ref.coefficients(p), buf.coef[p][0].cpu().numpy()) for p in range(3))
        oracle_checked = True
    return {"what": f"{N} x 4000x3000 4:4:4 baseline per GPU, DRI = 500 MCUs: entropy decode to Spectral, re-encode from the coefficients "
                    "(the 512-frame configuration is 64 frames on each of 8 GPUs)", "ms": round(ms_dec + ms_enc, 3), "decode_ms": round(ms_dec, 3),
            "encode_ms": round(ms_enc, 3), "pixels_per_gpu": N * W * H, "Mpixels_per_s": round(N * W * H / (ms_dec + ms_enc) / 1e3, 1),
            "ecs_bytes_per_frame": int(di.ecs_bytes // N), "parity_checked": bool(ok), "oracle_checked_on_rank0": oracle_checked}


PROGRESSION10 = [((0, 1), (1, None), [0, 1, 2]), ((1, 6), (1, None), [0]), ((1, 64), (1, None), [1]), ((1, 64), (1, None), [2]),
                 ((6, 64), (1, None), [0]), ((0, 1), (0, 1), [0, 1, 2]), ((1, 6), (0, 1), [0]), ((6, 64), (0, 1), [0]),
                 ((1, 64), (0, 1), [1]), ((1, 64), (0, 1), [2])]


def layer_a(device, q):
    """Layer A, the staged host API the Swift shim binds (one image, host buffers): whole-file decode to RGB8 of a 4K baseline
    file without DRI (what the reference's encoder writes: one entropy-coded segment, pushed with extend: true) and of the same
    image as a 10-scan progressive file with successive approximation, next to the CPU oracle's time on the same bytes."""
    from jpeg_b200 import host
    from oracle import oracle as O
    W, H = 3840, 2160
    factors = [(2, 2), (1, 1), (1, 1)]
    rgb = synth.frame(9001, W, H, "cpu").numpy()
    sp = host.Rectangular.pack(rgb, factors).decomposed().fdct([q[0], q[1], q[2]])
    files = {}
    sp.scans = [host.Scan((0, 64), (0, None), [(0, 0, 0), (1, 1, 1), (2, 1, 1)])]
    files["baseline_4k_no_dri"] = sp.compress()
    sp.process = 2
    sp.scans = [host.Scan(band, bits, [(c, 0 if c == 0 else 1, 0 if c == 0 else 1) for c in comps]) for band, bits, comps in PROGRESSION10]
    files["progressive_4k_10_scans"] = sp.compress()
    out = {}
    for name, data in files.items():
        best, got = None, None
        for _ in range(3):
            t0 = time.perf_counter()
            s = host.Spectral.decompress(data, gpu_lexer=True, resident=True)
            got = s.to_rgb8()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        stats = host.transfer_stats(s)
        t0 = time.perf_counter()
        ref = O.Spectral.decompress(data)
        want = O.unpack_rgb(ref.to_rectangular())
        t_cpu = time.perf_counter() - t0
        out[name] = {"file_bytes": len(data), "gpu_ms": round(best * 1e3, 2), "cpu_oracle_ms": round(t_cpu * 1e3, 1),
                     "speedup_vs_one_cpu_thread": round(t_cpu / best, 1), "Mpixels_per_s": round(W * H / best / 1e6, 1),
                     "equals_oracle": bool(np.array_equal(got, want)), **stats}
    out["call"] = "host.Spectral.decompress(bytes, gpu_lexer=True, resident=True).to_rgb8(): scans pushed one by one through the C-ABI " \
                  "(jpeg_sm100_spectral_* handle: planes stay in HBM between scans and stages), wall clock incl. Python container parsing"
    return out
