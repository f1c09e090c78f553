"""Full-size configurations of BASELINE.json on the GPU, checked through size-independent properties (entropy coding is
lossless, so encode -> decode must return the coefficients bit for bit; re-encoding decoded coefficients must return
the bytes) plus a direct oracle comparison on one frame per configuration."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    assert torch.cuda.is_available()
    from jpeg_b200 import batch, lib, synth
    from oracle import oracle as O
    dev = torch.device("cuda:0")
    ctx = lib.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    return dict(torch=torch, batch=batch, lib=lib, synth=synth, O=O, dev=dev, ctx=ctx)


def _quanta(O, level=0.25):
    return np.stack([O.quanta(level, 0), O.quanta(level, 1), O.quanta(level, 1)])


def _tables_to_oracle(O, tabs, i):
    mk = lambda t: O.HuffSpec.make(bytes(t.counts), bytes(t.values)) if t.present else O.HuffSpec()
    return [mk(t) for t in tabs[8 * i:8 * i + 4]], [mk(t) for t in tabs[8 * i + 4:8 * i + 8]]


def test_config2_4k_420_baseline_decode(env):
    """config #2: 3840x2160 baseline 4:2:0, DRI = 240 MCUs: GPU encode -> lexer -> GPU batch decode -> RGB."""
    t, b, lib, O, ctx, dev = env["torch"], env["batch"], env["lib"], env["O"], env["ctx"], env["dev"]
    W, H, N = 3840, 2160, 3
    geo = b.Geometry((W, H), [(2, 2), (1, 1), (1, 1)])
    assert geo.total_blocks == 194400 and geo.blocks == (240, 135)
    q = _quanta(O)
    frames = t.stack([env["synth"].frame(100 + i, W, H, dev) for i in range(N)])
    ecs, tabs, enc = b.encode_frames(ctx, frames, geo, q, geo.blocks[0])
    inputs = b.DecodeInputs(ecs, list(tabs), n_ecs_expected=135)
    # host-buffer batch entry point (H2D + K3 + K1 + K2 + D2H)
    rgb = np.zeros((N, H, W, 3), dtype=np.uint8)
    status = np.zeros(N, dtype=np.int32)
    desc = b.sequential_scan(geo)
    tarr = (lib.HuffTable * (8 * N))(*list(tabs))
    ctx.check(ctx.L.jpeg_sm100_decode_batch_rgb8(ctx.h, C.byref(desc), N, inputs.ecs.ctypes.data, inputs.offsets.ctypes.data,
                                                 inputs.n_ecs, geo.blocks[0], tarr, 0, q.ctypes.data, W, H, 0,
                                                 rgb.ctypes.data, status.ctypes.data))
    assert status.tolist() == [0] * N
    # layer B on the same inputs: coefficients must equal what the encoder was given
    buf = b.DeviceBuffers(geo, N, dev)
    d_ecs = t.from_numpy(inputs.ecs).to(dev)
    d_off = t.from_numpy(inputs.offsets.view(np.int64)).to(dev)
    d_st = t.zeros(N, dtype=t.int32, device=dev)
    ctx.check(ctx.L.jpeg_sm100_dev_decode_scan(ctx.h, C.byref(desc), d_ecs.data_ptr(), d_off.data_ptr(), inputs.n_ecs,
                                               geo.blocks[0], 0, tarr, 0, C.byref(buf.sp), d_st.data_ptr()))
    t.cuda.synchronize()
    for p in range(3):
        assert t.equal(buf.coef[p], enc.coef[p]), p
    # one frame against the oracle, end to end (coefficients from the stream, then IDCT + upsample + colour)
    dct, act = _tables_to_oracle(O, tabs, 1)
    data, lens = b.unstuff_split(ecs[1])
    offs = np.concatenate([[0], np.cumsum(lens)])
    parts = [data[offs[k]:offs[k + 1]].tobytes() for k in range(len(lens))]
    s = O.Spectral.create((W, H), geo.factors)
    for p in range(3):
        s.set_quanta(p, q[p])
    s.decode_scan((0, 64), (0, None), [0, 1, 2], [0, 1, 1], [0, 1, 1], dct, act, parts, interval=240)
    for p in range(3):
        assert np.array_equal(s.coefficients(p), enc.coef[p][1].cpu().numpy()), p
    assert np.array_equal(O.unpack_rgb(s.to_rectangular()), rgb[1])
    # pixels stay close to the source (level 0.25 is mild): a sanity bound, not a parity claim
    err = np.abs(rgb[1].astype(np.int32) - frames[1].cpu().numpy().astype(np.int32))
    assert err.mean() < 12.0


@pytest.mark.parametrize("tshift", [0, 4, 5, 6, 7])
def test_config2_fresh_planes_and_interval_threads(env, monkeypatch, tshift):
    """SCAN_FRESH: the library clears the planes itself (here they hold garbage), for every threads-per-interval setting of
    the parallel decoder (0 = the library's own choice) -- and for the sequential decoder."""
    t, b, lib, O, ctx, dev = env["torch"], env["batch"], env["lib"], env["O"], env["ctx"], env["dev"]
    if tshift:
        monkeypatch.setenv("JPEG_SM100_PAR_T", str(tshift))
    W, H, N = 3840, 2160, 2
    geo = b.Geometry((W, H), [(2, 2), (1, 1), (1, 1)])
    q = _quanta(O)
    frames = t.stack([env["synth"].frame(300 + i, W, H, dev) for i in range(N)])
    ecs, tabs, enc = b.encode_frames(ctx, frames, geo, q, geo.blocks[0])
    inputs = b.DecodeInputs(ecs, list(tabs), n_ecs_expected=135)
    desc = b.sequential_scan(geo)
    tarr = (lib.HuffTable * (8 * N))(*list(tabs))
    buf = b.DeviceBuffers(geo, N, dev)
    d_ecs = t.from_numpy(inputs.ecs).to(dev)
    d_off = t.from_numpy(inputs.offsets.view(np.int64)).to(dev)
    d_st = t.zeros(N, dtype=t.int32, device=dev)
    for c in buf.coef:
        c.fill_(0x5a5a)
    ctx.check(ctx.L.jpeg_sm100_dev_decode_scan(ctx.h, C.byref(desc), d_ecs.data_ptr(), d_off.data_ptr(), inputs.n_ecs,
                                               geo.blocks[0], lib.SCAN_FRESH, tarr, 0, C.byref(buf.sp), d_st.data_ptr()))
    t.cuda.synchronize()
    assert d_st.cpu().tolist() == [0] * N
    for p in range(3):
        assert t.equal(buf.coef[p], enc.coef[p]), p
    # fewer intervals than the image has rows: the rows nobody decodes must still come out zero
    short = 100
    # (decode only the first image: with fewer intervals the offsets of consecutive images are no longer contiguous)
    one = np.ascontiguousarray(inputs.offsets[:short + 1])
    for c in buf.coef:
        c.fill_(0x1234)
    one_sp = lib.DevSpectral()
    C.memmove(C.byref(one_sp), C.byref(buf.sp), C.sizeof(one_sp))
    one_sp.n_images = 1
    d_off1 = t.from_numpy(one.view(np.int64)).to(dev)
    ctx.check(ctx.L.jpeg_sm100_dev_decode_scan(ctx.h, C.byref(desc), d_ecs.data_ptr(), d_off1.data_ptr(), short,
                                               geo.blocks[0], lib.SCAN_FRESH, tarr, 0, C.byref(one_sp), d_st.data_ptr()))
    t.cuda.synchronize()
    for p, (ux, uy) in enumerate(geo.units):
        rows = short * geo.factors[p][1]
        assert t.equal(buf.coef[p][0, :rows], enc.coef[p][0, :rows]), p
        assert int(buf.coef[p][0, rows:].abs().sum().item()) == 0, p
        assert int((buf.coef[p][1] != 0x1234).sum().item()) == 0, p  # the second image was not part of the call


def test_single_ecs_without_restart_intervals(env, monkeypatch):
    """a reference-faithful file (no DRI: one entropy-coded segment per scan) through the parallel decoder: one CTA cuts the
    whole segment into 128 subsequences."""
    t, b, lib, O, ctx, dev = env["torch"], env["batch"], env["lib"], env["O"], env["ctx"], env["dev"]
    W, H = 1920, 1080
    geo = b.Geometry((W, H), [(2, 2), (1, 1), (1, 1)])
    q = _quanta(O)
    frames = t.stack([env["synth"].frame(400 + i, W, H, dev) for i in range(2)])
    ecs, tabs, enc = b.encode_frames(ctx, frames, geo, q, 0)
    inputs = b.DecodeInputs(ecs, list(tabs), n_ecs_expected=1)
    desc = b.sequential_scan(geo)
    tarr = (lib.HuffTable * 16)(*list(tabs))
    buf = b.DeviceBuffers(geo, 2, dev)
    d_ecs = t.from_numpy(inputs.ecs).to(dev)
    d_off = t.from_numpy(inputs.offsets.view(np.int64)).to(dev)
    d_st = t.zeros(2, dtype=t.int32, device=dev)
    for warm in ("64", "2048"):
        monkeypatch.setenv("JPEG_SM100_PAR_WARM", warm)
        for c in buf.coef:
            c.fill_(-1)
        ctx.check(ctx.L.jpeg_sm100_dev_decode_scan(ctx.h, C.byref(desc), d_ecs.data_ptr(), d_off.data_ptr(), 1,
                                                   lib.INTERVAL_NONE, lib.SCAN_FRESH, tarr, 0, C.byref(buf.sp), d_st.data_ptr()))
        t.cuda.synchronize()
        assert d_st.cpu().tolist() == [0, 0]
        for p in range(3):
            assert t.equal(buf.coef[p], enc.coef[p]), (warm, p)
    # the second image's segment cut short: the parallel decoder flags it, the sequential one reports the reference's error
    # (DecodingError.truncatedEntropyCodedSegment); the first image is unaffected
    cut = inputs.offsets.copy()
    cut[2] = cut[1] + (cut[2] - cut[1]) // 2
    d_cut = t.from_numpy(cut.view(np.int64)).to(dev)
    for c in buf.coef:
        c.fill_(-1)
    ctx.check(ctx.L.jpeg_sm100_dev_decode_scan(ctx.h, C.byref(desc), d_ecs.data_ptr(), d_cut.data_ptr(), 1, lib.INTERVAL_NONE,
                                               lib.SCAN_FRESH, tarr, 0, C.byref(buf.sp), d_st.data_ptr()))
    t.cuda.synchronize()
    assert d_st.cpu().tolist() == [0, lib.ERR_TRUNCATED_ECS]
    for p in range(3):
        assert t.equal(buf.coef[p][0], enc.coef[p][0]), p
    # garbage in the middle of the first image's segment: same verdict and, if it decodes, the same coefficients as the oracle
    junk = inputs.ecs.copy()
    rng = np.random.default_rng(2)
    mid = int(inputs.offsets[1]) // 2
    junk[mid:mid + 4096] = rng.integers(0, 256, 4096, dtype=np.uint8)
    d_junk = t.from_numpy(junk).to(dev)
    ctx.check(ctx.L.jpeg_sm100_dev_decode_scan(ctx.h, C.byref(desc), d_junk.data_ptr(), d_off.data_ptr(), 1, lib.INTERVAL_NONE,
                                               lib.SCAN_FRESH, tarr, 0, C.byref(buf.sp), d_st.data_ptr()))
    t.cuda.synchronize()
    dct, act = _tables_to_oracle(O, tabs, 0)
    ref = O.Spectral.create((W, H), geo.factors)
    try:
        ref.decode_scan((0, 64), (0, None), [0, 1, 2], [0, 1, 1], [0, 1, 1], dct, act, [junk[:int(inputs.offsets[1])].tobytes()])
        want = 0
    except O.OracleError as e:
        want = e.code
    assert d_st.cpu().tolist()[0] == want
    if want == 0:
        for p in range(3):
            assert np.array_equal(buf.coef[p][0].cpu().numpy(), ref.coefficients(p)), p


def test_config3_4k_420_encode(env):
    """config #3: 3840x2160 baseline 4:2:0 encode at level 0.25: coefficients and ECS bytes equal the oracle's, with the
    reference-faithful single ECS and with DRI = 240."""
    t, b, O, ctx, dev = env["torch"], env["batch"], env["O"], env["ctx"], env["dev"]
    W, H = 3840, 2160
    geo = b.Geometry((W, H), [(2, 2), (1, 1), (1, 1)])
    q = _quanta(O)
    frame = env["synth"].frame(7, W, H, dev)
    rgb = frame.cpu().numpy()
    planes = O.decompose(O.pack_rgb(rgb), geo.factors)
    s = O.Spectral.create((W, H), geo.factors)
    for p in range(3):
        s.coefficients(p)[...] = O.fdct_plane(planes[p], q[p])
    for interval in (0, 240):
        ecs, tabs, enc = b.encode_frames(ctx, frame[None], geo, q, interval)
        for p in range(3):
            assert np.array_equal(enc.coef[p][0].cpu().numpy(), s.coefficients(p)), p
        want, dct, act = s.encode_scan((0, 64), (0, None), [0, 1, 2], [0, 1, 1], [0, 1, 1], interval)
        assert ecs[0].tobytes() == want, interval
        for k in range(4):
            if dct[k].present:
                assert tabs[k].as_tuple() == dct[k].as_tuple()
            if act[k].present:
                assert tabs[4 + k].as_tuple() == act[k].as_tuple()


def test_config4_1080p_progressive_decode(env):
    """config #4: 1920x1080 4:2:0, the 4-scan progression of examples/recompress (DC all components, then Y / Cb / Cr
    AC 1..<64), per-row DRI.  Y has 135 block rows while the MCU grid covers 136 (hazard H3)."""
    t, b, lib, O, ctx, dev = env["torch"], env["batch"], env["lib"], env["O"], env["ctx"], env["dev"]
    from jpeg_b200 import host
    W, H = 1920, 1080
    geo = b.Geometry((W, H), [(2, 2), (1, 1), (1, 1)])
    assert geo.blocks == (120, 68) and geo.units[0] == (240, 135)
    q = _quanta(O)
    frame = env["synth"].frame(11, W, H, dev)
    _, _, enc = b.encode_frames(ctx, frame[None], geo, q, 0)  # coefficients via K4 + K5
    src = host.Spectral((W, H), geo.factors, process=2)
    for p in range(3):
        src.planes[p].coef = enc.coef[p][0].cpu().numpy()
    dst = host.Spectral((W, H), geo.factors, process=2)
    ref = O.Spectral.create((W, H), geo.factors, progressive=True)
    for p in range(3):
        ref.coefficients(p)[...] = src.planes[p].coef
    scans = [((0, 1), (0, None), [(0, 0, 0), (1, 1, 0), (2, 1, 0)], geo.blocks[0]),
             ((1, 64), (0, None), [(0, 0, 0)], geo.units[0][0]),
             ((1, 64), (0, None), [(1, 0, 0)], geo.units[1][0]),
             ((1, 64), (0, None), [(2, 0, 0)], geo.units[2][0])]
    import sys
    sys.path.insert(0, __file__.rsplit("/", 1)[0])
    import jpegfile as J
    for band, bits, comps, width in scans:
        ecs, dct, act = src.encode_scan(band, bits, comps, width)          # GPU encoder, DRI = one row
        want, _, _ = ref.encode_scan(band, bits, [c[0] for c in comps], [c[1] for c in comps], [c[2] for c in comps], width)
        assert ecs == want, (band, comps)
        dst.decode_scan(band, bits, comps, list(dct), list(act), J.unstuff_split(ecs), width)
    for p in range(3):
        assert np.array_equal(dst.planes[p].coef, src.planes[p].coef), p
    # pixels: the decoded image through the fused and the staged CUDA paths against the oracle's IDCT + upsampling + colour
    dst.quanta = [q[0].copy(), q[1].copy()]
    dst.planes[0].q, dst.planes[1].q, dst.planes[2].q = 0, 1, 1
    for p in range(3):
        ref.set_quanta(p, q[p])
    want = O.unpack_rgb(ref.to_rectangular())
    assert np.array_equal(dst.to_rgb8(), want)
    assert np.array_equal(dst.idct().interleaved().unpack_rgb(), want)


def test_config5_12mpix_444_roundtrip(env):
    """config #5: 4000x3000 4:4:4 baseline, DRI = 500 MCUs: decode to coefficients, re-encode from them: same bytes."""
    t, b, lib, O, ctx, dev = env["torch"], env["batch"], env["lib"], env["O"], env["ctx"], env["dev"]
    W, H = 4000, 3000
    geo = b.Geometry((W, H), [(1, 1), (1, 1), (1, 1)])
    assert geo.blocks == (500, 375) and geo.total_blocks == 562500
    q = _quanta(O)
    frame = env["synth"].frame(5, W, H, dev)
    ecs, tabs, enc = b.encode_frames(ctx, frame[None], geo, q, 500)
    inputs = b.DecodeInputs(ecs, list(tabs), n_ecs_expected=375)
    buf = b.DeviceBuffers(geo, 1, dev)
    desc = b.sequential_scan(geo)
    tarr = (lib.HuffTable * 8)(*list(tabs))
    d_ecs = t.from_numpy(inputs.ecs).to(dev)
    d_off = t.from_numpy(inputs.offsets.view(np.int64)).to(dev)
    d_st = t.zeros(1, dtype=t.int32, device=dev)
    ctx.check(ctx.L.jpeg_sm100_dev_decode_scan(ctx.h, C.byref(desc), d_ecs.data_ptr(), d_off.data_ptr(), inputs.n_ecs, 500,
                                               0, tarr, 0, C.byref(buf.sp), d_st.data_ptr()))
    t.cuda.synchronize()
    assert d_st.item() == 0
    for p in range(3):
        assert t.equal(buf.coef[p], enc.coef[p]), p
    # re-encode from the decoded coefficients
    stride = 64 << 20
    out = t.zeros(stride, dtype=t.uint8, device=dev)
    ln = t.zeros(1, dtype=t.int64, device=dev)
    tabs2 = (lib.HuffTable * 8)()
    ctx.check(ctx.L.jpeg_sm100_dev_encode_scan(ctx.h, C.byref(desc), C.byref(buf.sp), 500, tabs2, out.data_ptr(), stride,
                                               ln.data_ptr()))
    t.cuda.synchronize()
    again = out[:ln.item()].cpu().numpy()
    assert again.tobytes() == ecs[0].tobytes()
    # the oracle on the same frame: coefficients from the stream, the re-encoded bytes and tables, and (below) the pixels
    dct, act = _tables_to_oracle(O, tabs, 0)
    data, lens = b.unstuff_split(ecs[0])
    offs = np.concatenate([[0], np.cumsum(lens)])
    parts = [data[offs[k]:offs[k + 1]].tobytes() for k in range(len(lens))]
    ref = O.Spectral.create((W, H), geo.factors)
    for p in range(3):
        ref.set_quanta(p, q[p])
    ref.decode_scan((0, 64), (0, None), [0, 1, 2], [0, 1, 1], [0, 1, 1], dct, act, parts, interval=500)
    for p in range(3):
        assert np.array_equal(ref.coefficients(p), buf.coef[p][0].cpu().numpy()), p
    want, dct2, act2 = ref.encode_scan((0, 64), (0, None), [0, 1, 2], [0, 1, 1], [0, 1, 1], 500)
    assert again.tobytes() == want
    for k in range(4):
        if dct2[k].present:
            assert tabs2[k].as_tuple() == dct2[k].as_tuple()
        if act2[k].present:
            assert tabs2[4 + k].as_tuple() == act2[k].as_tuple()
    # 4:4:4 colour fast path against the generic kernel on the same planes
    ctx.check(ctx.L.jpeg_sm100_dev_idct(ctx.h, C.byref(buf.sp), q.ctypes.data, 8, C.byref(buf.pl)))
    ctx.check(ctx.L.jpeg_sm100_dev_planar_to_rgb8(ctx.h, C.byref(buf.pl), W, H, 0, buf.rgb.data_ptr()))
    il = t.zeros((H, W, 3), dtype=t.int16, device=dev)
    ctx.check(ctx.L.jpeg_sm100_dev_interleave(ctx.h, C.byref(buf.pl), W, H, 0, il.data_ptr()))
    rgb2 = t.zeros((H, W, 3), dtype=t.uint8, device=dev)
    ctx.check(ctx.L.jpeg_sm100_dev_unpack_rgb8(ctx.h, il.data_ptr(), W * H, 3, rgb2.data_ptr()))
    t.cuda.synchronize()
    assert t.equal(buf.rgb[0], rgb2)
    assert np.array_equal(buf.rgb[0].cpu().numpy(), O.unpack_rgb(ref.to_rectangular()))


@pytest.mark.parametrize("size,factors", [((3840, 2160), [(2, 2), (1, 1), (1, 1)]), ((4000, 3000), [(1, 1), (1, 1), (1, 1)]),
                                          ((1936, 1081), [(2, 2), (1, 1), (1, 1)]), ((48, 33), [(2, 2), (1, 1), (1, 1)]),
                                          ((1040, 7), [(1, 1), (1, 1), (1, 1)])])
def test_colour_kernels_agree(env, monkeypatch, size, factors):
    """K2: the TMA-store kernels (rows staged in shared memory, bulk copies out), the direct-store fast kernels and the
    reference-literal generic kernel produce the same RGB from the same random planes."""
    t, b, ctx, dev = env["torch"], env["batch"], env["ctx"], env["dev"]
    W, H = size
    geo = b.Geometry(size, factors)
    buf = b.DeviceBuffers(geo, 2, dev)
    g = t.Generator(device=dev)
    g.manual_seed(5)
    for s_ in buf.samples:
        s_.copy_(t.randint(0, 256, s_.shape, generator=g, device=dev, dtype=t.uint8))
    out = {}
    for mode in ("tma", "direct", "default", "generic"):
        monkeypatch.setenv("JPEG_SM100_COLOR", mode)
        rgb = t.full((2, H, W, 3), 7, dtype=t.uint8, device=dev)
        ctx.check(ctx.L.jpeg_sm100_dev_planar_to_rgb8(ctx.h, C.byref(buf.pl), W, H, 0, rgb.data_ptr()))
        t.cuda.synchronize()
        out[mode] = rgb
    assert t.equal(out["tma"], out["generic"])
    assert t.equal(out["direct"], out["generic"])
    assert t.equal(out["default"], out["generic"])
    # ... and the oracle's interleaved(cosite: false) + unpack(as: RGB) on the first image's planes
    O = env["O"]
    planes = [s_[0].cpu().numpy().astype(np.uint16) for s_ in buf.samples]
    want = O.unpack_rgb(O.interleave(planes, geo.units, geo.factors, size, False))
    assert np.array_equal(out["default"][0].cpu().numpy(), want)


@pytest.mark.parametrize("size,factors", [((3840, 2160), [(2, 2), (1, 1), (1, 1)]), ((4000, 3000), [(1, 1), (1, 1), (1, 1)]),
                                          ((64, 48), [(2, 2), (1, 1), (1, 1)]), ((40, 24), [(1, 1), (1, 1), (1, 1)]),
                                          ((50, 34), [(2, 2), (1, 1), (1, 1)])])
def test_forward_colour_kernels_agree(env, monkeypatch, size, factors):
    """K4: the fused RGB8 -> YCbCr-planes kernels against the reference-literal per-plane kernel (itself checked against the
    oracle in test_gpu_parity) on random pixels; (50, 34) is not a whole number of MCUs and must take the generic path."""
    t, b, ctx, dev = env["torch"], env["batch"], env["ctx"], env["dev"]
    W, H = size
    geo = b.Geometry(size, factors)
    g = t.Generator(device=dev)
    g.manual_seed(17)
    rgb = t.randint(0, 256, (2, H, W, 3), generator=g, device=dev, dtype=t.uint8)
    out = {}
    for mode in ("default", "generic"):
        monkeypatch.setenv("JPEG_SM100_COLOR", mode)
        buf = b.DeviceBuffers(geo, 2, dev)
        for s_ in buf.samples:
            s_.fill_(9)
        ctx.check(ctx.L.jpeg_sm100_dev_rgb8_to_planar(ctx.h, rgb.data_ptr(), W, H, C.byref(buf.pl)))
        t.cuda.synchronize()
        out[mode] = [s_.clone() for s_ in buf.samples]
    for p in range(3):
        assert t.equal(out["default"][p], out["generic"][p]), p
    # ... and the oracle's pack + decomposed() on the first image
    O = env["O"]
    want = O.decompose(O.pack_rgb(rgb[0].cpu().numpy()), geo.factors)
    for p in range(3):
        assert np.array_equal(out["default"][p][0].cpu().numpy().astype(np.uint16), want[p]), p


def test_whole_file_4k_baseline_without_dri_through_host(env):
    """The reference's real call path for a baseline file: Spectral.decompress pushes the (only) scan with extend: true
    (decode.swift:3892-3904), one entropy-coded segment for the whole image (its encoder never writes DRI).  The scan must run
    through the cluster form of the subsequence-parallel decoder, and equal the oracle's decode of the same file."""
    O = env["O"]
    from jpeg_b200 import host
    W, H = 3840, 2160
    factors = [(2, 2), (1, 1), (1, 1)]
    q = _quanta(O)
    rgb = env["synth"].frame(21, W, H, env["dev"]).cpu().numpy()
    sp = host.Rectangular.pack(rgb, factors).decomposed().fdct([q[0], q[1], q[2]])
    sp.scans = [host.Scan((0, 64), (0, None), [(0, 0, 0), (1, 1, 1), (2, 1, 1)])]
    data = sp.compress()
    s = host.Spectral.decompress(data)
    ref = O.Spectral.decompress(data)
    for p in range(3):
        assert np.array_equal(s.planes[p].coef, ref.coefficients(p)), p
        assert np.array_equal(s.planes[p].coef, sp.planes[p].coef), p
    assert np.array_equal(s.to_rgb8(), O.unpack_rgb(ref.to_rectangular()))
    # the same file cut after 60 % of its bytes: the lexer fails on the missing EOI exactly as the reference's does; the scan
    # alone, pushed with extend, ends in truncatedEntropyCodedSegment in both
    import sys
    sys.path.insert(0, __file__.rsplit("/", 1)[0])
    import jpegfile as J
    ecs = [e for m, _, e in J.split(data) if m == 0xDA][0]
    part = J.unstuff_split(ecs)[0]
    part = part[:int(len(part) * 0.6)]
    dct = [t if t.present else None for t in _scan_tables(data, 0)]
    act = [t if t.present else None for t in _scan_tables(data, 1)]
    cut = host.Spectral((W, H), factors)
    with pytest.raises(env["lib"].JpegSm100Error) as ei:
        cut.decode_scan((0, 64), (0, None), [(0, 0, 0), (1, 1, 1), (2, 1, 1)], dct, act, [part], None, extend=True)
    assert ei.value.code == env["lib"].ERR_TRUNCATED_ECS


def _scan_tables(data, cls):
    """the four table slots of class `cls` (0 = DC, 1 = AC) as the file defines them"""
    import jpegfile as J
    from jpeg_b200 import lib
    slots = [lib.HuffTable() for _ in range(4)]
    for m, body, _ in J.split(data):
        if m == 0xC4:
            for c, tgt, counts, values in J.parse_dht(body):
                if c == cls:
                    slots[tgt] = lib.HuffTable.make(counts, values)
    return slots


@pytest.mark.parametrize("size,n,band", [((3840, 2160), 2, 0), ((1936, 1081), 3, 0), ((48, 33), 2, 0), ((16, 16), 1, 0), ((8, 8), 1, 0),
                                         ((400, 300), 2, 8), ((1000, 250), 2, 8), ((391, 517), 1, 10), ((776, 129), 1, 9)])
def test_fused_idct_colour_equals_staged_kernels(env, monkeypatch, size, n, band):
    """K1+K2 fused (k_idct_rgb420 behind jpeg_sm100_dev_spectral_to_rgb8): the same bytes as jpeg_sm100_dev_idct followed by
    jpeg_sm100_dev_planar_to_rgb8 (JPEG_SM100_FUSE=0), on random sparse coefficients -- strips of 24 MCUs with partial last
    strips, bands of MCU rows (forced short so that several bands and their re-transformed halo rows are exercised), odd sizes
    whose luma plane is one block short of the MCU grid -- and the oracle's idct + interleaved + unpack on the first image."""
    t, b, ctx, dev, O = env["torch"], env["batch"], env["ctx"], env["dev"], env["O"]
    W, H = size
    factors = [(2, 2), (1, 1), (1, 1)]
    geo = b.Geometry(size, factors)
    buf = b.DeviceBuffers(geo, n, dev)
    g = t.Generator(device=dev)
    g.manual_seed(W * 31 + H)
    for c in buf.coef:
        dense = t.rand(c.shape, generator=g, device=dev) < 0.15
        c.copy_(t.where(dense, t.randint(-40, 41, c.shape, generator=g, device=dev), t.zeros((), dtype=t.int64, device=dev)).to(t.int16))
        c[..., 0] = t.randint(-200, 201, c.shape[:-1], generator=g, device=dev).to(t.int16)
    q = _quanta(O)
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("JPEG_SM100_FUSE", mode)
        if band:
            monkeypatch.setenv("JPEG_SM100_FUSE_BAND", str(band))
        rgb = t.full((n, H, W, 3), 7, dtype=t.uint8, device=dev)
        ctx.check(ctx.L.jpeg_sm100_dev_spectral_to_rgb8(ctx.h, C.byref(buf.sp), q.ctypes.data, W, H, 0, rgb.data_ptr()))
        t.cuda.synchronize()
        out[mode] = rgb
    assert t.equal(out["1"], out["0"])
    if W * H <= 1936 * 1081:
        planes = [O.idct_plane(c[0].cpu().numpy(), q[p]) for p, c in enumerate(buf.coef)]
        want = O.unpack_rgb(O.interleave(planes, geo.units, geo.factors, size, False))
        assert np.array_equal(out["1"][0].cpu().numpy(), want)
