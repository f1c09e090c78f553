"""GPU parity tests: every stage of the CUDA path, called through the C-ABI, against the CPU oracle (which is itself
pinned bit-exactly to the reference's golden vectors by test_oracle_golden.py) and against the committed golden
digests directly.  Bar: bit-exact everywhere -- the kernels evaluate the reference's binary32 arithmetic op for op
(no FMA), so even the float stages are expected to match exactly, not merely within +-1.
"""
import hashlib

import numpy as np
import pytest

import jpegfile as J
from conftest import golden_bytes

pytestmark = pytest.mark.gpu


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


@pytest.fixture(scope="module")
def H():
    import torch
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    from jpeg_b200 import host
    host.default_context()  # raises if libjpeg_sm100.so is missing or no sm_100 device is present
    return host


def _to_lib_tables(H, specs):
    from jpeg_b200 import lib
    return [lib.HuffTable.make(bytes(t.counts), bytes(t.values)) if t.present else None for t in specs]


# ------------------------------------------------------------------------------------------------ whole files
def test_golden_files_through_gpu(manifest, H, O):
    """tests/regression/tests.swift:39-138 run through the CUDA path: coefficients, planes, YCbCr and RGB."""
    for v in manifest["decode"]:
        data = golden_bytes(v["jpeg"])
        s = H.Spectral.decompress(data)
        ref = O.Spectral.decompress(data)
        assert s.size == ref.size and s.ncomp == ref.ncomp
        for p in range(s.ncomp):
            assert np.array_equal(s.planes[p].coef, ref.coefficients(p)), (v["jpeg"], "coefficients", p)
            assert np.array_equal(s.quanta[s.planes[p].q], ref.quanta(p))
        planar = s.idct()
        for p, pl in enumerate(ref.idct()):
            assert np.array_equal(planar.planes[p], pl), (v["jpeg"], "idct", p)
        rect = planar.interleaved()
        rgb, ycc = rect.unpack_rgb(), rect.unpack_ycc()
        assert sha(rgb.tobytes()) == v["rgb_sha256"], v["jpeg"]
        if "ycc_sha256" in v:
            assert sha(ycc.tobytes()) == v["ycc_sha256"], v["jpeg"]
        if "planes_sha256" in v:
            for p, pl in enumerate(s.idct_u8()):
                assert sha(pl.tobytes()) == v["planes_sha256"][p]
        # fused coefficients -> RGB8 path
        assert sha(s.to_rgb8().tobytes()) == v["rgb_sha256"], (v["jpeg"], "fused")


def test_resident_spectral_through_gpu(manifest, H, O):
    """The image kept in HBM across scans and stages (jpeg_sm100_spectral, what JPEG.Context does with its one Spectral per
    file, decode.swift:3565-3587): same coefficients, planes and pixels as the oracle; only scan bytes go up, only status words
    come down until a stage result is asked for; the host planes materialise on demand and are read-only."""
    files = [v for v in manifest["decode"]] + [{"jpeg": r} for r in manifest["restart"]]
    for v in files:
        data = golden_bytes(v["jpeg"])
        ref = O.Spectral.decompress(data)
        for gpu_lexer in (False, True):
            s = H.Spectral.decompress(data, resident=True, gpu_lexer=gpu_lexer)
            t = H.transfer_stats(s)
            assert t["h2d_bytes"] <= len(data) + 64 * 1024, (v["jpeg"], t)      # scan bytes (+ offsets, table sets)
            assert t["d2h_bytes"] <= 64 * (len(s.scans) + 1), (v["jpeg"], t)    # one status word per scan (+ the lexer's counts)
            rgb = s.to_rgb8()
            assert np.array_equal(rgb, O.unpack_rgb(ref.to_rectangular())), v["jpeg"]
            if "rgb_sha256" in v:
                assert sha(rgb.tobytes()) == v["rgb_sha256"], v["jpeg"]
            assert H.transfer_stats(s)["d2h_bytes"] - t["d2h_bytes"] == rgb.size                 # exactly the pixels
            for a, b in zip(s.idct().planes, ref.idct()):
                assert np.array_equal(a, b), v["jpeg"]
            for p in range(s.ncomp):
                assert np.array_equal(s.planes[p].coef, ref.coefficients(p)), (v["jpeg"], p)   # materialised here, once
            with pytest.raises(ValueError):
                s.planes[0].coef[0, 0, 0] = 1
            with pytest.raises(ValueError):
                s.planes[0].coef = np.zeros(1)
            # the first scan re-encoded from the resident image and from host planes: the same bytes and tables
            sc = s.scans[0]
            host_side = H.Spectral(s.size, [p.factor for p in s.planes], process=s.process)
            for p in range(s.ncomp):
                host_side.planes[p].coef = np.array(s.planes[p].coef)
            got, want = s.encode_scan(sc.band, sc.bits, sc.comps), host_side.encode_scan(sc.band, sc.bits, sc.comps)
            assert got[0] == want[0] and [t_.as_tuple() for t_ in got[1] + got[2]] == [t_.as_tuple() for t_ in want[1] + want[2]]
    # DNL: the resident image is resized on the device (Spectral.set(height:), decode.swift:2484-2495)
    data = golden_bytes("gold/color-sequential-1.jpg")
    h = O.Spectral.decompress(data).size[1]
    for fh, dh in ((0, h), (h, h - 21), (h, h + 40)):
        mod = J.with_dnl(data, fh, dh)
        s, ref = H.Spectral.decompress(mod, resident=True), O.Spectral.decompress(mod)
        assert s.size == ref.size
        for p in range(s.ncomp):
            assert np.array_equal(s.planes[p].coef, ref.coefficients(p)), (fh, dh, p)
    s = H.Spectral.decompress(data, resident=True)
    s.set_size((s.size[0], s.size[1] - 100))                       # after the scans: crop on the device, the rest intact
    ref = O.Spectral.decompress(data)
    for p in range(s.ncomp):
        uy = s.planes[p].units[1]
        assert np.array_equal(s.planes[p].coef, ref.coefficients(p)[:uy]), p


def test_restart_files_through_gpu(manifest, H, O):
    for rel in manifest["restart"]:
        data = golden_bytes(rel)
        s = H.Spectral.decompress(data)
        ref = O.Spectral.decompress(data)
        for p in range(s.ncomp):
            assert np.array_equal(s.planes[p].coef, ref.coefficients(p)), rel
        rgb_ref = O.unpack_rgb(ref.to_rectangular())
        assert np.array_equal(s.to_rgb8(), rgb_ref), rel


def test_height_defined_by_dnl_through_gpu(H, O):
    """decode.swift:3905-3924: frame height 0 + DNL, and DNL heights below / above the frame's (crop / zero extension)."""
    for rel in ("gold/color-sequential-1.jpg", "gold/color-progressive-2.jpg", "gold/grayscale-sequential-1.jpg"):
        data = golden_bytes(rel)
        h = O.Spectral.decompress(data).size[1]
        for fh, dh in ((0, h), (h, h - 21), (h, h + 40)):
            mod = J.with_dnl(data, fh, dh)
            if dh > h and "progressive" in rel:  # the later scans of a progressive file run out of data on the taller image
                from jpeg_b200 import lib
                with pytest.raises(lib.JpegSm100Error) as ei:
                    H.Spectral.decompress(mod)
                assert ei.value.code == lib.ERR_TRUNCATED_ECS
                continue
            s, ref = H.Spectral.decompress(mod), O.Spectral.decompress(mod)
            assert s.size == ref.size
            for p in range(s.ncomp):
                assert np.array_equal(s.planes[p].coef, ref.coefficients(p)), (rel, fh, dh, p)
            assert np.array_equal(s.to_rgb8(), O.unpack_rgb(ref.to_rectangular())), (rel, fh, dh)


def test_cosited_and_generic_upsampling(H, O, manifest):
    """interleaved(cosite: true) and non-4:2:0 factors go through the generic kernel."""
    eb = manifest["encode_basic"]
    w, h = eb["size"]
    rgb = np.frombuffer(golden_bytes(eb["rgb"]), dtype=np.uint8).reshape(h, w, 3)[:203, :187]
    il = O.pack_rgb(rgb)
    for factors in ([(2, 2), (1, 1), (1, 1)], [(2, 1), (1, 1), (1, 1)], [(1, 2), (1, 1), (1, 1)], [(4, 2), (1, 1), (2, 1)],
                    [(1, 1), (1, 1), (1, 1)]):
        planes = O.decompose(il, factors)
        us = [(p.shape[1] // 8, p.shape[0] // 8) for p in planes]
        for cosited in (False, True):
            ref = O.interleave(planes, us, factors, (187, 203), cosited)
            got = H.Planar((187, 203), us, factors, [p.copy() for p in planes]).interleaved(cosite=cosited)
            assert np.array_equal(got.values, ref), (factors, cosited)
            assert np.array_equal(got.unpack_rgb(), O.unpack_rgb(ref))


def test_four_planes_of_12_bit_samples(H, O):
    """N4 (kernel level): a custom JPEG.Format with four components and 12-bit samples (examples/custom-color) reaches the same
    kernels -- interleaved(cosite:) over four 16-bit planes with mixed sampling factors, centred and co-sited -- and decomposed()
    back."""
    rng = np.random.default_rng(12)
    size = (150, 70)
    factors = [(2, 2), (1, 1), (1, 2), (2, 1)]
    sx, sy = 2, 2
    us = [(-(-size[0] * fx // (8 * sx)), -(-size[1] * fy // (8 * sy))) for fx, fy in factors]
    planes = [rng.integers(0, 4096, size=(8 * uy, 8 * ux)).astype(np.uint16) for ux, uy in us]
    for cosited in (False, True):
        ref = O.interleave(planes, us, factors, size, cosited)
        got = H.Planar(size, us, factors, [p.copy() for p in planes]).interleaved(cosite=cosited)
        assert np.array_equal(got.values, ref), cosited
    il = rng.integers(0, 4096, size=(size[1], size[0], 4)).astype(np.uint16)
    back = H.Rectangular(size, factors, il).decomposed()
    want = O.decompose(il, factors)
    for p in range(4):
        assert np.array_equal(back.planes[p], want[p]), p


def test_online_decoding_through_gpu(manifest, H):
    """examples/decode-online: the ten per-scan RGB dumps of the reference, reproduced from the image as it stands after each
    scan (staged idct().interleaved().unpack(as: RGB) and the fused coefficients -> RGB8 call); a progressive file with DRI."""
    v = manifest["decode_online"]
    data = golden_bytes(v["jpeg"])
    for gpu_lexer in (False, True):
        got = []

        def capture(s, scan):
            got.append((sha(s.idct().interleaved().unpack_rgb().tobytes()), sha(s.to_rgb8().tobytes())))

        H.Spectral.decompress(data, gpu_lexer=gpu_lexer, on_scan=capture)
        assert [g[0] for g in got] == v["rgb_sha256"], gpu_lexer
        assert [g[1] for g in got] == v["rgb_sha256"], gpu_lexer


def test_custom_format_file_through_gpu(manifest, H, O):
    """N4 (host level): examples/custom-color/output.jpg -- a user-defined JPEG.Format (components 4-7, 12-bit samples, 16-bit
    DQT), ten progressive scans with two-component DC scans of unequal sampling -- decoded, inverse-transformed, forward-
    transformed from the example's pixels and written back through the CUDA path.  The writer must return the reference's
    file byte for byte."""
    from test_oracle_golden import custom_color_pixels
    cc = manifest["custom_color"]
    fmt = H.Format(tuple(cc["format"][0]), cc["format"][1])
    w, h = cc["size"]
    factors = [tuple(f) for f in cc["factors"]]
    data = golden_bytes(cc["jpeg"])
    ref = O.Spectral.decompress(data, format=(list(fmt.components), fmt.precision))
    with pytest.raises(H.DecodingError, match="unrecognizedColorFormat"):
        H.Spectral.decompress(data)
    for gpu_lexer in (False, True):
        s = H.Spectral.decompress(data, format=fmt, gpu_lexer=gpu_lexer)
        assert s.size == (w, h) and s.precision == 12 and [p.comp_id for p in s.planes] == [4, 5, 6, 7]
        for p in range(4):
            assert np.array_equal(s.planes[p].coef, ref.coefficients(p)), (gpu_lexer, p)
            assert np.array_equal(s.quanta[s.planes[p].q], ref.quanta(p)), (gpu_lexer, p)
    # inverse path: 12-bit IDCT, four planes interleaved
    planar = s.idct()
    for p, pl in enumerate(ref.idct()):
        assert np.array_equal(planar.planes[p], pl), ("idct", p)
    rect = planar.interleaved()
    assert rect.precision == 12 and np.array_equal(rect.values, ref.to_rectangular())
    # forward path from the example's pixels (main.swift:123-134): decomposed() over four planes, 12-bit FDCT + quantiser
    px = custom_color_pixels()
    q = [np.array([1, 2, 2, 3, 3, 3] + [10] * 58, dtype=np.uint16), np.array([1] + [100] * 63, dtype=np.uint16)]
    pl = H.Rectangular((w, h), factors, px, precision=12).decomposed()
    for p, want in enumerate(O.decompose(px, factors)):
        assert np.array_equal(pl.planes[p], want), ("decomposed", p)
    enc = pl.fdct([q[0], q[0], q[0], q[1]], comp_ids=list(fmt.components), process=2)
    for p in range(4):
        assert np.array_equal(enc.planes[p].coef, ref.coefficients(p)), ("fdct", p)
    # the writer: same progression, quanta 0 shared by R, G, B (main.swift:143-149)
    enc.planes[1].q = enc.planes[2].q = enc.planes[0].q
    out = enc.compress(scans=s.scans, quanta_slots={enc.planes[0].q: 0, enc.planes[3].q: 1}, jfif=False)
    assert out == data and sha(out) == cc["file_sha256"]


# ------------------------------------------------------------------------------------------------ single stages
@pytest.mark.parametrize("ux,uy", [(1, 1), (3, 2), (17, 5), (128, 1), (129, 3), (40, 40)])
def test_idct_random_blocks(H, O, ux, uy):
    """K1 on random coefficients incl. values that drive the output far outside 0..255 (clamp + trunc)."""
    rng = np.random.default_rng(ux * 100 + uy)
    coef = rng.integers(-64, 64, size=(uy, ux, 64)).astype(np.int16)
    coef[..., 0] = rng.integers(-1024, 1024, size=(uy, ux))
    coef[rng.random((uy, ux)) < 0.1] *= 30
    q = rng.integers(1, 255, size=64).astype(np.uint16)
    s = H.Spectral((8 * ux, 8 * uy), [(1, 1)])
    s.planes[0].coef = coef
    s.quanta.append(q)
    s.planes[0].q = 1
    ref = O.idct_plane(coef, q)
    assert np.array_equal(s.idct().planes[0], ref)
    assert np.array_equal(s.idct_u8()[0], ref.astype(np.uint8))


@pytest.mark.parametrize("precision", [8, 12, 16])
def test_idct_fdct_other_precisions(H, O, precision):
    """N4 (first step): Plane.idct(quanta:precision:) / Plane.fdct(_:quanta:precision:) with the precision of a 12-bit or 16-bit
    frame (level shift 2^(P-1), clamp to 2^P - 1) and 16-bit quantisation tables, 16-bit samples in and out."""
    import ctypes as C
    ctx = H.default_context()
    rng = np.random.default_rng(precision)
    ux, uy = 19, 7
    top = (1 << precision) - 1
    coef = rng.integers(-40, 40, size=(uy, ux, 64)).astype(np.int16)
    coef[..., 0] = rng.integers(-top // 16, top // 16, size=(uy, ux))
    coef[rng.random((uy, ux)) < 0.1] *= 20
    q = rng.integers(1, 600 if precision > 8 else 255, size=64).astype(np.uint16)
    got = np.zeros((8 * uy, 8 * ux), dtype=np.uint16)
    ctx.check(ctx.L.jpeg_sm100_idct(ctx.h, C.c_void_p(coef.ctypes.data), ux, uy, C.c_void_p(q.ctypes.data), precision,
                                    C.c_void_p(got.ctypes.data)))
    assert np.array_equal(got, O.idct_plane(coef, q, precision)), precision
    samples = rng.integers(0, top + 1, size=(8 * uy, 8 * ux)).astype(np.uint16)
    back = np.zeros((uy, ux, 64), dtype=np.int16)
    ctx.check(ctx.L.jpeg_sm100_fdct(ctx.h, C.c_void_p(samples.ctypes.data), ux, uy, C.c_void_p(q.ctypes.data), precision,
                                    C.c_void_p(back.ctypes.data)))
    assert np.array_equal(back, O.fdct_plane(samples, q, precision)), precision


def test_idct_empty_plane(H):
    s = H.Spectral((8, 8), [(1, 1)])
    s.planes[0].coef = np.zeros((0, 0, 64), np.int16)
    s.planes[0].units = (0, 0)
    assert s.idct().planes[0].size == 0


def test_encode_front_end_stages(H, O, manifest):
    """K4/K5: RGB.pack, decomposed() for four sampling modes, fdct at two compression levels -- bit-exact."""
    eb = manifest["encode_basic"]
    w, h = eb["size"]
    rgb = np.frombuffer(golden_bytes(eb["rgb"]), dtype=np.uint8).reshape(h, w, 3)
    il = O.pack_rgb(rgb)
    for lum in ((1, 1), (1, 2), (2, 1), (2, 2)):
        factors = [lum, (1, 1), (1, 1)]
        rect = H.Rectangular.pack(rgb, factors)
        assert np.array_equal(rect.values, il)
        planar = rect.decomposed()
        ref_planes = O.decompose(il, factors)
        for p in range(3):
            assert np.array_equal(planar.planes[p], ref_planes[p]), (lum, p)
        for level in (0.0, 0.25, 2.0):
            q = [O.quanta(level, 0), O.quanta(level, 1), O.quanta(level, 1)]
            sp = planar.fdct(q)
            for p in range(3):
                assert np.array_equal(sp.planes[p].coef, O.fdct_plane(ref_planes[p], q[p])), (lum, level, p)


# ------------------------------------------------------------------------------------------------ scan decode, all kinds
def _oracle_scans(O, src, progression, rows):
    """encode `src` (oracle Spectral) scan by scan with restart intervals of `rows` rows; returns per-scan inputs"""
    out = []
    for band, bits, comps in progression:
        width = src.blocks[0] if len(comps) > 1 else src.units(comps[0])[0]
        ival = rows * width if rows else 0
        ecs, dct, act = src.encode_scan(band, bits, comps, [0] * len(comps), [0] * len(comps), ival)
        out.append((band, bits, comps, dct, act, J.unstuff_split(ecs), ival or None))
    return out


PROGRESSIVE = [((0, 1), (1, None), [0, 1, 2]), ((1, 6), (1, None), [0]), ((1, 64), (1, None), [1]),
               ((1, 64), (1, None), [2]), ((6, 64), (1, None), [0]), ((0, 1), (0, 1), [0, 1, 2]),
               ((1, 64), (0, 1), [0]), ((1, 64), (0, 1), [1]), ((1, 64), (0, 1), [2])]
BASELINE = [((0, 64), (0, None), [0, 1, 2])]
BASELINE_SPLIT = [((0, 64), (0, None), [0]), ((0, 64), (0, None), [1, 2])]


@pytest.mark.parametrize("name,progression", [("baseline", BASELINE), ("split", BASELINE_SPLIT),
                                              ("progressive", PROGRESSIVE)])
@pytest.mark.parametrize("rows", [0, 1, 2])
def test_scan_decode_all_kinds(H, O, name, progression, rows):
    """K3 against the oracle for the five scan kinds, interleaved and not, with and without restart intervals,
    on an odd-sized 4:2:0 image (partial MCUs: out-of-plane blocks are decoded and dropped, decode.swift:1470)."""
    src = O.Spectral.decompress(golden_bytes("gold/color-progressive-1.jpg"))
    fac = [src.factor(p) for p in range(3)]
    dst = H.Spectral(src.size, fac, process=2)
    for band, bits, comps, dct, act, parts, ival in _oracle_scans(O, src, progression, rows):
        dst.decode_scan(band, bits, [(c, 0, 0) for c in comps], _to_lib_tables(H, dct), _to_lib_tables(H, act), parts, ival)
    for p in range(3):
        assert np.array_equal(dst.planes[p].coef, src.coefficients(p)), (name, rows, p)


@pytest.mark.parametrize("tshift", [4, 5, 6, 7])
@pytest.mark.parametrize("warm", [32, 2048])
def test_scan_decode_parallel_variants(H, O, monkeypatch, tshift, warm):
    """the subsequence-parallel decoder (K3p) under every threads-per-interval setting and with a warm-up so short that the
    synchronisation rounds have real work: same coefficients and same error codes as the oracle."""
    from jpeg_b200 import lib
    monkeypatch.setenv("JPEG_SM100_PAR_T", str(tshift))
    monkeypatch.setenv("JPEG_SM100_PAR_WARM", str(warm))
    src = O.Spectral.decompress(golden_bytes("gold/color-progressive-1.jpg"))
    fac = [src.factor(p) for p in range(3)]
    for progression in (BASELINE, BASELINE_SPLIT):
        for rows in (0, 1, 3):
            dst = H.Spectral(src.size, fac, process=2)
            for band, bits, comps, dct, act, parts, ival in _oracle_scans(O, src, progression, rows):
                dst.decode_scan(band, bits, [(c, 0, 0) for c in comps], _to_lib_tables(H, dct), _to_lib_tables(H, act), parts, ival)
            for p in range(3):
                assert np.array_equal(dst.planes[p].coef, src.coefficients(p)), (rows, p)
    # DC-first scans (kind 1): from 64 subsequences per interval up, every subsequence is parsed once per block-in-MCU hypothesis
    # and the chain of exits picks the right one; Y and the chroma planes use different DC tables here, and the same one below
    for tables in ([0, 1, 1], [0, 0, 0]):
        for comps in ([0, 1, 2], [0], [1, 2]):
            for rows in (0, 2):
                width = src.blocks[0] if len(comps) > 1 else src.units(comps[0])[0]
                ecs, dct, act = src.encode_scan((0, 1), (1, None), comps, [tables[c] for c in comps], [0] * len(comps), rows * width)
                dst = H.Spectral(src.size, fac, process=2)
                ref = O.Spectral.create(src.size, fac, progressive=True)
                dst.decode_scan((0, 1), (1, None), [(c, tables[c], 0) for c in comps], _to_lib_tables(H, dct), _to_lib_tables(H, act),
                                J.unstuff_split(ecs), (rows * width) or None)
                ref.decode_scan((0, 1), (1, None), comps, [tables[c] for c in comps], [0] * len(comps), dct, act, J.unstuff_split(ecs),
                                interval=(rows * width) or O.INTERVAL_NONE)
                for p in range(3):
                    assert np.array_equal(dst.planes[p].coef, ref.coefficients(p)), (tables, comps, rows, p)
    # corrupt intervals: flagged by the parallel decoder, diagnosed by the sequential one
    (band, bits, comps, dct, act, parts, ival), = _oracle_scans(O, src, BASELINE, 1)
    rng = np.random.default_rng(11)
    for k in range(4):
        junk = list(parts)
        junk[k] = bytes(rng.integers(0, 256, len(parts[k]), dtype=np.uint8))
        if k == 3:
            junk[k] = junk[k][:len(junk[k]) // 3]
        got = H.Spectral(src.size, fac)
        want = O.Spectral.create(src.size, fac)
        try:
            got.decode_scan(band, bits, [(c, 0, 0) for c in comps], _to_lib_tables(H, dct), _to_lib_tables(H, act), junk, ival)
            code = 0
        except lib.JpegSm100Error as e:
            code = e.code
        try:
            want.decode_scan(band, bits, comps, [0] * 3, [0] * 3, dct, act, junk, interval=ival)
            wcode = 0
        except O.OracleError as e:
            wcode = e.code
        assert code == wcode, k
        if code == 0:
            for p in range(3):
                assert np.array_equal(got.planes[p].coef, want.coefficients(p)), (k, p)


@pytest.mark.parametrize("size,factors", [((97, 61), [(1, 1)]), ((200, 120), [(2, 1), (1, 1), (1, 1)]),
                                          ((131, 77), [(1, 2), (1, 1), (1, 1)]), ((160, 96), [(4, 1), (1, 1), (1, 1)]),
                                          ((150, 70), [(4, 2), (1, 1), (2, 1)]), ((96, 64), [(2, 2), (1, 1), (1, 1), (2, 2)]),
                                          ((1000, 40), [(1, 1), (1, 1), (1, 1)])])
@pytest.mark.parametrize("rows", [0, 1, 2])
def test_sequential_scan_geometries(H, O, size, factors, rows):
    """K3p over the geometries the benchmark does not touch -- up to 12 blocks per MCU, four components, partial MCUs on both
    edges, one interval per image or per 1-2 MCU rows -- on random sparse coefficients: entropy coding is lossless, so the
    oracle's encoder followed by the GPU decoder must return the coefficients, and the GPU encoder the oracle's bytes."""
    rng = np.random.default_rng(hash((size, len(factors), rows)) & 0xffff)
    src = O.Spectral.create(size, factors)
    n = len(factors)
    for p in range(n):
        c = src.coefficients(p)
        dense = rng.random(c.shape) < 0.08
        c[...] = np.where(dense, rng.integers(-60, 60, c.shape), 0).astype(np.int16)
        c[..., 0] = rng.integers(-500, 500, c.shape[:2])
        c[rng.random(c.shape[:2]) < 0.3, 1:] = 0           # DC-only blocks
        hot = rng.random(c.shape[:2]) < 0.05
        c[hot] = rng.integers(-1000, 1000, (int(hot.sum()), 64)).astype(np.int16)  # a few dense blocks with long codes
    comps = list(range(n))
    ival = rows * src.blocks[0] if n > 1 else rows * src.units(0)[0]
    ecs, dct, act = src.encode_scan((0, 64), (0, None), comps, [0] * n, [0] * n, ival)
    dst = H.Spectral(size, factors)
    dst.decode_scan((0, 64), (0, None), [(c, 0, 0) for c in comps], _to_lib_tables(H, dct), _to_lib_tables(H, act),
                    J.unstuff_split(ecs), ival or None)
    for p in range(n):
        assert np.array_equal(dst.planes[p].coef, src.coefficients(p)), p
    got, _, _ = dst.encode_scan((0, 64), (0, None), [(c, 0, 0) for c in comps], ival)
    assert got == ecs


def test_scan_decode_errors(H, O):
    """Error parity: truncated data, undefined tables, EOB run in a sequential scan -- same code as the oracle."""
    from jpeg_b200 import lib
    src = O.Spectral.decompress(golden_bytes("gold/color-sequential-1.jpg"))
    fac = [src.factor(p) for p in range(3)]
    (band, bits, comps, dct, act, parts, ival), = _oracle_scans(O, src, BASELINE, 1)
    dcl, acl = _to_lib_tables(H, dct), _to_lib_tables(H, act)
    cc = [(c, 0, 0) for c in comps]

    def run(parts_, dc_=dcl, ac_=acl, ival_=ival, H=H):
        dst = H.Spectral(src.size, fac)
        try:
            dst.decode_scan(band, bits, cc, dc_, ac_, parts_, ival_)
            return 0
        except lib.JpegSm100Error as e:
            return e.code

    def run_oracle(parts_, dc_=dct, ac_=act):
        dst = O.Spectral.create(src.size, fac)
        try:
            dst.decode_scan(band, bits, comps, [0] * 3, [0] * 3, dc_, ac_, parts_, interval=ival)
            return 0
        except O.OracleError as e:
            return e.code

    assert run(parts) == 0
    cut = [p for p in parts]
    cut[5] = cut[5][:len(cut[5]) // 2]
    assert run(cut) == run_oracle(cut) == lib.ERR_TRUNCATED_ECS
    empty = list(parts)
    empty[0] = b""
    assert run(empty) == run_oracle(empty) == lib.ERR_TRUNCATED_ECS
    assert run(parts, dc_=[None] * 4) == lib.ERR_UNDEFINED_DC
    assert run(parts, ac_=[None] * 4) == lib.ERR_UNDEFINED_AC
    # garbage bytes: whatever the oracle says (error code or success), we say too
    rng = np.random.default_rng(7)
    for k in range(6):
        junk = list(parts)
        junk[k] = bytes(rng.integers(0, 256, len(parts[k]), dtype=np.uint8))
        assert run(junk) == run_oracle(junk), k


def test_scan_decode_extend_flag(H, O):
    """extend = true (first scan): rows stop silently where the data ends (decode.swift:3214-3220)."""
    src = O.Spectral.decompress(golden_bytes("gold/grayscale-sequential-1.jpg"))
    (band, bits, comps, dct, act, parts, ival), = _oracle_scans(O, src, [((0, 64), (0, None), [0])], 0)
    half = [parts[0][:len(parts[0]) // 2]]
    from jpeg_b200 import lib
    dst = H.Spectral(src.size, [(1, 1)])
    with pytest.raises(lib.JpegSm100Error):
        dst.decode_scan(band, bits, [(0, 0, 0)], _to_lib_tables(H, dct), _to_lib_tables(H, act), half, None, extend=False)


@pytest.mark.parametrize("kind", ["baseline", "luma", "dc_first", "dc_first_luma"])
@pytest.mark.parametrize("rows", [0, 2])
def test_first_scan_extend_through_parallel_decoder(H, O, monkeypatch, kind, rows):
    """The reference pushes the first scan of every file with extend: true (decode.swift:3892-3904): rows stop silently where the
    data ends (3214-3220, 2906-2912).  Complete scans, scans that end after 25 / 50 / 75 % of the MCU rows (silent stop) and
    scans cut in the middle of a row (truncation error) -- same planes and same codes as the oracle, through the
    subsequence-parallel decoders (an interval that runs dry is flagged and redone by the sequential kernel)."""
    from jpeg_b200 import lib
    monkeypatch.setenv("JPEG_SM100_PAR_T", "4")  # (short DC-first intervals would otherwise take the one-thread-per-interval kernel)
    src = O.Spectral.decompress(golden_bytes("gold/color-progressive-1.jpg"))
    fac = [src.factor(p) for p in range(3)]
    band, bits = ((0, 64), (0, None)) if kind in ("baseline", "luma") else ((0, 1), (1, None))
    comps = [0, 1, 2] if kind in ("baseline", "dc_first") else [0]
    w, h = src.size
    mcu_h = 8 * src.scale[1] if len(comps) > 1 else 8

    def scan_of(height):
        """the scan of the image cropped to `height` pixel rows, encoded by the oracle"""
        part = O.Spectral.create((w, height), fac, progressive=True)
        for p in range(3):
            c = part.coefficients(p)
            c[...] = src.coefficients(p)[:c.shape[0], :c.shape[1]]
        width = part.blocks[0] if len(comps) > 1 else part.units(comps[0])[0]
        ecs, dct, act = part.encode_scan(band, bits, comps, [0] * len(comps), [0] * len(comps), rows * width)
        return J.unstuff_split(ecs), dct, act, (rows * width) or None

    def both(parts, dct, act, ival):
        got, want = H.Spectral(src.size, fac, process=2), O.Spectral.create(src.size, fac, progressive=True)
        try:
            got.decode_scan(band, bits, [(c, 0, 0) for c in comps], _to_lib_tables(H, dct), _to_lib_tables(H, act), parts, ival, extend=True)
            code = 0
        except lib.JpegSm100Error as e:
            code = e.code
        try:
            want.decode_scan(band, bits, comps, [0] * len(comps), [0] * len(comps), dct, act, parts,
                             interval=O.INTERVAL_NONE if ival is None else ival, extend=True)
            wcode = 0
        except O.OracleError as e:
            wcode = e.code
        assert code == wcode
        if code == 0:
            assert want.size == src.size  # (the oracle grows planes under `extend`; it must not have had to)
            for p in range(3):
                assert np.array_equal(got.planes[p].coef, want.coefficients(p)), p
        return code

    total_rows = -(-h // mcu_h)
    assert both(*scan_of(h)) == 0
    for frac in (0.25, 0.5, 0.75):
        k = max(1, int(total_rows * frac))
        assert both(*scan_of(k * mcu_h)) == 0, frac      # the data ends at a row boundary: silent stop
    parts, dct, act, ival = scan_of(h)
    for frac in (0.25, 0.5, 0.75):                        # cut mid-row: whatever the oracle says
        cut = list(parts)
        j = int(len(cut) * frac)
        cut[j] = cut[j][:max(1, int(len(cut[j]) * frac))]
        both(cut[:j + 1] if rows == 0 else cut, dct, act, ival)


def test_dc_refinement_scan_is_a_bit_gather(H, O):
    """kind 2 through k_decode_dc_refine: interleaved and single-component, with and without restart intervals, odd geometry,
    plus a scan that is one bit short (truncation, decode.swift:2775-2778)."""
    from jpeg_b200 import lib
    src = O.Spectral.decompress(golden_bytes("gold/color-progressive-1.jpg"))
    fac = [src.factor(p) for p in range(3)]
    for comps in ([0, 1, 2], [0], [2]):
        for rows in (0, 1, 3):
            dst = H.Spectral(src.size, fac, process=2)
            ref = O.Spectral.create(src.size, fac, progressive=True)
            scans = _oracle_scans(O, src, [((0, 1), (1, None), comps), ((0, 1), (0, 1), comps)], rows)
            for band, bits, cs, dct, act, parts, ival in scans:
                dst.decode_scan(band, bits, [(c, 0, 0) for c in cs], _to_lib_tables(H, dct), _to_lib_tables(H, act), parts, ival)
                ref.decode_scan(band, bits, cs, [0] * len(cs), [0] * len(cs), dct, act, parts, interval=O.INTERVAL_NONE if ival is None else ival)
            for p in range(3):
                assert np.array_equal(dst.planes[p].coef, ref.coefficients(p)), (comps, rows, p)
                if p in comps:
                    assert np.array_equal(dst.planes[p].coef[..., 0], src.coefficients(p)[..., 0]), (comps, rows, p)
            band, bits, cs, dct, act, parts, ival = scans[1]
            short = list(parts)
            short[-1] = short[-1][:-1]
            with pytest.raises(lib.JpegSm100Error) as ei:
                dst.decode_scan(band, bits, [(c, 0, 0) for c in cs], _to_lib_tables(H, dct), _to_lib_tables(H, act), short, ival)
            assert ei.value.code == lib.ERR_TRUNCATED_ECS


DEEP_REFINEMENT = [((0, 1), (0, None), [0, 1, 2]), ((1, 64), (2, None), [0]), ((1, 20), (1, None), [1]), ((20, 64), (1, None), [1]),
                   ((1, 64), (1, None), [2]), ((1, 64), (1, 2), [0]), ((1, 9), (0, 1), [0]), ((9, 64), (0, 1), [0]),
                   ((1, 64), (0, 1), [1]), ((1, 64), (0, 1), [2])]


@pytest.mark.parametrize("forced", ["1", "0"])
@pytest.mark.parametrize("rows", [0, 1, 3])
def test_ac_refinement_three_phase(H, O, monkeypatch, forced, rows):
    """kind 4 through k_acr_masks / k_acr_parse / k_acr_apply (forced on, and the one-thread-per-interval kernel forced for the same
    inputs): two refinement passes over split bands, with and without restart intervals, on a golden image and on sparse random
    planes whose end-of-band runs span hundreds of blocks; then truncated and garbage refinement data -- same planes and same
    error codes as the oracle (a flagged interval is left untouched and redone by the sequential kernel)."""
    from jpeg_b200 import lib
    monkeypatch.setenv("JPEG_SM100_ACR", forced)
    rng = np.random.default_rng(5)
    sources = [O.Spectral.decompress(golden_bytes("gold/color-progressive-1.jpg"))]
    sparse = O.Spectral.create((1000, 72), [(1, 1), (1, 1), (1, 1)], progressive=True)
    for p in range(3):
        c = sparse.coefficients(p)
        c[...] = np.where(rng.random(c.shape) < 0.02, rng.integers(-9, 10, c.shape), 0).astype(np.int16)
        c[rng.random(c.shape[:2]) < 0.9, 1:] = 0      # long runs of blocks without AC coefficients
        c[0, :40, 1:] = rng.integers(-3, 4, (40, 63))  # and a dense stretch: many correction bits per symbol
    sources.append(sparse)
    for src in sources:
        fac = [src.factor(p) for p in range(3)]
        for progression in (PROGRESSIVE, DEEP_REFINEMENT):
            dst = H.Spectral(src.size, fac, process=2)
            ref = O.Spectral.create(src.size, fac, progressive=True)
            scans = _oracle_scans(O, src, progression, rows)
            for band, bits, comps, dct, act, parts, ival in scans:
                dst.decode_scan(band, bits, [(c, 0, 0) for c in comps], _to_lib_tables(H, dct), _to_lib_tables(H, act), parts, ival)
                ref.decode_scan(band, bits, comps, [0] * len(comps), [0] * len(comps), dct, act, parts,
                                interval=O.INTERVAL_NONE if ival is None else ival)
            for p in range(3):
                assert np.array_equal(dst.planes[p].coef, ref.coefficients(p)), (rows, p)
                assert np.array_equal(dst.planes[p].coef, src.coefficients(p)), (rows, p)
        # damaged refinement scans: the last scan of the progression decoded from a good state, with bad data
        scans = _oracle_scans(O, src, PROGRESSIVE, rows)
        band, bits, comps, dct, act, parts, ival = scans[6]  # luma AC refinement
        cases = []
        cut = list(parts)
        cut[-1] = cut[-1][:len(cut[-1]) // 2]
        cases.append(cut)
        short = list(parts)
        short[0] = short[0][:-1]
        cases.append(short)
        for k in range(3):
            junk = list(parts)
            j = k % len(junk)
            junk[j] = bytes(rng.integers(0, 256, len(parts[j]), dtype=np.uint8))
            cases.append(junk)
        for ci, bad in enumerate(cases):
            got, want = H.Spectral(src.size, fac, process=2), O.Spectral.create(src.size, fac, progressive=True)
            for b_, bi_, cs_, d_, a_, pa_, iv_ in scans[:6]:
                got.decode_scan(b_, bi_, [(c, 0, 0) for c in cs_], _to_lib_tables(H, d_), _to_lib_tables(H, a_), pa_, iv_)
                want.decode_scan(b_, bi_, cs_, [0] * len(cs_), [0] * len(cs_), d_, a_, pa_, interval=O.INTERVAL_NONE if iv_ is None else iv_)
            try:
                got.decode_scan(band, bits, [(c, 0, 0) for c in comps], _to_lib_tables(H, dct), _to_lib_tables(H, act), bad, ival)
                code = 0
            except lib.JpegSm100Error as e:
                code = e.code
            try:
                want.decode_scan(band, bits, comps, [0], [0], dct, act, bad, interval=O.INTERVAL_NONE if ival is None else ival)
                wcode = 0
            except O.OracleError as e:
                wcode = e.code
            assert code == wcode, (ci, code, wcode)
            if code == 0:
                for p in range(3):
                    assert np.array_equal(got.planes[p].coef, want.coefficients(p)), (ci, p)


_oracle_on_virtual_grid = J.oracle_on_virtual_grid


@pytest.mark.parametrize("name,progression", [("baseline", BASELINE), ("split", BASELINE_SPLIT), ("progressive", PROGRESSIVE)])
def test_restart_intervals_that_are_not_whole_rows(H, O, name, progression):
    """N2: interval e starts at MCU e * Ri wherever that falls in its row (ITU-T T.81 E.1.4; JPEG_SM100_SCAN_T81 on decode, any
    interval_mcus on encode).  The reference cannot be the oracle here (it places intervals by rows, decode.swift:3205-3207, and
    never writes DRI), so: the GPU encoder's bytes and tables equal the oracle's on the virtual grid of the same MCUs, and the
    GPU decoder returns the coefficients the scans were made from -- all five scan kinds, partial MCUs, intervals of 1 MCU,
    primes, just over a row, and longer than the scan."""
    from jpeg_b200 import lib
    rng = np.random.default_rng(23)
    odd = O.Spectral.create((131, 77), [(2, 2), (1, 1), (1, 1)], progressive=True)  # 17 x 10 luma blocks on a 9 x 5 MCU grid
    for p in range(3):
        c = odd.coefficients(p)
        c[...] = np.where(rng.random(c.shape) < 0.2, rng.integers(-30, 31, c.shape), 0).astype(np.int16)
        c[..., 0] = rng.integers(-300, 301, c.shape[:2])
    for src in (O.Spectral.decompress(golden_bytes("gold/color-progressive-1.jpg")), odd):
        _t81_roundtrip(H, O, lib, src, progression)


def test_whole_file_with_odd_restart_interval(H, O):
    """compress(interval_mcus = 7 and 33) -> DRI + RSTn every 7 / 33 MCUs -> decompress(t81_intervals=True), host lexer and GPU
    lexer, host planes and resident image: the coefficients and the pixels of the file without restart intervals."""
    data = golden_bytes("gold/color-sequential-1.jpg")
    plain = H.Spectral.decompress(data)
    rgb = plain.to_rgb8()
    for ival in (7, 33):
        assert ival % (-(-plain.size[0] // 16)) != 0
        again = plain.compress(interval_mcus=ival)
        assert sum(again.count(bytes([0xff, 0xd0 + k])) for k in range(8)) > 10
        for kw in (dict(), dict(gpu_lexer=True), dict(resident=True), dict(resident=True, gpu_lexer=True)):
            back = H.Spectral.decompress(again, t81_intervals=True, **kw)
            for p in range(plain.ncomp):
                assert np.array_equal(back.planes[p].coef, plain.planes[p].coef), (ival, kw, p)
            assert np.array_equal(back.to_rgb8(), rgb), (ival, kw)


def _t81_roundtrip(H, O, lib, src, progression):
    fac = [src.factor(p) for p in range(3)]
    dev = H.Spectral(src.size, fac, process=2)
    for p in range(3):
        dev.planes[p].coef = np.array(src.coefficients(p))
    for k, ival_of in enumerate((lambda W, M: 7, lambda W, M: W + 3, lambda W, M: 1 if M < 700 else 101, lambda W, M: M + 5)):
        dst = H.Spectral(src.size, fac, process=2)
        for band, bits, comps in progression:
            W, Hh = src.blocks if len(comps) > 1 else src.units(comps[0])
            ival = ival_of(W, W * Hh)
            assert ival % W != 0
            sel = [0, 1, 1][:len(comps)] if len(comps) > 1 else [0]
            v = _oracle_on_virtual_grid(O, src, comps, ival)
            want, dct, act = v.encode_scan(band, bits, list(range(len(comps))), sel, sel, ival)
            got, gdc, gac = dev.encode_scan(band, bits, [(c, d, d) for c, d in zip(comps, sel)], ival)
            for t in range(4):
                if dct[t].present:
                    assert gdc[t].as_tuple() == dct[t].as_tuple(), (k, band, bits, comps)
                if act[t].present:
                    assert gac[t].as_tuple() == act[t].as_tuple(), (k, band, bits, comps)
            assert got == want, (k, band, bits, comps, ival, len(got), len(want))
            parts = J.unstuff_split(got)
            assert len(parts) == -(-W * Hh // ival)
            dst.decode_scan(band, bits, [(c, d, d) for c, d in zip(comps, sel)], _to_lib_tables(H, dct), _to_lib_tables(H, act), parts, ival,
                            extend=lib.SCAN_T81)
        for p in range(3):
            assert np.array_equal(dst.planes[p].coef, src.coefficients(p)), (k, p)


# ------------------------------------------------------------------------------------------------ layer B: batches
def test_batched_device_pipeline(H, O, manifest):
    """Layer B on a small batch of different images with identical geometry and per-image tables:
    dev_decode_scan -> dev_idct (8-bit) -> dev_planar_to_rgb8, all on device memory."""
    import ctypes as C

    import torch

    from jpeg_b200 import lib
    eb = manifest["encode_basic"]
    w0, h0 = eb["size"]
    full = np.frombuffer(golden_bytes(eb["rgb"]), dtype=np.uint8).reshape(h0, w0, 3)
    W, Hh, N = 136, 104, 5
    factors = [(2, 2), (1, 1), (1, 1)]
    q = [O.quanta(0.25, 0), O.quanta(0.25, 1), O.quanta(0.25, 1)]
    specs, all_parts, tables, want = [], [], [], []
    for i in range(N):
        rgb = np.ascontiguousarray(full[40 * i:40 * i + Hh, 30 * i:30 * i + W])
        planes = O.decompose(O.pack_rgb(rgb), factors)
        s = O.Spectral.create((W, Hh), factors)
        for p in range(3):
            s.coefficients(p)[...] = O.fdct_plane(planes[p], q[p])
            s.set_quanta(p, q[p])
        ecs, dct, act = s.encode_scan((0, 64), (0, None), [0, 1, 2], [0, 1, 1], [0, 1, 1], s.blocks[0])
        all_parts.append(J.unstuff_split(ecs))
        tables += _to_lib_tables(H, dct) + _to_lib_tables(H, act)
        specs.append(s)
        want.append(O.unpack_rgb(s.to_rectangular()))
    n_ecs = len(all_parts[0])
    assert all(len(p) == n_ecs for p in all_parts)
    cat = b"".join(b"".join(p) for p in all_parts)
    offs = np.zeros(N * n_ecs + 1, dtype=np.uint64)
    np.cumsum([len(e) for p in all_parts for e in p], out=offs[1:])
    dev = torch.device("cuda:0")
    d_ecs = torch.from_numpy(np.frombuffer(cat + b"\0" * 16, dtype=np.uint8).copy()).to(dev)
    d_off = torch.from_numpy(offs.view(np.int64)).to(dev)
    ctx = lib.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    s0 = specs[0]
    sp, pl = lib.DevSpectral(), lib.DevPlanar()
    sp.n_images = pl.n_images = N
    sp.n_planes = pl.n_planes = 3
    pl.sample_bytes = 1
    coefs, samples = [], []
    for p in range(3):
        ux, uy = s0.units(p)
        c = torch.zeros((N, uy, ux, 64), dtype=torch.int16, device=dev)
        sm = torch.zeros((N, 8 * uy, 8 * ux), dtype=torch.uint8, device=dev)
        coefs.append(c)
        samples.append(sm)
        for dst, ptr, stride in ((sp.plane[p], c, 64 * ux * uy), (pl.plane[p], sm, 64 * ux * uy)):
            dst.image_stride = stride
            dst.units_x, dst.units_y = ux, uy
            dst.factor_x, dst.factor_y = factors[p]
        sp.plane[p].coef = c.data_ptr()
        pl.plane[p].samples = sm.data_ptr()
    desc = lib.ScanDesc()
    desc.band_lo, desc.band_hi, desc.bit_lo, desc.bit_hi, desc.n_comp = 0, 64, 0, -1, 3
    for i, (dcs, acs) in enumerate(((0, 0), (1, 1), (1, 1))):
        desc.comp[i].plane = i
        desc.comp[i].factor_x, desc.comp[i].factor_y = factors[i]
        desc.comp[i].dc, desc.comp[i].ac = dcs, acs
    desc.blocks_x, desc.blocks_y = s0.blocks
    tabs = (lib.HuffTable * (8 * N))(*[t if t is not None else lib.HuffTable() for t in tables])
    status = torch.zeros(N, dtype=torch.int32, device=dev)
    ctx.check(ctx.L.jpeg_sm100_dev_decode_scan(ctx.h, C.byref(desc), d_ecs.data_ptr(), d_off.data_ptr(), n_ecs,
                                               s0.blocks[0], 0, tabs, 0, C.byref(sp), status.data_ptr()))
    qz = np.ascontiguousarray(np.stack(q), dtype=np.uint16)
    ctx.check(ctx.L.jpeg_sm100_dev_idct(ctx.h, C.byref(sp), qz.ctypes.data, 8, C.byref(pl)))
    d_rgb = torch.zeros((N, Hh, W, 3), dtype=torch.uint8, device=dev)
    ctx.check(ctx.L.jpeg_sm100_dev_planar_to_rgb8(ctx.h, C.byref(pl), W, Hh, 0, d_rgb.data_ptr()))
    torch.cuda.synchronize()
    assert status.cpu().tolist() == [0] * N
    for i in range(N):
        for p in range(3):
            assert np.array_equal(coefs[p][i].cpu().numpy(), specs[i].coefficients(p)), (i, p)
        assert np.array_equal(d_rgb[i].cpu().numpy(), want[i]), i
    assert ctx.launches >= 2 + 3 + 1  # decode + status reduce, one IDCT per plane, colour


# ------------------------------------------------------------------------------------------------ entropy encode (K6/K7)
@pytest.mark.parametrize("name,progression", [("baseline", BASELINE), ("split", BASELINE_SPLIT),
                                              ("progressive", PROGRESSIVE)])
@pytest.mark.parametrize("rows", [0, 1, 3])
def test_scan_encode_all_kinds(H, O, name, progression, rows):
    """GPU Spectral.encode(scan:) against the oracle: optimal tables and stuffed ECS bytes identical, for the five
    scan kinds, with the reference's single-ECS form (rows = 0) and our RSTn extension."""
    data = golden_bytes("gold/color-progressive-1.jpg")
    src = O.Spectral.decompress(data)
    dev = H.Spectral.decompress(data)
    for band, bits, comps in progression:
        width = src.blocks[0] if len(comps) > 1 else src.units(comps[0])[0]
        sel = [0, 1, 1][:len(comps)] if len(comps) > 1 else [0]
        want, dct, act = src.encode_scan(band, bits, comps, sel, sel, rows * width)
        got, gdc, gac = dev.encode_scan(band, bits, [(c, d, d) for c, d in zip(comps, sel)], rows * width)
        for k in range(4):
            assert bool(gdc[k].present) == bool(dct[k].present) and bool(gac[k].present) == bool(act[k].present)
            if dct[k].present:
                assert gdc[k].as_tuple() == dct[k].as_tuple(), (name, band, bits, "dc", k)
            if act[k].present:
                assert gac[k].as_tuple() == act[k].as_tuple(), (name, band, bits, "ac", k)
        assert got == want, (name, rows, band, bits, comps, len(got), len(want))


@pytest.mark.parametrize("mode", ["blocks", "seq"])
@pytest.mark.parametrize("rows", [0, 2, 40])
def test_progressive_ac_encoders_block_parallel(H, O, monkeypatch, mode, rows):
    """kinds 3 and 4 through k_enc_acp_flags / k_enc_acp_chain + the per-block passes (and the one-thread-per-interval kernels
    forced on the same input): two refinement passes over split bands on a plane of 16384 blocks that is mostly empty, so that
    end-of-band runs exceed the 4096-block limit several times over, with dense stretches, blocks whose last coefficient is
    non-zero (no EOB) and blocks that only hold correction bits -- bytes and tables identical to the oracle's."""
    if mode == "seq":
        monkeypatch.setenv("JPEG_SM100_ENC_AC", "seq")
    rng = np.random.default_rng(17)
    size, fac = (2048, 512), [(1, 1), (1, 1), (1, 1)]
    src = O.Spectral.create(size, fac, progressive=True)
    for p in range(3):
        c = src.coefficients(p)
        c[...] = np.where(rng.random(c.shape) < 0.03, rng.integers(-12, 13, c.shape), 0).astype(np.int16)
        c[rng.random(c.shape[:2]) < (0.97 if p == 0 else 0.6), 1:] = 0   # plane 0: runs of thousands of empty blocks
        c[3, 10:60, 1:] = rng.integers(-5, 6, (50, 63))                    # a dense stretch
        c[5, 7, 63] = 9                                                    # last coefficient set: the block needs no EOB
        c[7, :, 1:] = 0
        c[7, ::3, 5] = 6                                                   # significant since the first pass: correction bits only
    dev = H.Spectral(size, fac, process=2)
    for p in range(3):
        dev.planes[p].coef = np.array(src.coefficients(p))
    for band, bits, comps in DEEP_REFINEMENT[1:]:
        width = src.units(comps[0])[0]
        want, dct, act = src.encode_scan(band, bits, comps, [0], [0], rows * width)
        got, gdc, gac = dev.encode_scan(band, bits, [(comps[0], 0, 0)], rows * width)
        assert gac[0].as_tuple() == act[0].as_tuple(), (band, bits, comps)
        assert got == want, (mode, rows, band, bits, comps, len(got), len(want))


def test_encode_basic_golden_through_gpu(manifest, H):
    """examples/encode-basic end to end on the GPU: RGB -> pack -> decomposed -> fdct -> encode(scan:) must give the
    reference's own DHT tables and entropy-coded bytes (digests from the 32 committed JPEG files)."""
    from jpeg_b200 import lib
    eb = manifest["encode_basic"]
    w, h = eb["size"]
    rgb = np.frombuffer(golden_bytes(eb["rgb"]), dtype=np.uint8).reshape(h, w, 3)
    levels = {"0.0": 0.0, "0.25": 0.25, "1.0": 1.0, "8.0": 8.0}
    for name, lum in (("4-4-4", (1, 1)), ("4-4-0", (1, 2)), ("4-2-2", (2, 1)), ("4-2-0", (2, 2))):
        factors = [lum, (1, 1), (1, 1)]
        planar = H.Rectangular.pack(rgb, factors).decomposed()
        for tag, level in levels.items():
            exp = eb["files"][f"{name}-{tag}"]
            q = [np.array(exp["dqt"][0][1], np.uint16), np.array(exp["dqt"][1][1], np.uint16)]
            sp = planar.fdct([q[0], q[1], q[1]])
            for sc, comps in zip(exp["scans"], ([(0, 0, 0)], [(1, 1, 1), (2, 1, 1)])):
                ecs, dct, act = sp.encode_scan((0, 64), (0, None), comps)
                assert len(ecs) == sc["ecs_len"] and sha(ecs) == sc["ecs_sha256"], (name, tag)
                for cls, tgt, counts, values in sc["dht"]:
                    tab = (dct if cls == 0 else act)[tgt]
                    assert tab.as_tuple() == (bytes.fromhex(counts), bytes.fromhex(values)), (name, tag)


def test_reencode_in_memory_through_gpu(manifest, H):
    """examples/in-memory: the 14-scan progressive file decoded to coefficients on the GPU and re-encoded from them."""
    exp = manifest["reencode"]["in-memory"]
    s = H.Spectral.decompress(golden_bytes(exp["source"]))
    ids = [p.comp_id for p in s.planes]
    for k, sc in enumerate(exp["scans"]):
        sos = bytes.fromhex(sc["sos"])
        n = sos[0]
        comps = [(ids.index(sos[1 + 2 * i]), sos[2 + 2 * i] >> 4, sos[2 + 2 * i] & 15) for i in range(n)]
        band = (sos[2 * n + 1], sos[2 * n + 2] + 1)
        al, ah = sos[2 * n + 3] & 15, sos[2 * n + 3] >> 4
        ecs, dct, act = s.encode_scan(band, (al, None if ah == 0 else ah), comps)
        assert len(ecs) == sc["ecs_len"] and sha(ecs) == sc["ecs_sha256"], k
        for cls, tgt, counts, values in sc["dht"]:
            tab = (dct if cls == 0 else act)[tgt]
            assert tab.as_tuple() == (bytes.fromhex(counts), bytes.fromhex(values)), k


def test_compress_decompress_roundtrip_on_gpu(H, O):
    """Spectral.compress() with DRI -> Spectral.decompress(): coefficients survive; the oracle decodes our file to
    the same pixels (the file is a valid JPEG for an independent decoder)."""
    src = H.Spectral.decompress(golden_bytes("gold/color-sequential-2.jpg"))
    blob = src.compress(scans=[H.Scan((0, 64), (0, None), [(0, 0, 0), (1, 1, 1), (2, 1, 1)])],
                        interval_mcus=src.blocks[0])
    back = H.Spectral.decompress(blob)
    for a, b in zip(src.planes, back.planes):
        assert np.array_equal(a.coef, b.coef)
    rgb, _, _ = O.decode_rgb(blob)
    assert np.array_equal(rgb, back.to_rgb8())


# ------------------------------------------------------------------------------------------------ N3: spectral-domain operations
@pytest.mark.parametrize("kind", ["ii", "iii", "iv"])
def test_rotate_through_gpu(manifest, H, O, kind):
    """examples/rotate: the block-transform kernel against the oracle (itself pinned to the reference's three rotated files):
    coefficients, permuted quantisation tables, and the re-encoded scans byte for byte."""
    data = golden_bytes(manifest["rotate"]["source"])
    got = H.Spectral.decompress(data).rotated(kind)
    want = O.rotated(O.Spectral.decompress(data), kind)
    assert got.size == want.size
    for p in range(3):
        assert np.array_equal(got.planes[p].coef, want.coefficients(p)), (kind, p)
        assert np.array_equal(got.quanta[got.planes[p].q], want.quanta(p)), (kind, p)
    exp = manifest["rotate"]["outputs"][kind]
    for k, sc in enumerate(exp["scans"]):
        sos = bytes.fromhex(sc["sos"])
        n = sos[0]
        ids = [sos[1 + 2 * i] for i in range(n)]
        cid = [pl.comp_id for pl in got.planes]
        comps = [(cid.index(ids[i]), sos[2 + 2 * i] >> 4, sos[2 + 2 * i] & 15) for i in range(n)]
        band = (sos[2 * n + 1], sos[2 * n + 2] + 1)
        al, ah = sos[2 * n + 3] & 15, sos[2 * n + 3] >> 4
        ecs, _, _ = got.encode_scan(band, (al, None if ah == 0 else ah), comps)
        assert len(ecs) == sc["ecs_len"] and sha(ecs) == sc["ecs_sha256"], (kind, k)


def test_requantize_through_gpu(manifest, H, O):
    """examples/recompress: the requantisation kernel against the oracle's restatement (pinned to recompressed-requantized.jpg)."""
    data = golden_bytes(manifest["reencode"]["recompress-requantized"]["source"])
    s = H.Spectral.decompress(data)
    ref = O.Spectral.decompress(data)
    new_q = [np.concatenate([q[:1], np.minimum(q[1:].astype(np.int64) * 3, 255)]).astype(np.uint16) for q in s.quanta]
    got = s.requantized(new_q)
    for p in range(3):
        q = ref.quanta(p)
        nq = np.concatenate([q[:1], np.minimum(q[1:].astype(np.int64) * 3, 255)]).astype(np.uint16)
        assert np.array_equal(got.planes[p].coef, O.requantize(ref.coefficients(p), q, nq)), p
    # random coefficients and tables, including negative values and quotients that land on the rounding boundary
    rng = np.random.default_rng(3)
    coef = rng.integers(-1024, 1024, (5, 7, 64), dtype=np.int16)
    qo = rng.integers(1, 31, 64).astype(np.uint16)
    qn = rng.integers(1, 256, 64).astype(np.uint16)
    t = H.Spectral((56, 40), [(1, 1)])
    t.planes[0].coef = coef
    t.quanta = [qo]
    assert np.array_equal(t.requantized([qn]).planes[0].coef, O.requantize(coef, qo, qn))


def test_spectral_ops_on_device_batches(H, O):
    """layer B of N3: a batch of two images, planes with an image stride, requantise in place, then a rotation."""
    import ctypes as C

    import torch
    from jpeg_b200 import batch, lib
    ctx, dev = H.default_context(), torch.device("cuda:0")
    geo = batch.Geometry((200, 128), [(2, 2), (1, 1), (1, 1)])
    src = batch.DeviceBuffers(geo, 2, dev)
    g = torch.Generator(device=dev)
    g.manual_seed(9)
    for c in src.coef:
        c.copy_(torch.randint(-300, 300, c.shape, generator=g, device=dev, dtype=torch.int16))
    host = [c.cpu().numpy().copy() for c in src.coef]
    qo = np.stack([np.arange(1, 65), np.arange(2, 66), np.arange(3, 67)]).astype(np.uint16)
    qn = np.stack([np.arange(64, 0, -1), np.full(64, 7), np.full(64, 255)]).astype(np.uint16)
    ctx.check(ctx.L.jpeg_sm100_dev_requantize(ctx.h, C.byref(src.sp), qo.ctypes.data, qn.ctypes.data, C.byref(src.sp)))
    torch.cuda.synchronize()
    req = [np.stack([O.requantize(host[p][i], qo[p], qn[p]) for i in range(2)]) for p in range(3)]
    for p in range(3):
        assert np.array_equal(src.coef[p].cpu().numpy(), req[p]), p
    zmap, mul, matrix = H.Spectral.block_mapping("iv")
    rot = batch.Geometry((128, 200), [(2, 2), (1, 1), (1, 1)])  # (the height is a whole number of MCUs: no crop)
    dst = batch.DeviceBuffers(rot, 2, dev)
    for c in dst.coef:
        c.fill_(77)
    ctx.check(ctx.L.jpeg_sm100_dev_transform_blocks(ctx.h, C.byref(src.sp), matrix.ctypes.data, zmap.ctypes.data, mul.ctypes.data,
                                                    C.byref(dst.sp)))
    torch.cuda.synchronize()
    m = ((int(matrix[0]), int(matrix[1])), (int(matrix[2]), int(matrix[3])))
    for p in range(3):
        for i in range(2):
            want = O.transform_blocks(req[p][i], m, zmap, mul, rot.units[p])
            assert np.array_equal(dst.coef[p][i].cpu().numpy(), want), (p, i)
