"""N1: the GPU lexer (jpeg_b200/csrc/lexer.cu) against Bytestream.segment(prefix: true) + the scan loop of
Context.decompress (decode.swift:130-190, 3895-3933) as restated by tests/jpegfile.py::unstuff_split and the oracle."""
import ctypes as C

import numpy as np
import pytest

import jpegfile as J
from conftest import golden_bytes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    assert torch.cuda.is_available()
    from jpeg_b200 import batch, host, lib, synth
    from oracle import oracle as O
    ctx = lib.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    return dict(torch=torch, lib=lib, ctx=ctx, host=host, batch=batch, synth=synth, O=O, dev=torch.device("cuda:0"))


def make_raw(rng, n_segments, seg_len, ff_rate=0.05, fill=True, wrong_phase_at=None):
    """random entropy-coded data with stuffed FFs, optional fill bytes before the RSTn markers"""
    out, parts = bytearray(), []
    for k in range(n_segments):
        n = int(rng.integers(0, seg_len + 1))
        data = rng.integers(0, 256, n, dtype=np.uint8)
        data[rng.random(n) < ff_rate] = 0xFF
        parts.append(data.tobytes())
        for b in data.tobytes():
            out.append(b)
            if b == 0xFF:
                out.append(0)
        if k + 1 < n_segments:
            if fill and rng.random() < 0.3:
                out += b"\xff" * int(rng.integers(1, 4))
            phase = k % 8 if wrong_phase_at != k else (k + 3) % 8
            out += bytes([0xFF, 0xD0 + phase])
    return bytes(out), parts


def lex_host_buffers(env, raw):
    ctx, L = env["ctx"], env["ctx"].L
    buf = np.frombuffer(raw + b"\0" * 32, dtype=np.uint8)
    ecs = np.zeros(len(raw) + 64, np.uint8)
    offs = np.zeros(len(raw) // 2 + 8, np.uint64)
    n = C.c_uint32()
    rc = L.jpeg_sm100_lex_scan(ctx.h, buf.ctypes.data, len(raw), ecs.ctypes.data, ecs.size, offs.ctypes.data, offs.size, C.byref(n))
    if rc:
        return rc, None
    o = offs[:n.value + 1].astype(np.int64)
    return 0, [ecs[o[k]:o[k + 1]].tobytes() for k in range(n.value)]


@pytest.mark.parametrize("n_segments,seg_len", [(1, 0), (1, 5), (1, 5000), (9, 700), (40, 3000), (300, 40), (2, 70000)])
def test_lex_scan_random_streams(env, n_segments, seg_len):
    rng = np.random.default_rng(n_segments * 1000 + seg_len)
    for trial in range(3):
        raw, parts = make_raw(rng, n_segments, seg_len, ff_rate=[0.01, 0.2, 0.9][trial])
        assert J.unstuff_split(raw) == parts  # the host restatement agrees with the construction
        rc, got = lex_host_buffers(env, raw)
        assert rc == 0
        assert got == parts


def test_lex_scan_errors(env):
    lib = env["lib"]
    rng = np.random.default_rng(5)
    raw, _ = make_raw(rng, 12, 300, wrong_phase_at=9)
    assert lex_host_buffers(env, raw)[0] == lib.ERR_RESTART_PHASE
    raw, _ = make_raw(rng, 3, 300)
    assert lex_host_buffers(env, raw[:100] + b"\xff\xc4" + raw[100:])[0] == lib.ERR_INVALID_ARGUMENT
    # a trailing FF (start of the next marker, or fill) emits nothing
    rc, got = lex_host_buffers(env, b"\x12\x34\xff")
    assert rc == 0 and got == [b"\x12\x34"]
    rc, got = lex_host_buffers(env, b"")
    assert rc == 0 and got == [b""]


def test_dev_lex_scan_batch_unaligned_images(env):
    """several images at arbitrary byte offsets of one device buffer, each with the same number of segments"""
    t, ctx, lib, dev = env["torch"], env["ctx"], env["lib"], env["dev"]
    rng = np.random.default_rng(11)
    n_images, n_ecs = 7, 13
    blobs, parts = zip(*[make_raw(rng, n_ecs, 2000, ff_rate=0.1) for _ in range(n_images)])
    offs, cat = [], bytearray()
    for b in blobs:
        cat += bytes(int(rng.integers(0, 23)))  # garbage gap, any alignment
        offs.append(len(cat))
        cat += b
    cat += bytes(32)
    d_raw = t.from_numpy(np.frombuffer(bytes(cat), dtype=np.uint8).copy()).to(dev)
    d_ecs = t.zeros(len(cat) + 64, dtype=t.uint8, device=dev)
    d_off = t.zeros(n_images * n_ecs + 1, dtype=t.int64, device=dev)
    d_st = t.full((n_images,), 99, dtype=t.int32, device=dev)
    ro = np.array(offs, np.uint64)
    rl = np.array([len(b) for b in blobs], np.uint64)
    ctx.check(ctx.L.jpeg_sm100_dev_lex_scan(ctx.h, d_raw.data_ptr(), ro.ctypes.data, rl.ctypes.data, n_images, n_ecs,
                                            d_ecs.data_ptr(), d_off.data_ptr(), d_st.data_ptr()))
    t.cuda.synchronize()
    assert d_st.tolist() == [0] * n_images
    o = d_off.cpu().numpy()
    e = d_ecs.cpu().numpy()
    assert o[0] == 0
    for i in range(n_images):
        for k in range(n_ecs):
            a, b = o[i * n_ecs + k], o[i * n_ecs + k + 1]
            assert e[a:b].tobytes() == parts[i][k], (i, k)
    # one image with a missing marker: its status is ERR_ECS_COUNT, the others are untouched
    ctx.check(ctx.L.jpeg_sm100_dev_lex_scan(ctx.h, d_raw.data_ptr(), ro.ctypes.data, rl.ctypes.data, n_images, n_ecs + 1,
                                            d_ecs.data_ptr(), d_off.data_ptr(), d_st.data_ptr()))
    t.cuda.synchronize()
    assert d_st.tolist() == [lib.ERR_ECS_COUNT] * n_images


def test_decode_scan_raw_on_restart_files(env, manifest):
    """whole files with restart intervals: the host only finds the end of each scan; lexing + decoding on the GPU"""
    O, H, ctx = env["O"], env["host"], env["ctx"]
    for name in list(manifest["restart"]) + ["gold/color-sequential-2.jpg", "gold/color-progressive-1.jpg"]:
        data = golden_bytes(name)
        ref = O.Spectral.decompress(data)
        s = H.Spectral.decompress(data, gpu_lexer=True)
        for p in range(ref.ncomp):
            assert np.array_equal(s.planes[p].coef, ref.coefficients(p)), (name, p)


def test_decode_scan_raw_errors(env):
    """missingRestartIntervalSegment (decode.swift:3719) and invalidRestartPhase (3931) through the raw entry point"""
    H, lib = env["host"], env["lib"]
    src = H.Spectral.decompress(golden_bytes("gold/color-sequential-2.jpg"))
    blob = src.compress(scans=[H.Scan((0, 64), (0, None), [(0, 0, 0), (1, 1, 1), (2, 1, 1)])], interval_mcus=src.blocks[0])
    segs = J.split(blob)
    # drop the DRI segment: RSTn markers without a restart interval definition
    i = blob.index(b"\xff\xdd")
    with pytest.raises(lib.JpegSm100Error) as e:
        H.Spectral.decompress(blob[:i] + blob[i + 6:], gpu_lexer=True)
    assert e.value.code == lib.ERR_MISSING_INTERVAL
    # corrupt the phase of the third marker
    j = blob.index(b"\xff\xda")
    for _ in range(3):
        j = next(k for k in range(j + 1, len(blob) - 1) if blob[k] == 0xFF and 0xD0 <= blob[k + 1] <= 0xD7)
    bad = bytearray(blob)
    bad[j + 1] = 0xD0 + ((bad[j + 1] - 0xD0 + 1) % 8)
    with pytest.raises(lib.JpegSm100Error) as e:
        H.Spectral.decompress(bytes(bad), gpu_lexer=True)
    assert e.value.code == lib.ERR_RESTART_PHASE
    assert segs[0][0] == 0xD8


def test_batch_raw_equals_batch_prelexed_4k(env):
    """config #2 at full size: raw scan bytes in -> RGB out, against the pre-lexed entry point and the host lexer"""
    t, b, lib, O, ctx, dev = env["torch"], env["batch"], env["lib"], env["O"], env["ctx"], env["dev"]
    W, H, N = 3840, 2160, 3
    geo = b.Geometry((W, H), [(2, 2), (1, 1), (1, 1)])
    q = np.stack([O.quanta(0.25, 0), O.quanta(0.25, 1), O.quanta(0.25, 1)])
    frames = t.stack([env["synth"].frame(300 + i, W, H, dev) for i in range(N)])
    ecs, tabs, _ = b.encode_frames(ctx, frames, geo, q, geo.blocks[0])
    inputs = b.DecodeInputs(ecs, list(tabs), n_ecs_expected=135)
    desc = b.sequential_scan(geo)
    tarr = (lib.HuffTable * (8 * N))(*list(tabs))
    want = np.zeros((N, H, W, 3), np.uint8)
    st = np.zeros(N, np.int32)
    ctx.check(ctx.L.jpeg_sm100_decode_batch_rgb8(ctx.h, C.byref(desc), N, inputs.ecs.ctypes.data, inputs.offsets.ctypes.data,
                                                 inputs.n_ecs, geo.blocks[0], tarr, 0, q.ctypes.data, W, H, 0,
                                                 want.ctypes.data, st.ctypes.data))
    raw_off, cat = [], bytearray()
    for e in ecs:
        cat += bytes(len(cat) % 7)  # odd alignments
        raw_off.append(len(cat))
        cat += e.tobytes()
    raw = np.frombuffer(bytes(cat) + bytes(64), np.uint8)
    ro, rl = np.array(raw_off, np.uint64), np.array([len(e) for e in ecs], np.uint64)
    got = np.zeros_like(want)
    ctx.check(ctx.L.jpeg_sm100_decode_batch_raw_rgb8(ctx.h, C.byref(desc), N, raw.ctypes.data, ro.ctypes.data, rl.ctypes.data,
                                                     inputs.n_ecs, geo.blocks[0], tarr, 0, q.ctypes.data, W, H, 0,
                                                     got.ctypes.data, st.ctypes.data))
    assert st.tolist() == [0] * N
    assert np.array_equal(got, want)


@pytest.mark.parametrize("size,factors,n", [((200, 120), [(2, 2), (1, 1), (1, 1)], 37), ((97, 61), [(1, 1)], 19),
                                             ((136, 104), [(2, 1), (1, 1), (1, 1)], 9)])
def test_batch_entry_points_many_small_images(env, size, factors, n):
    """the host-buffer batch entry points with more images than one group of their internal pipeline (groups of >= 8 images go
    through upload -> lexer -> K3 -> K1/K2 -> download one after the other), per-image tables, colour and greyscale, odd sizes:
    RGB equals the oracle's for every image; a corrupted image reports its own error and leaves the others intact."""
    t, b, lib, O, ctx = env["torch"], env["batch"], env["lib"], env["O"], env["ctx"]
    W, H = size
    rng = np.random.default_rng(n)
    geo = b.Geometry(size, factors)
    npl = len(factors)
    q = [O.quanta(0.5, 0)] + [O.quanta(0.5, 1)] * (npl - 1)
    ecs_list, tabs, want = [], [], []
    comps = list(range(npl))
    sel = [0, 1, 1][:npl]
    ival = geo.blocks[0] if npl > 1 else geo.units[0][0]
    for i in range(n):
        s = O.Spectral.create(size, factors)
        for p in range(npl):
            c = s.coefficients(p)
            c[...] = np.where(rng.random(c.shape) < 0.1, rng.integers(-40, 40, c.shape), 0).astype(np.int16)
            c[..., 0] = rng.integers(-100, 100, c.shape[:2])
            s.set_quanta(p, q[p])
        ecs, dct, act = s.encode_scan((0, 64), (0, None), comps, sel, sel, ival)
        ecs_list.append(np.frombuffer(ecs, np.uint8))
        mk = lambda x: lib.HuffTable.make(bytes(x.counts), bytes(x.values)) if x.present else lib.HuffTable()
        tabs += [mk(x) for x in dct] + [mk(x) for x in act]
        want.append(O.unpack_rgb(s.to_rectangular()))
    want = np.stack(want)
    desc = b.sequential_scan(geo, dc=sel, ac=sel)
    tarr = (lib.HuffTable * (8 * n))(*tabs)
    qz = np.ascontiguousarray(np.stack(q), dtype=np.uint16)
    # raw scan bytes at odd offsets
    raw_off, cat = [], bytearray()
    for e in ecs_list:
        cat += bytes(len(cat) % 5)
        raw_off.append(len(cat))
        cat += e.tobytes()
    raw = np.frombuffer(bytes(cat) + bytes(64), np.uint8)
    ro, rl = np.array(raw_off, np.uint64), np.array([len(e) for e in ecs_list], np.uint64)
    n_ecs = geo.blocks[1] if npl > 1 else geo.units[0][1]
    got, st = np.zeros_like(want), np.zeros(n, np.int32)
    ctx.check(ctx.L.jpeg_sm100_decode_batch_raw_rgb8(ctx.h, C.byref(desc), n, raw.ctypes.data, ro.ctypes.data, rl.ctypes.data, n_ecs, ival,
                                                     tarr, 0, qz.ctypes.data, W, H, 0, got.ctypes.data, st.ctypes.data))
    assert st.tolist() == [0] * n
    assert np.array_equal(got, want)
    # the pre-lexed entry point on the same images
    inputs = b.DecodeInputs(ecs_list, tabs, n_ecs_expected=n_ecs)
    got2, st2 = np.zeros_like(want), np.zeros(n, np.int32)
    ctx.check(ctx.L.jpeg_sm100_decode_batch_rgb8(ctx.h, C.byref(desc), n, inputs.ecs.ctypes.data, inputs.offsets.ctypes.data, n_ecs, ival,
                                                 tarr, 0, qz.ctypes.data, W, H, 0, got2.ctypes.data, st2.ctypes.data))
    assert st2.tolist() == [0] * n and np.array_equal(got2, want)
    # image n // 2 cut short inside its second interval: its status is the truncation error, every other image is unchanged
    bad = n // 2
    offs = inputs.offsets.copy()
    cut = int(offs[bad * n_ecs + 2] - offs[bad * n_ecs + 1]) // 2 + 1
    ecs_cut = np.concatenate([inputs.ecs[:int(offs[bad * n_ecs + 2]) - cut], inputs.ecs[int(offs[bad * n_ecs + 2]):]])
    offs[bad * n_ecs + 2:] -= np.uint64(cut)
    got3, st3 = np.zeros_like(want), np.zeros(n, np.int32)
    rc = ctx.L.jpeg_sm100_decode_batch_rgb8(ctx.h, C.byref(desc), n, ecs_cut.ctypes.data, offs.ctypes.data, n_ecs, ival, tarr, 0,
                                            qz.ctypes.data, W, H, 0, got3.ctypes.data, st3.ctypes.data)
    assert rc == lib.ERR_TRUNCATED_ECS
    assert st3.tolist() == [lib.ERR_TRUNCATED_ECS if i == bad else 0 for i in range(n)]
    keep = [i for i in range(n) if i != bad]
    assert np.array_equal(got3[keep], want[keep])
