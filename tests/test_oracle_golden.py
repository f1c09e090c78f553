"""Pins the CPU oracle (oracle/jpeg_oracle.c) to the reference's own golden vectors -- bit-exact.

Mirrors tests/regression/tests.swift:39-138 (decode -> .ycc / .rgb byte compare), tests/unit/tests.swift
(zig-zag, amplitude coding, Annex-K Huffman KAT, invalid codewords, Huffman round trip) and uses the committed
outputs of examples/{decode-basic,decode-advanced,in-memory,encode-basic,recompress,encode-advanced,rotate,custom-color}
as vectors.
"""
import hashlib
import os

import numpy as np
import pytest

import jpegfile as J
from conftest import GOLDEN, REFERENCE, golden_bytes, have_reference
from oracle import oracle as O


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


# ---------------------------------------------------------------- unit KATs (tests/unit/tests.swift)

def test_zigzag_table():
    expected = [[0, 1, 5, 6, 14, 15, 27, 28], [2, 4, 7, 13, 16, 26, 29, 42], [3, 8, 12, 17, 25, 30, 41, 43],
                [9, 11, 18, 24, 31, 40, 44, 53], [10, 19, 23, 32, 39, 45, 52, 54], [20, 22, 33, 38, 46, 51, 55, 60],
                [21, 34, 37, 47, 50, 56, 59, 61], [35, 36, 48, 49, 57, 58, 62, 63]]
    assert O.zigzag_table().tolist() == expected  # tests.swift:38-48


def test_amplitude_extend_kat():
    vec = [(1, 0, -1), (1, 1, 1), (2, 0, -3), (2, 1, -2), (2, 2, 2), (2, 3, 3), (5, 0, -31), (5, 1, -30),
           (5, 14, -17), (5, 15, -16), (5, 16, 16), (5, 17, 17), (5, 30, 30), (5, 31, 31), (11, 0, -2047),
           (11, 1, -2046), (11, 1023, -1024), (11, 1024, 1024), (11, 2046, 2046), (11, 2047, 2047),
           (15, 0, -32767), (15, 1, -32766), (15, 16383, -16384), (15, 16384, 16384), (15, 32766, 32766),
           (15, 32767, 32767)]  # tests.swift:72-104
    L = O.lib()
    for binade, tail, exp in vec:
        assert L.orc_extend(binade, tail) == exp


def test_amplitude_roundtrip_full_range():
    import ctypes as C
    L = O.lib()
    b, t = C.c_int(), C.c_uint()
    for x in list(range(-32767, 0)) + list(range(1, 32768)):
        L.orc_compact(x, C.byref(b), C.byref(t))
        assert L.orc_extend(b.value, t.value) == x
        assert b.value == abs(x).bit_length()


ANNEX_K_AC_COUNTS = [0x00, 0x02, 0x01, 0x03, 0x03, 0x02, 0x04, 0x03, 0x05, 0x05, 0x04, 0x04, 0x00, 0x00, 0x01, 0x7D]
ANNEX_K_AC_VALUES = [
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07,
    0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xA1, 0x08, 0x23, 0x42, 0xB1, 0xC1, 0x15, 0x52, 0xD1, 0xF0,
    0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0A, 0x16, 0x17, 0x18, 0x19, 0x1A, 0x25, 0x26, 0x27, 0x28,
    0x29, 0x2A, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3A, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49,
    0x4A, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5A, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69,
    0x6A, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7A, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89,
    0x8A, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9A, 0xA2, 0xA3, 0xA4, 0xA5, 0xA6, 0xA7,
    0xA8, 0xA9, 0xAA, 0xB2, 0xB3, 0xB4, 0xB5, 0xB6, 0xB7, 0xB8, 0xB9, 0xBA, 0xC2, 0xC3, 0xC4, 0xC5,
    0xC6, 0xC7, 0xC8, 0xC9, 0xCA, 0xD2, 0xD3, 0xD4, 0xD5, 0xD6, 0xD7, 0xD8, 0xD9, 0xDA, 0xE1, 0xE2,
    0xE3, 0xE4, 0xE5, 0xE6, 0xE7, 0xE8, 0xE9, 0xEA, 0xF1, 0xF2, 0xF3, 0xF4, 0xF5, 0xF6, 0xF7, 0xF8,
    0xF9, 0xFA]


def test_huffman_annex_k_kat(manifest):
    """tests.swift:141-361: every (length, codeword) of the K.3.3.2 table decodes to the expected run/size."""
    spec = O.HuffSpec.make(ANNEX_K_AC_COUNTS, ANNEX_K_AC_VALUES)
    expected = 0
    for length, codeword in manifest["unit"]["annex_k_ac_codewords"]:
        sym, ln = O.huff_lookup(spec, (codeword << (16 - length)) & 0xFFFF)
        assert (sym, ln) == (expected, length)
        if expected & 0x0F < 0x0A:
            expected = (expected & 0xF0) | ((expected & 0x0F) + 1)
        else:
            expected = (((expected & 0xF0) + 0x10) & 0xFF) | (0 if expected & 0xF0 == 0xE0 else 1)


def _decode_symbols(spec, data):
    bits = "".join(f"{b:08b}" for b in data) + "1" * 32
    out, b = [], 0
    while b < 8 * len(data):
        sym, ln = O.huff_lookup(spec, int(bits[b:b + 16], 2))
        out.append(sym)
        b += ln
    return out


def test_huffman_small_trees_and_invalid_codewords():
    """tests.swift:365-461, incl. 'invalid codeword => symbol 0, length 16' (decode.swift:1257-1261)."""
    t1 = O.HuffSpec.make([0, 3, 1, 1] + [0] * 12, [0x61, 0x62, 0x63, 0x64, 0x65])
    t2 = O.HuffSpec.make([1] * 16, list(range(0x61, 0x71)))
    t3 = O.HuffSpec.make([1, 1, 1, 1] + [0] * 12, [0x61, 0x62, 0x63, 0x64])
    assert _decode_symbols(t1, [0b11011100, 0b00110110]) == [0x64, 0x65, 0x61, 0x62, 0x63, 0x64]
    assert _decode_symbols(t2, [0b11111111, 0b11001111, 0b11111111, 0b11101110]) == [0x6B, 0x61, 0x70, 0x64]
    assert _decode_symbols(t3, [0b11110110, 0b11111110, 0b10101110, 0b11111111, 0b11111110]) == \
        [0x00, 0x62, 0x62, 0x64, 0x00]


def test_huffman_invalid_tree_rejected():
    bad = O.HuffSpec.make([3] + [0] * 15, [1, 2, 3])
    with pytest.raises(O.OracleError):
        O.huff_lookup(bad, 0)


@pytest.mark.parametrize("count", [16, 256, 4096, 65536])
def test_huffman_build_roundtrip(count):
    """tests.swift:464-503: random symbols -> init(frequencies:) -> encoder() -> bits -> decoder()."""
    rng = np.random.default_rng(count)
    syms = (rng.integers(0, 128, count) + rng.integers(0, 129, count)).astype(np.int64)
    freq = np.bincount(syms, minlength=256)
    spec = O.huff_from_frequencies(freq)
    code, ln = O.huff_encoder(spec)
    assert all(1 <= ln[s] <= 16 for s in syms)
    assert sum(spec.counts) == int((freq > 0).sum())
    # all-ones codeword is never assigned
    assert all(int(code[s]) != (1 << int(ln[s])) - 1 for s in set(syms.tolist()))
    for s in syms[:2000]:
        sym, l = O.huff_lookup(spec, (int(code[s]) << (16 - int(ln[s]))) & 0xFFFF)
        assert (sym, l) == (s, ln[s])


def test_huffman_length_limit_16():
    freq = np.zeros(256, dtype=np.int64)
    a, b = 1, 1
    for i in range(40):  # fibonacci frequencies force a >16-level tree
        freq[i] = a
        a, b = b, a + b
    spec = O.huff_from_frequencies(freq)
    assert sum(spec.counts) == 40
    code, ln = O.huff_encoder(spec)
    assert ln[:40].max() == 16 and ln[:40].min() >= 1
    kraft = sum(2.0 ** -int(l) for l in ln[:40])
    assert kraft < 1.0


# ---------------------------------------------------------------- decode regression (tests/regression/tests.swift)

def test_decode_golden(manifest):
    for v in manifest["decode"]:
        rgb, ycc, s = O.decode_rgb(golden_bytes(v["jpeg"]))
        assert sha(rgb.tobytes()) == v["rgb_sha256"], v["jpeg"]
        if "ycc_sha256" in v:
            assert sha(ycc.tobytes()) == v["ycc_sha256"], v["jpeg"]
        if "planes_sha256" in v:  # examples/decode-advanced: per-plane IDCT output before upsampling
            planes = s.idct()
            for p, h in enumerate(v["planes_sha256"]):
                assert sha(planes[p].astype(np.uint8).tobytes()) == h


def test_decode_online_golden(manifest):
    """examples/decode-online/main.swift: the image after each of the ten scans of a progressive file with a restart
    interval (DC first at bit 1, band-limited AC first scans, AC / DC refinements), as RGB -- every snapshot equals the
    reference's committed dump.  Planes whose quanta are not bound yet dequantise with the default table (zeros)."""
    v = manifest["decode_online"]
    data = golden_bytes(v["jpeg"])
    interval = [b for m, b, e in J.split(data) if m == 0xDD]
    assert interval and int.from_bytes(interval[0], "big") > 0
    for k, want in enumerate(v["rgb_sha256"]):
        s = O.Spectral.decompress_scans(data, k + 1)
        assert sha(O.unpack_rgb(s.to_rectangular()).tobytes()) == want, k
    full = O.Spectral.decompress(data)
    assert sha(O.unpack_rgb(full.to_rectangular()).tobytes()) == v["rgb_sha256"][-1]


def test_decode_restart_files(manifest):
    """tests/integration/tests.swift:100-187: restart-interval files decode without error (no reference output
    exists); the row-granular interval placement is cross-checked against a DRI-less re-encode of the same
    coefficients."""
    for rel in manifest["restart"]:
        rgb, ycc, s = O.decode_rgb(golden_bytes(rel))
        assert s.size == (256, 384)


# ---------------------------------------------------------------- encode vectors

def _scan_params(sos, s):
    n = sos[0]
    ids = [sos[1 + 2 * i] for i in range(n)]
    dcs = [sos[2 + 2 * i] >> 4 for i in range(n)]
    acs = [sos[2 + 2 * i] & 15 for i in range(n)]
    band = (sos[2 * n + 1], sos[2 * n + 2] + 1)
    al, ah = sos[2 * n + 3] & 15, sos[2 * n + 3] >> 4
    cid = [s.plane_info(p)[2] for p in range(s.ncomp)]
    return band, (al, None if ah == 0 else ah), [cid.index(i) for i in ids], dcs, acs


def _check_scans(s, expect):
    for k, sc in enumerate(expect["scans"]):
        band, bits, comps, dcs, acs = _scan_params(bytes.fromhex(sc["sos"]), s)
        ecs, dct, act = s.encode_scan(band, bits, comps, dcs, acs)
        assert len(ecs) == sc["ecs_len"] and sha(ecs) == sc["ecs_sha256"], f"scan {k}"
        for cls, tgt, counts, values in sc["dht"]:
            tab = (dct if cls == 0 else act)[tgt]
            assert tab.present and tab.as_tuple() == (bytes.fromhex(counts), bytes.fromhex(values)), f"scan {k}"


def test_encode_basic_golden(manifest):
    """examples/encode-basic: colour, 4 sampling modes, 8 compression levels: quanta, optimal tables and ECS."""
    eb = manifest["encode_basic"]
    w, h = eb["size"]
    rgb = np.frombuffer(golden_bytes(eb["rgb"]), dtype=np.uint8).reshape(h, w, 3)
    il = O.pack_rgb(rgb)
    for name, lum in (("4-4-4", (1, 1)), ("4-4-0", (1, 2)), ("4-2-2", (2, 1)), ("4-2-0", (2, 2))):
        factors = [lum, (1, 1), (1, 1)]
        planes = O.decompose(il, factors)
        for level in ("0.0", "0.125", "0.25", "0.5", "1.0", "2.0", "4.0", "8.0"):
            exp = eb["files"][f"{name}-{level}"]
            q = [O.quanta(float(level), 0), O.quanta(float(level), 1)]
            assert [t[1] for t in exp["dqt"]] == [q[0].tolist(), q[1].tolist()]
            s = O.Spectral.create((w, h), factors)
            for p in range(3):
                s.coefficients(p)[...] = O.fdct_plane(planes[p], q[0 if p == 0 else 1])
            _check_scans(s, exp)


def test_reencode_in_memory(manifest):
    """examples/in-memory: 14-scan progressive decoded to Spectral and re-encoded from coefficients."""
    exp = manifest["reencode"]["in-memory"]
    s = O.Spectral.decompress(golden_bytes(exp["source"]))
    _check_scans(s, exp)


def test_recompress_requantized(manifest):
    """examples/recompress/main.swift:15-60: requantise in the spectral domain, 4-scan progression (config #4)."""
    exp = manifest["reencode"]["recompress-requantized"]
    orig = O.Spectral.decompress(golden_bytes(exp["source"]))
    rec = O.Spectral.create(orig.size, [orig.factor(p) for p in range(3)], progressive=True)
    for p in range(3):
        q = orig.quanta(p).astype(np.int64)
        nq = np.concatenate([q[:1], np.minimum(q[1:] * 3, 255)])
        resc = (orig.coefficients(p).astype(np.int64) * q).astype(np.float64) / nq.astype(np.float64)
        rec.coefficients(p)[...] = np.trunc(resc + 0.3 * np.where(resc < 0, -1.0, 1.0)).astype(np.int16)
    _check_scans(rec, exp)


@pytest.mark.parametrize("kind", ["ii", "iii", "iv"])
def test_rotate_golden(manifest, kind):
    """examples/rotate/main.swift: lossless rotation in the spectral domain; the re-encoded scans (tables + entropy-coded
    bytes) and the permuted quantisation tables equal the reference's committed outputs."""
    exp = manifest["rotate"]["outputs"][kind]
    s = O.Spectral.decompress(golden_bytes(manifest["rotate"]["source"]))
    r = O.rotated(s, kind)
    cid = [r.plane_info(p)[2] for p in range(r.ncomp)]
    assert [t[1] for t in exp["dqt"]][:2] == [r.quanta(0).tolist(), r.quanta(1).tolist()] or \
        all(r.quanta(p).tolist() in [t[1] for t in exp["dqt"]] for p in range(r.ncomp)), cid
    _check_scans(r, exp)


@pytest.mark.skipif(not have_reference(), reason="input (1.6 MB) lives only in the reference checkout")
def test_encode_advanced_golden(manifest):
    """examples/encode-advanced: 4:2:2, 11-scan progressive with successive approximation, custom quanta."""
    exp = manifest["encode_advanced"]
    raw = open(os.path.join(REFERENCE, "examples/encode-advanced/karlie-cfdas-2011.png.rgb"), "rb").read()
    assert sha(raw) == exp["rgb_sha256"]
    w, h = exp["size"]
    rgb = np.frombuffer(raw, dtype=np.uint8).reshape(h, w, 3)
    factors = [(2, 1), (1, 1), (1, 1)]
    planes = O.decompose(O.pack_rgb(rgb), factors)
    q0 = np.array([1, 2, 2, 3, 3, 3] + [4] * 58, dtype=np.uint16)
    q1 = np.array([1, 2, 2, 5, 5, 5] + [30] * 58, dtype=np.uint16)
    s = O.Spectral.create((w, h), factors, progressive=True)
    for p in range(3):
        s.coefficients(p)[...] = O.fdct_plane(planes[p], q0 if p == 0 else q1)
    _check_scans(s, exp)


# ---------------------------------------------------------------- user-defined format, 12-bit samples (N4)

def custom_color_pixels():
    """examples/custom-color/main.swift:123-134: the gradient the example encodes, as 12-bit (r, g, b, a) samples."""
    def stride(a, b, step):  # Swift's stride(from:to:by:) over Double: a + i * step
        n = 0
        while a + n * step < b:
            n += 1
        return a + np.arange(n, dtype=np.float64) * step

    def wave(x):  # UInt16(0x0fff * (sin(2 pi x) * 0.5 + 0.5)): truncating conversion
        return np.trunc(0x0fff * (np.sin(2.0 * np.pi * x) * 0.5 + 0.5)).astype(np.uint16)
    t = stride(0.0, 1.0, 0.005)[:, None] + stride(0.0, 1.0, 0.001)[None, :]
    return np.stack([wave(t - 0.15), wave(t), wave(t + 0.15), np.full(t.shape, 0x0fff, np.uint16)], axis=-1)


def test_custom_color_golden(manifest):
    """examples/custom-color: a user-defined JPEG.Format (components 4-7, 12-bit precision) written as a 10-scan
    progression (two-component DC scans of unequal sampling, AC first + refinement per component, 16-bit DQT).
    The committed output.jpg pins, bit for bit: Rectangular.decomposed() over four planes, the 12-bit FDCT + quantiser,
    the progressive scan decoders on that file, and the progressive encoders (tables + entropy-coded bytes)."""
    cc = manifest["custom_color"]
    fmt = (cc["format"][0], cc["format"][1])
    w, h = cc["size"]
    factors = [tuple(f) for f in cc["factors"]]
    px = custom_color_pixels()
    assert px.shape == (h, w, 4)
    # main.swift:196-204: the .rgb the example writes next to the file is its INPUT, 12 bits left-aligned in two bytes
    rgb = px[..., :3]
    dump = np.stack([(rgb >> 4).astype(np.uint8), ((rgb << 4) & 0xff).astype(np.uint8)], axis=-1)
    assert sha(dump.tobytes()) == cc["rgb_sha256"]
    q = [np.array([1, 2, 2, 3, 3, 3] + [10] * 58, dtype=np.uint16), np.array([1] + [100] * 63, dtype=np.uint16)]
    assert [t[1] for t in cc["dqt"]] == [q[0].tolist(), q[1].tolist()]
    jpeg = golden_bytes(cc["jpeg"])
    # JPEG.Common does not recognise the file (jpeg.swift:370-397): DecodingError.unrecognizedColorFormat
    with pytest.raises(O.OracleError) as e:
        O.Spectral.decompress(jpeg)
    assert e.value.code == -12
    dec = O.Spectral.decompress(jpeg, format=fmt)
    assert dec.size == (w, h) and dec.ncomp == 4 and dec.precision == 12
    assert [dec.plane_info(p)[2] for p in range(4)] == fmt[0] and [dec.factor(p) for p in range(4)] == factors
    # forward path from the pixels == what the decoders read back from the reference's file
    planes = O.decompose(px, factors)
    enc = O.Spectral.create((w, h), factors, progressive=True, format=fmt)
    for p in range(4):
        enc.set_quanta(p, q[1 if p == 3 else 0])
        enc.coefficients(p)[...] = O.fdct_plane(planes[p], q[1 if p == 3 else 0], precision=12)
        assert np.array_equal(enc.coefficients(p), dec.coefficients(p)), p
        assert np.array_equal(dec.quanta(p), q[1 if p == 3 else 0]), p
    _check_scans(enc, cc)
    _check_scans(dec, cc)
    # inverse path (no committed output for it: the example only shows a difference image): 12-bit samples, small error
    out = dec.to_rectangular()
    assert out.shape == (h, w, 4) and int(out.max()) <= 0x0fff
    assert int(np.abs(out[..., :3].astype(np.int32) - rgb.astype(np.int32)).max()) <= 4
    assert np.all(out[..., 3] == 0x0fff)


def test_16bit_quanta_need_a_deep_format(manifest):
    """Spectral.push(qi:quanta:) decode.swift:2546-2557: a 16-bit DQT in an 8-bit image is
    DecodingError.invalidScanQuantizationPrecision; 8-bit tables in a 12-bit image are fine (examples/custom-color would
    have written them had its format been 8-bit: decode.swift:2530)."""
    src = golden_bytes(manifest["decode"][0]["jpeg"])
    out = bytearray()
    for m, body, ecs in J.split(src):
        if m == 0xDB:  # rewrite every table as Pq = 1
            wide = bytearray()
            for tgt, vals in J.parse_dqt(body):
                wide.append(0x10 | tgt)
                for v in vals:
                    wide += bytes([v >> 8, v & 0xff])
            body = bytes(wide)
        out += bytes([0xFF, m])
        if m not in (0xD8, 0xD9):
            out += (len(body) + 2).to_bytes(2, "big") + body + ecs
    O.Spectral.decompress(src)
    with pytest.raises(O.OracleError) as e:
        O.Spectral.decompress(bytes(out))
    assert e.value.code == -12


# ---------------------------------------------------------------- our restart-interval extension of the encoder

@pytest.mark.parametrize("rows", [1, 3])
def test_restart_extension_roundtrip(manifest, rows):
    """The reference encoder never emits DRI (encode.swift:1952-1968); ours can (whole MCU rows only, the one form
    the reference decoder places correctly, decode.swift:3205-3207).  Decoding the DRI stream must reproduce the
    coefficients exactly, for sequential and all four progressive scan kinds."""
    src = O.Spectral.decompress(golden_bytes("gold/color-progressive-1.jpg"))
    fac = [src.factor(p) for p in range(3)]
    dst = O.Spectral.create(src.size, fac, progressive=True)
    scans = [((0, 1), (1, None), [0, 1, 2]), ((1, 6), (1, None), [0]), ((1, 64), (1, None), [1]),
             ((1, 64), (1, None), [2]), ((6, 64), (1, None), [0]), ((0, 1), (0, 1), [0, 1, 2]),
             ((1, 64), (0, 1), [0]), ((1, 64), (0, 1), [1]), ((1, 64), (0, 1), [2])]
    empty = [O.HuffSpec() for _ in range(4)]
    for band, bits, comps in scans:
        width = src.blocks[0] if len(comps) > 1 else src.units(comps[0])[0]
        ecs, dct, act = src.encode_scan(band, bits, comps, [0] * len(comps), [0] * len(comps), rows * width)
        parts = J.unstuff_split(ecs)
        height = src.blocks[1] if len(comps) > 1 else src.units(comps[0])[1]
        assert len(parts) == -(-height // rows)
        dst.decode_scan(band, bits, comps, [0] * len(comps), [0] * len(comps), dct if dct[0].present else empty,
                        act if act[0].present else empty, parts, interval=rows * width)
    for p in range(3):
        assert np.array_equal(dst.coefficients(p), src.coefficients(p))


# ---------------------------------------------------------------- differential / fuzz (tests/fuzz, tests/compare of the reference)

def test_transforms_against_an_independent_float_dct():
    """tests/compare/main.swift compares the decoder with ImageMagick's float DCT (not available here): the same idea with
    scipy's orthonormal DCT-II/III in float64 as the independent implementation.  The AAN networks of decode.swift:4042-4133
    and encode.swift:123-248 must be a DCT, not merely each other's inverse: quanta 1, error below one grey level."""
    from scipy.fft import dctn, idctn
    rng = np.random.default_rng(7)
    ones = np.ones(64, dtype=np.uint16)
    zz = np.array(O.zigzag_table())
    for precision in (8, 12):
        top = (1 << precision) - 1
        # forward: random sample blocks -> coefficients (zig-zag order, level shift 2^(P-1))
        samples = rng.integers(0, top + 1, size=(8 * 6, 8 * 9)).astype(np.uint16)
        coef = O.fdct_plane(samples, ones, precision)
        for by in range(6):
            for bx in range(9):
                blk = samples[8 * by:8 * by + 8, 8 * bx:8 * bx + 8].astype(np.float64) - (1 << (precision - 1))
                want = dctn(blk, norm="ortho")
                got = np.empty((8, 8))
                for h in range(8):
                    for k in range(8):
                        got[h, k] = coef[by, bx, zz[h][k]]
                # zz[h][k] is indexed (vertical, horizontal) or (horizontal, vertical): accept the orientation that matches
                err = min(np.abs(got - want).max(), np.abs(got.T - want).max())
                assert err <= 0.5 + 1e-3, (precision, by, bx, err)
        # inverse: smooth random coefficients -> samples
        c = np.zeros((4, 5, 64), dtype=np.int16)
        c[..., 0] = rng.integers(-200, 200, size=(4, 5)) * (1 << (precision - 8))
        c[..., 1:10] = rng.integers(-30, 30, size=(4, 5, 9)) * (1 << (precision - 8))
        out = O.idct_plane(c, ones, precision)
        for by in range(4):
            for bx in range(5):
                m = np.zeros((8, 8))
                for h in range(8):
                    for k in range(8):
                        m[h, k] = c[by, bx, zz[h][k]]
                a = idctn(m, norm="ortho") + (1 << (precision - 1))
                b = idctn(m.T, norm="ortho") + (1 << (precision - 1))
                got = out[8 * by:8 * by + 8, 8 * bx:8 * bx + 8].astype(np.float64)
                err = min(np.abs(got - np.clip(a, 0, top)).max(), np.abs(got - np.clip(b, 0, top)).max())
                assert err <= 1.0 + 1e-3, (precision, by, bx, err)


@pytest.mark.parametrize("seed", range(6))
def test_fuzz_small_progressive_images_round_trip(seed):
    """tests/fuzz/main.swift: small random progressive images, quanta all 1.  Here the property is the codec's own: random
    coefficients -> every scan kind (with and without our restart-interval extension) -> the decoders return them exactly."""
    rng = np.random.default_rng(100 + seed)
    w, h = int(rng.integers(1, 40)), int(rng.integers(1, 40))
    factors = [[(1, 1), (1, 1), (1, 1)], [(2, 2), (1, 1), (1, 1)], [(2, 1), (1, 1), (1, 2)]][seed % 3]
    src = O.Spectral.create((w, h), factors, progressive=True)
    for p in range(3):
        c = src.coefficients(p)
        dense = rng.random(c.shape) < 0.25
        c[...] = np.where(dense, rng.integers(-1023, 1024, size=c.shape), 0).astype(np.int16)
        c[..., 0] = rng.integers(-1023, 1024, size=c.shape[:2])
    scans = [((0, 1), (2, None), [0, 1, 2]), ((1, 9), (1, None), [0]), ((9, 64), (1, None), [0]), ((1, 64), (1, None), [1]),
             ((1, 64), (1, None), [2]), ((0, 1), (1, 2), [0, 1, 2]), ((0, 1), (0, 1), [0, 1, 2]),
             ((1, 64), (0, 1), [0]), ((1, 64), (0, 1), [1]), ((1, 64), (0, 1), [2])]
    empty = [O.HuffSpec() for _ in range(4)]
    for rows in (0, 1):
        dst = O.Spectral.create((w, h), factors, progressive=True)
        for band, bits, comps in scans:
            width = src.blocks[0] if len(comps) > 1 else src.units(comps[0])[0]
            ecs, dct, act = src.encode_scan(band, bits, comps, [0] * len(comps), [0] * len(comps), rows * width)
            dst.decode_scan(band, bits, comps, [0] * len(comps), [0] * len(comps), dct if dct[0].present else empty,
                            act if act[0].present else empty, J.unstuff_split(ecs),
                            interval=rows * width if rows else O.INTERVAL_NONE)
        for p in range(3):
            assert np.array_equal(dst.coefficients(p), src.coefficients(p)), (seed, rows, p)
