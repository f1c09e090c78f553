"""The C++ host mirror (jpeg_b200/host/jpeg_host.{hpp,cpp} -> libjpeg_host.so): jpeg::Data::{Spectral,Planar,Rectangular}
over the C-ABI, driven through its flat jpegh_* facade.  The reference is compiled Swift, so the host side above the
C-ABI is compiled code too; these tests read like tests/regression/tests.swift:39-138 (decode every gold file, compare
pixels) and examples/encode-basic / recompress."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

import jpegfile as J
from conftest import golden_bytes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LEXING, PARSING, DECODING = -201, -202, -203


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


@pytest.fixture(scope="module")
def hostlib():
    lib = C.CDLL(os.path.join(ROOT, "jpeg_b200", "libjpeg_host.so"))
    u8p, i16p = C.POINTER(C.c_uint8), C.POINTER(C.c_int16)
    lib.jpegh_free.argtypes = [C.c_void_p]
    lib.jpegh_free.restype = None
    lib.jpegh_decompress_pixels8.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(u8p), C.POINTER(C.c_int32),
                                             C.POINTER(C.c_int32), C.c_char_p, C.c_size_t]
    lib.jpegh_decompress_coefficients.argtypes = [C.c_char_p, C.c_size_t, C.c_int32, C.POINTER(i16p), C.POINTER(C.c_int32),
                                                  C.POINTER(C.c_int32), C.c_void_p, C.c_char_p, C.c_size_t]
    lib.jpegh_recompress.argtypes = [C.c_char_p, C.c_size_t, C.c_uint64, C.POINTER(u8p), C.POINTER(C.c_size_t), C.c_char_p,
                                     C.c_size_t]
    lib.jpegh_compress_rgb8.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                        C.c_void_p, C.c_int32, C.c_uint64, C.POINTER(u8p), C.POINTER(C.c_size_t), C.c_char_p,
                                        C.c_size_t]
    lib.jpegh_rotate.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(u8p), C.POINTER(C.c_size_t), C.c_char_p, C.c_size_t]
    i32p = C.POINTER(C.c_int32)
    lib.jpegh_recompress_format.argtypes = [C.c_char_p, C.c_size_t, i32p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(u8p),
                                            C.POINTER(C.c_size_t), C.c_char_p, C.c_size_t]
    lib.jpegh_decompress_samples16.argtypes = [C.c_char_p, C.c_size_t, i32p, C.c_int32, C.c_int32,
                                               C.POINTER(C.POINTER(C.c_uint16)), i32p, i32p, C.c_char_p, C.c_size_t]
    return lib


class HostError(Exception):
    def __init__(self, code, what):
        super().__init__(f"{code}: {what}")
        self.code, self.what = code, what


def pixels(lib, data, mode=0, cosite=False):
    out, w, h = C.POINTER(C.c_uint8)(), C.c_int32(), C.c_int32()
    err = C.create_string_buffer(256)
    rc = lib.jpegh_decompress_pixels8(data, len(data), mode, int(cosite), C.byref(out), C.byref(w), C.byref(h), err, 256)
    if rc:
        raise HostError(rc, err.value.decode())
    a = np.ctypeslib.as_array(out, shape=(h.value, w.value, 3)).copy()
    lib.jpegh_free(out)
    return a


def coefficients(lib, data, plane):
    out, ux, uy = C.POINTER(C.c_int16)(), C.c_int32(), C.c_int32()
    q = np.zeros(64, np.uint16)
    err = C.create_string_buffer(256)
    rc = lib.jpegh_decompress_coefficients(data, len(data), plane, C.byref(out), C.byref(ux), C.byref(uy), q.ctypes.data, err, 256)
    if rc:
        raise HostError(rc, err.value.decode())
    a = np.ctypeslib.as_array(out, shape=(uy.value, ux.value, 64)).copy()
    lib.jpegh_free(out)
    return a, q


def recompress(lib, data, interval=0):
    out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
    err = C.create_string_buffer(256)
    rc = lib.jpegh_recompress(data, len(data), interval, C.byref(out), C.byref(n), err, 256)
    if rc:
        raise HostError(rc, err.value.decode())
    b = bytes(np.ctypeslib.as_array(out, shape=(n.value,)))
    lib.jpegh_free(out)
    return b


def compress_rgb(lib, rgb, factors, quanta, scans, progressive=False, interval=0):
    h, w, _ = rgb.shape
    rgb = np.ascontiguousarray(rgb)
    f = np.array(factors, np.int32).reshape(-1)
    q = np.ascontiguousarray(np.stack(quanta), dtype=np.uint16)
    sc = np.zeros((len(scans), 17), np.int32)
    for k, (band, bits, comps) in enumerate(scans):
        sc[k, :5] = band[0], band[1], bits[0], (-1 if bits[1] is None else bits[1]), len(comps)
        for i, c in enumerate(comps):
            sc[k, 5 + 3 * i:8 + 3 * i] = c
    out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
    err = C.create_string_buffer(256)
    rc = lib.jpegh_compress_rgb8(rgb.ctypes.data, w, h, len(factors), f.ctypes.data, q.ctypes.data, int(progressive),
                                 sc.ctypes.data, len(scans), interval, C.byref(out), C.byref(n), err, 256)
    if rc:
        raise HostError(rc, err.value.decode())
    b = bytes(np.ctypeslib.as_array(out, shape=(n.value,)))
    lib.jpegh_free(out)
    return b


# ------------------------------------------------------------------------------------------------------- CPU
def test_host_library_loads_and_exports(hostlib):
    for name in ("jpegh_decompress_pixels8", "jpegh_decompress_coefficients", "jpegh_recompress", "jpegh_compress_rgb8",
                 "jpegh_free"):
        assert hasattr(hostlib, name)


def test_lexer_and_parser_errors_need_no_device(hostlib):
    """decode.swift:130-190 / 475-1005: container errors are raised by the host before any device work."""
    with pytest.raises(HostError) as e:
        pixels(hostlib, b"\x00\x01\x02")
    assert e.value.code == LEXING and e.value.what == "truncatedEntropyCodedSegment"
    with pytest.raises(HostError) as e:
        pixels(hostlib, b"\x12\xff\xd8")
    assert e.value.code == LEXING and e.value.what == "invalidMarkerSegmentPrefix"
    with pytest.raises(HostError) as e:
        pixels(hostlib, b"\xff\xd9")
    assert e.value.code == DECODING and e.value.what == "missingStartOfImage"
    with pytest.raises(HostError) as e:
        pixels(hostlib, b"\xff\xd8\xff\xc4\x00\x05\x20\x00\x00")
    assert e.value.code == PARSING
    with pytest.raises(HostError) as e:
        pixels(hostlib, b"\xff\xd8\xff\xdb\x00\x03\x05")
    assert e.value.code == PARSING and e.value.what == "invalidQuantizationTargetCode"
    with pytest.raises(HostError) as e:
        pixels(hostlib, b"\xff\xd8\xff\xe0\x00\x10")
    assert e.value.code == LEXING and e.value.what == "truncatedMarkerSegmentBody"


def recompress_format(lib, data, components, precision, jfif=False):
    out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
    err = C.create_string_buffer(256)
    ids = (C.c_int32 * len(components))(*components)
    rc = lib.jpegh_recompress_format(data, len(data), ids, len(components), precision, int(jfif), C.byref(out), C.byref(n), err, 256)
    if rc:
        raise HostError(rc, err.value.decode())
    b = bytes(np.ctypeslib.as_array(out, shape=(n.value,)))
    lib.jpegh_free(out)
    return b


def samples16(lib, data, components, precision):
    out, w, h = C.POINTER(C.c_uint16)(), C.c_int32(), C.c_int32()
    err = C.create_string_buffer(256)
    ids = (C.c_int32 * len(components))(*components)
    rc = lib.jpegh_decompress_samples16(data, len(data), ids, len(components), precision, C.byref(out), C.byref(w), C.byref(h),
                                        err, 256)
    if rc:
        raise HostError(rc, err.value.decode())
    a = np.ctypeslib.as_array(out, shape=(h.value, w.value, len(components))).copy()
    lib.jpegh_free(out)
    return a


def test_format_recognition_needs_no_device(manifest, hostlib):
    """jpeg::Format (examples/custom-color/main.swift:41-63): a file is decoded only by a format that recognises its component
    keys and precision -- DecodingError.unrecognizedColorFormat otherwise (decode.swift:2374-2381), before any device work."""
    data = golden_bytes(manifest["custom_color"]["jpeg"])
    for comps, prec in (([4, 5, 6, 7], 8), ([4, 5, 6], 12), ([1, 2, 3, 4], 12)):
        with pytest.raises(HostError) as e:
            recompress_format(hostlib, data, comps, prec)
        assert e.value.code == DECODING and e.value.what == "unrecognizedColorFormat", (comps, prec)
    with pytest.raises(HostError) as e:
        pixels(hostlib, data)  # JPEG.Common
    assert e.value.code == DECODING and e.value.what == "unrecognizedColorFormat"
    with pytest.raises(HostError) as e:  # an 8-bit file is not a 12-bit format's
        recompress_format(hostlib, golden_bytes("gold/color-sequential-2.jpg"), [1, 2, 3], 12)
    assert e.value.code == DECODING and e.value.what == "unrecognizedColorFormat"


def test_compression_level_quanta_cpp(manifest, hostlib):
    """jpeg::CompressionLevel (encode.swift:260-333): the DQT tables of the 32 files examples/encode-basic wrote"""
    hostlib.jpegh_compression_level_quanta.argtypes = [C.c_int32, C.c_double, C.c_void_p]
    hostlib.jpegh_compression_level_quanta.restype = None
    for name, exp in manifest["encode_basic"]["files"].items():
        level = float(name.rsplit("-", 1)[1])
        got = []
        for chroma in (0, 1):
            q = np.zeros(64, np.uint16)
            hostlib.jpegh_compression_level_quanta(chroma, level, q.ctypes.data)
            got.append(q.tolist())
        assert [t[1] for t in exp["dqt"]] == got, name


def test_no_cpu_fallback_without_device(hostlib):
    """Without a GPU the first hot-path stage fails loudly with the CUDA error code; nothing is computed on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(HostError) as e:
        pixels(hostlib, golden_bytes("gold/color-sequential-2.jpg"))
    assert e.value.code == -100


# ------------------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


@pytest.mark.gpu
def test_gold_files_through_cpp_host(manifest, hostlib, O):
    """tests/regression/tests.swift:39-138 through jpeg::Data::Spectral::decompress -> idct -> interleaved -> unpack."""
    for v in manifest["decode"]:
        data = golden_bytes(v["jpeg"])
        rgb = pixels(hostlib, data, 0)
        assert sha(rgb.tobytes()) == v["rgb_sha256"], v["jpeg"]
        if "ycc_sha256" in v:
            assert sha(pixels(hostlib, data, 1).tobytes()) == v["ycc_sha256"], v["jpeg"]
        assert np.array_equal(pixels(hostlib, data, 2), rgb), v["jpeg"]  # fused Spectral -> RGB8
        assert np.array_equal(pixels(hostlib, data, 4), rgb), v["jpeg"]  # host lexer instead of the GPU lexer
        ref = O.Spectral.decompress(data)
        for p in range(ref.ncomp):
            coef, q = coefficients(hostlib, data, p)
            assert np.array_equal(coef, ref.coefficients(p)), (v["jpeg"], p)
            assert np.array_equal(q, ref.quanta(p))
    for name in manifest["restart"]:
        data = golden_bytes(name)
        want, _, _ = O.decode_rgb(data)
        assert np.array_equal(pixels(hostlib, data, 0), want), name
        assert np.array_equal(pixels(hostlib, data, 4), want), name


@pytest.mark.gpu
def test_cosited_through_cpp_host(hostlib, O):
    data = golden_bytes("gold/color-sequential-2.jpg")
    ref = O.Spectral.decompress(data)
    want = O.unpack_rgb(ref.to_rectangular(cosited=True))
    assert np.array_equal(pixels(hostlib, data, 0, cosite=True), want)
    assert np.array_equal(pixels(hostlib, data, 2, cosite=True), want)


@pytest.mark.gpu
def test_recompress_through_cpp_host(manifest, hostlib, O):
    """examples/recompress: decompress -> compress with the file's own progression; an independent decoder (the
    oracle) must read our file back to the same coefficients, with and without restart intervals."""
    for name in ("gold/color-progressive-1.jpg", "gold/color-sequential-2.jpg", "gold/grayscale-progressive-1.jpg"):
        data = golden_bytes(name)
        src = O.Spectral.decompress(data)
        for interval in (0, src.blocks[0]):
            if interval and name != "gold/color-sequential-2.jpg":
                continue  # per-scan row widths differ in non-interleaved progressive scans
            blob = recompress(hostlib, data, interval)
            back = O.Spectral.decompress(blob)
            for p in range(src.ncomp):
                assert np.array_equal(back.coefficients(p), src.coefficients(p)), (name, interval, p)
            assert np.array_equal(pixels(hostlib, blob, 0), O.decode_rgb(data)[0])


@pytest.mark.gpu
def test_encode_basic_golden_through_cpp_host(manifest, hostlib):
    """examples/encode-basic through jpeg::Data::Rectangular::pack -> decomposed -> fdct -> compress: the DHT tables
    and entropy-coded bytes inside our file equal the reference's (digests of the 32 committed files)."""
    eb = manifest["encode_basic"]
    w, h = eb["size"]
    rgb = np.frombuffer(golden_bytes(eb["rgb"]), dtype=np.uint8).reshape(h, w, 3)
    scans = [((0, 64), (0, None), [(0, 0, 0)]), ((0, 64), (0, None), [(1, 1, 1), (2, 1, 1)])]
    for name, lum in (("4-4-4", (1, 1)), ("4-4-0", (1, 2)), ("4-2-2", (2, 1)), ("4-2-0", (2, 2))):
        for tag in ("0.0", "0.25", "1.0", "8.0"):
            exp = eb["files"][f"{name}-{tag}"]
            q = [np.array(exp["dqt"][0][1], np.uint16), np.array(exp["dqt"][1][1], np.uint16)]
            blob = compress_rgb(hostlib, rgb, [lum, (1, 1), (1, 1)], [q[0], q[1], q[1]], scans)
            segs = J.split(blob)
            got_scans = [ecs for m, _, ecs in segs if m == 0xDA]
            assert len(got_scans) == 2
            for ecs, sc in zip(got_scans, exp["scans"]):
                assert len(ecs) == sc["ecs_len"] and sha(ecs) == sc["ecs_sha256"], (name, tag)
            dht = {}
            for m, body, _ in segs:
                if m == 0xC4:
                    for cls, tgt, counts, values in J.parse_dht(body):
                        dht[(cls, tgt)] = (bytes(counts), bytes(values))
            for sc in exp["scans"]:
                for cls, tgt, counts, values in sc["dht"]:
                    assert dht[(cls, tgt)] == (bytes.fromhex(counts), bytes.fromhex(values)), (name, tag)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,code", [("ii", 2), ("iii", 3), ("iv", 4)])
def test_rotate_through_cpp_host(manifest, hostlib, kind, code):
    """examples/rotate as one call of the C++ host (decompress -> Spectral::rotated -> compress): every scan of our file --
    entropy-coded bytes and the Huffman tables in front of it -- and the permuted quantisation tables equal the reference's
    committed output."""
    data = golden_bytes(manifest["rotate"]["source"])
    out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
    err = C.create_string_buffer(256)
    rc = hostlib.jpegh_rotate(data, len(data), code, C.byref(out), C.byref(n), err, 256)
    assert rc == 0, err.value.decode()
    blob = bytes(np.ctypeslib.as_array(out, shape=(n.value,)))
    hostlib.jpegh_free(out)
    exp = manifest["rotate"]["outputs"][kind]
    segs = J.split(blob)
    got = [(body, ecs) for m, body, ecs in segs if m == 0xDA]
    assert len(got) == len(exp["scans"])
    for (body, ecs), sc in zip(got, exp["scans"]):
        assert body.hex() == sc["sos"]
        assert len(ecs) == sc["ecs_len"] and sha(ecs) == sc["ecs_sha256"], kind
    dqt = [[t, q] for m, body, _ in segs if m == 0xDB for t, q in J.parse_dqt(body)]
    assert dqt == exp["dqt"]


@pytest.mark.gpu
def test_custom_format_through_cpp_host(manifest, hostlib, O):
    """examples/custom-color through jpeg::Data::Spectral with a jpeg::Format: components 4-7, 12-bit samples, 16-bit DQT, ten
    progressive scans.  compress() of the decoded image returns the reference's file byte for byte; the decoded 16-bit values
    equal the oracle's."""
    cc = manifest["custom_color"]
    comps, prec = cc["format"]
    data = golden_bytes(cc["jpeg"])
    out = recompress_format(hostlib, data, comps, prec, jfif=False)
    assert out == data and sha(out) == cc["file_sha256"]
    ref = O.Spectral.decompress(data, format=(comps, prec)).to_rectangular()
    got = samples16(hostlib, data, comps, prec)
    assert got.shape == ref.shape and np.array_equal(got, ref)
    # planes follow the format's order, not the frame header's
    rev = samples16(hostlib, data, comps[::-1], prec)
    assert np.array_equal(rev, ref[..., ::-1])

