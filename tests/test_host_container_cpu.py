"""The Python host mirror's container logic (lexer, DQT/DHT/SOF/SOS parsers, progression bookkeeping, format recognition,
the writer) on CPU: jpeg_b200/host.py run with the ORACLE's scan codec in place of the two C-ABI calls that need a GPU
(jpeg_sm100_decode_scan / jpeg_sm100_encode_scan).  What is checked is the host code, not the codec: every golden file must
parse to the oracle's own coefficients, and the one file the reference wrote without metadata (examples/custom-color) must
come back out of the writer byte for byte.  The GPU twins of these tests live in test_gpu_parity.py.
"""
import hashlib

import numpy as np
import pytest

import jpegfile as J
from conftest import golden_bytes
from oracle import oracle as O


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


@pytest.fixture()
def host(monkeypatch):
    from jpeg_b200 import host as H
    from jpeg_b200 import lib as L

    class OracleBacked(H.Spectral):
        def _mirror(self):
            o = O.Spectral.create(self.size, [p.factor for p in self.planes], self.process == 2,
                                  format=([p.comp_id for p in self.planes], self.precision))
            for i, p in enumerate(self.planes):
                o.coefficients(i)[...] = p.coef
            return o

        def decode_scan(self, band, bits, comps, dc_tables, ac_tables, ecss, interval, extend=False):
            def spec(t):
                return O.HuffSpec.make(bytes(t.counts), bytes(t.values)[:sum(t.counts)]) if t is not None and t.present \
                    else O.HuffSpec()
            o = self._mirror()
            try:
                o.decode_scan(band, bits, [c[0] for c in comps], [c[1] for c in comps], [c[2] for c in comps],
                              [spec(t) for t in dc_tables], [spec(t) for t in ac_tables], ecss,
                              O.INTERVAL_NONE if interval is None else interval, extend)
            except O.OracleError as e:
                raise H.DecodingError(str(e))
            for i, p in enumerate(self.planes):  # rows grown by `extend` are cropped by set(height:) (decode.swift:3949)
                p.coef = o.coefficients(i)[:p.units[1], :p.units[0]].copy()

        def encode_scan(self, band, bits, comps, interval_mcus=0):
            ecs, d, a = self._mirror().encode_scan(band, bits, [c[0] for c in comps], [c[1] for c in comps],
                                                   [c[2] for c in comps], interval_mcus)

            def tab(t):
                return L.HuffTable.make(*t.as_tuple()) if t.present else L.HuffTable()
            return ecs, [tab(t) for t in d], [tab(t) for t in a]

    monkeypatch.setattr(H, "Spectral", OracleBacked)
    monkeypatch.setattr(H, "default_context", lambda: object())
    return H


def test_golden_files_parse_to_the_oracles_coefficients(manifest, host):
    files = [v["jpeg"] for v in manifest["decode"]] + list(manifest["restart"])
    for rel in files:
        data = golden_bytes(rel)
        s = host.Spectral.decompress(data)
        ref = O.Spectral.decompress(data)
        assert s.size == ref.size and s.ncomp == ref.ncomp and s.precision == 8, rel
        for p in range(s.ncomp):
            assert s.planes[p].comp_id == ref.plane_info(p)[2] and s.planes[p].factor == ref.factor(p), rel
            assert np.array_equal(s.planes[p].coef, ref.coefficients(p)), (rel, p)
            assert np.array_equal(s.quanta[s.planes[p].q], ref.quanta(p)), (rel, p)


def test_height_defined_by_dnl(manifest, host):
    """decode.swift:3905-3924: a frame height of 0 is resolved by the DNL segment that follows the first scan; a DNL height
    other than the frame's crops / extends the image.  The host resolves it before the scan is decoded."""
    for rel in ("gold/color-sequential-1.jpg", "gold/color-progressive-2.jpg", "gold/grayscale-sequential-1.jpg"):
        data = golden_bytes(rel)
        h = O.Spectral.decompress(data).size[1]
        for fh, dh in ((0, h), (h, h - 21), (h, h + 40)):
            mod = J.with_dnl(data, fh, dh)
            if dh > h and "progressive" in rel:  # the later scans of a progressive file run out of data on the taller image
                with pytest.raises(O.OracleError):
                    O.Spectral.decompress(mod)
                with pytest.raises(host.DecodingError, match="truncatedEntropyCodedSegment"):
                    host.Spectral.decompress(mod)
                continue
            s, ref = host.Spectral.decompress(mod), O.Spectral.decompress(mod)
            assert s.size == ref.size == (ref.size[0], dh), (rel, fh, dh)
            for p in range(s.ncomp):
                assert np.array_equal(s.planes[p].coef, ref.coefficients(p)), (rel, fh, dh, p)
        with pytest.raises(host.DecodingError, match="missingHeightRedefinitionSegment"):
            host.Spectral.decompress(J.with_dnl(data, 0, h).replace(bytes([0xFF, 0xDC, 0, 4]) + h.to_bytes(2, "big"), b""))


def test_online_decoding_hook(manifest, host):
    """examples/decode-online/main.swift:252-282: the capture closure sees the image after every scan"""
    v = manifest["decode_online"]
    data = golden_bytes(v["jpeg"])
    seen = []

    def capture(s, scan):
        k = len(seen)
        ref = O.Spectral.decompress_scans(data, k + 1)
        assert len(s.scans) == k + 1 and s.scans[-1] is scan
        for p in range(s.ncomp):
            assert np.array_equal(s.planes[p].coef, ref.coefficients(p)), (k, p)
            assert np.array_equal(s.quanta[s.planes[p].q], ref.quanta(p)), (k, p)  # not bound yet -> the default table
        seen.append((scan.band, scan.bits, [c[0] for c in scan.comps]))

    host.Spectral.decompress(data, on_scan=capture)
    assert len(seen) == len(v["rgb_sha256"]) == 10
    assert seen[0] == ((0, 1), (1, None), [0, 1, 2]) and seen[-1] == ((1, 64), (0, 1), [0])


def test_custom_format_file_round_trips_byte_for_byte(manifest, host):
    """examples/custom-color/main.swift: components 4-7, 12-bit precision, 16-bit DQT, ten scans, no metadata."""
    cc = manifest["custom_color"]
    fmt = host.Format(tuple(cc["format"][0]), cc["format"][1])
    data = golden_bytes(cc["jpeg"])
    with pytest.raises(host.DecodingError, match="unrecognizedColorFormat"):
        host.Spectral.decompress(data)  # JPEG.Common knows neither the components nor the precision
    with pytest.raises(host.DecodingError, match="unrecognizedColorFormat"):
        host.Spectral.decompress(data, format=host.Format((4, 5, 6, 7), 8))
    with pytest.raises(host.DecodingError, match="unrecognizedColorFormat"):
        host.Spectral.decompress(data, format=host.Format((4, 5, 6), 12))
    s = host.Spectral.decompress(data, format=fmt)
    ref = O.Spectral.decompress(data, format=(list(fmt.components), fmt.precision))
    assert s.precision == 12 and [p.comp_id for p in s.planes] == [4, 5, 6, 7] and len(s.scans) == 10
    for p in range(4):
        assert np.array_equal(s.planes[p].coef, ref.coefficients(p)), p
    # a format may order its planes differently from the frame header (jpeg.swift:1296-1307)
    t = host.Spectral.decompress(data, format=host.Format((7, 6, 5, 4), 12))
    assert [p.comp_id for p in t.planes] == [7, 6, 5, 4] and np.array_equal(t.planes[0].coef, ref.coefficients(3))
    out = s.compress(quanta_slots={s.planes[0].q: 0, s.planes[3].q: 1}, jfif=False)
    assert out == data and sha(out) == cc["file_sha256"]


def test_sixteen_bit_tables_in_an_eight_bit_image_are_rejected(manifest, host):
    """Spectral.push(qi:quanta:) decode.swift:2546-2557 -> DecodingError.invalidScanQuantizationPrecision"""
    src = golden_bytes(manifest["decode"][0]["jpeg"])
    out = bytearray()
    for m, body, ecs in J.split(src):
        if m == 0xDB:
            wide = bytearray()
            for tgt, vals in J.parse_dqt(body):
                wide.append(0x10 | tgt)
                for v in vals:
                    wide += bytes([v >> 8, v & 0xff])
            body = bytes(wide)
        out += bytes([0xFF, m])
        if m not in (0xD8, 0xD9):
            out += (len(body) + 2).to_bytes(2, "big") + body + ecs
    host.Spectral.decompress(src)
    with pytest.raises(host.DecodingError, match="invalidScanQuantizationPrecision"):
        host.Spectral.decompress(bytes(out))


def test_compression_level_quanta(manifest):
    """JPEG.CompressionLevel (encode.swift:260-333) in the host mirror: equal to the oracle's for a sweep of levels, and to the
    DQT segments of the 32 files examples/encode-basic wrote."""
    from jpeg_b200.host import CompressionLevel
    for level in (0.0, 0.01, 0.125, 0.25, 1 / 3, 0.5, 0.75, 1.0, 1.5, 2.0, 2.5, 4.0, 8.0, 100.0):
        assert np.array_equal(CompressionLevel.luminance(level).quanta, O.quanta(level, 0)), level
        assert np.array_equal(CompressionLevel.chrominance(level).quanta, O.quanta(level, 1)), level
    for name, exp in manifest["encode_basic"]["files"].items():
        level = float(name.rsplit("-", 1)[1])
        assert [t[1] for t in exp["dqt"]] == [CompressionLevel.luminance(level).quanta.tolist(),
                                              CompressionLevel.chrominance(level).quanta.tolist()], name

