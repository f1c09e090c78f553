"""The oracle extended to ITU-T T.81 restart-interval placement (interval e starts at MCU e * Ri wherever that falls in its row)
-- what the GPU tests of JPEG_SM100_SCAN_T81 compare with -- checked on the CPU: the reference itself places intervals by rows
(decode.swift:3205-3207) and cannot decode such files, so the extension is pinned to the reference where both are defined."""
import numpy as np
import pytest

import jpegfile as J
from conftest import golden_bytes


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


SCANS = [((0, 64), (0, None), [0, 1, 2]), ((0, 1), (1, None), [0, 1, 2]), ((1, 6), (1, None), [0]), ((1, 64), (0, 1), [1])]


def test_virtual_grid_equals_the_reference_for_whole_row_intervals(O):
    """An interval of whole MCU rows means the same thing under both placements: the scan encoded on the virtual grid of
    gcd(Ri, MCUs) columns must be byte-identical to the reference-order encoding, tables included."""
    src = O.Spectral.decompress(golden_bytes("gold/color-progressive-1.jpg"))
    for band, bits, comps in SCANS:
        W, H = src.blocks if len(comps) > 1 else src.units(comps[0])
        sel = [0, 1, 1][:len(comps)] if len(comps) > 1 else [0]
        for rows in (1, 2, 7):
            ival = rows * W
            want, dct, act = src.encode_scan(band, bits, comps, sel, sel, ival)
            v = J.oracle_on_virtual_grid(O, src, comps, ival)
            assert v.blocks[0] % W == 0 or W % v.blocks[0] == 0
            got, vdc, vac = v.encode_scan(band, bits, list(range(len(comps))), sel, sel, ival)
            assert got == want, (band, bits, comps, rows)
            for a, b in zip(vdc + vac, dct + act):
                assert a.present == b.present and (not a.present or a.as_tuple() == b.as_tuple())


def test_virtual_grid_round_trip_for_odd_intervals(O):
    """Intervals that are not whole rows: encode on the virtual grid, decode on the virtual grid, copy back -> the coefficients
    (within the scan's band and bits) of the source; the segment count is ceil(MCUs / Ri)."""
    src = O.Spectral.decompress(golden_bytes("gold/color-progressive-1.jpg"))
    band, bits, comps = (0, 64), (0, None), [0, 1, 2]
    W, H = src.blocks
    for ival in (1, 7, W + 3, W * H - 1, W * H + 5):
        v = J.oracle_on_virtual_grid(O, src, comps, ival)
        ecs, dct, act = v.encode_scan(band, bits, [0, 1, 2], [0, 1, 1], [0, 1, 1], ival)
        parts = J.unstuff_split(ecs)
        assert len(parts) == -(-W * H // ival)
        back = O.Spectral.create(v.size, [v.factor(p) for p in range(3)])
        back.decode_scan(band, bits, [0, 1, 2], [0, 1, 1], [0, 1, 1], dct, act, parts, interval=ival)
        for p in range(3):
            assert np.array_equal(back.coefficients(p), v.coefficients(p)), (ival, p)
