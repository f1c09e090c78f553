import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import jpegfile as J
from oracle import oracle as O
import test_gpu_parity as T
from jpeg_b200 import host as H, lib
src = O.Spectral.decompress(T.golden_bytes("gold/color-progressive-1.jpg"))
fac = [src.factor(p) for p in range(3)]
for mode in ("", "seq", "flat"):
    if mode: os.environ["JPEG_SM100_HUFF"] = mode
    for ival in (7, 23, 1, 605, 40, 20):
        band, bits, comps = (0, 64), (0, None), [0, 1, 2]
        sel = [0, 1, 1]
        W, Hh = src.blocks
        if ival % W:
            v = T._oracle_on_virtual_grid(O, src, comps, ival)
            want, dct, act = v.encode_scan(band, bits, [0, 1, 2], sel, sel, ival)
        else:
            want, dct, act = src.encode_scan(band, bits, comps, sel, sel, ival)
        parts = J.unstuff_split(want)
        dst = H.Spectral(src.size, fac, process=2)
        try:
            dst.decode_scan(band, bits, [(c, d, d) for c, d in zip(comps, sel)], T._to_lib_tables(H, dct), T._to_lib_tables(H, act), parts, ival, extend=lib.SCAN_T81)
            ok = all(np.array_equal(dst.planes[p].coef, src.coefficients(p)) for p in range(3))
            print(mode, ival, len(parts), "decoded, equal:", ok)
        except lib.JpegSm100Error as e:
            print(mode, ival, len(parts), "error", e)
