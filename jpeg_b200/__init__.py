"""jpeg_b200 -- B200-native JPEG block-transform hot path behind the tayloraswift/jpeg staged API.

The product is `libjpeg_sm100.so` (hand-written CUDA for sm_100a, C-ABI in include/jpeg_sm100.h).  This package is
the thin host side used by the tests and the benchmark: a ctypes binding of that C-ABI (`jpeg_b200.lib`) and a
Python mirror of the reference's staged interface (`jpeg_b200.host`: Spectral / Planar / Rectangular with
decompress / idct / interleaved / unpack and pack / decomposed / fdct / compress), whose hot-path bodies are calls
into the library.  There is no CPU fallback: importing works anywhere, but every compute entry point raises if
the shared library or a B200 is missing.
"""
from . import lib  # noqa: F401

__all__ = ["lib"]
