"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports exactly the symbols include/jpeg_sm100.h
declares, and fails loudly (no fallback) when there is no B200."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "jpeg_sm100.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(jpeg_sm100_[a-z0-9_]+)\s*\(", text))


def test_library_exports_every_declared_symbol():
    from jpeg_b200 import lib
    declared = _declared()
    assert declared == set(lib.SYMBOLS), declared ^ set(lib.SYMBOLS)
    L = lib.load()
    for name in declared:
        assert hasattr(L, name), name
    assert L.jpeg_sm100_abi_version() == 1
    assert L.jpeg_sm100_error_string(-1) == b"truncatedEntropyCodedSegment"


def test_struct_layouts_match_header():
    from jpeg_b200 import lib
    assert C.sizeof(lib.HuffTable) == 4 + 16 + 256
    assert C.sizeof(lib.ScanDesc) == 4 * 5 + 4 * 20 + 8
    assert C.sizeof(lib.PlaneI16) == 16 and C.sizeof(lib.PlaneU16) == 24
    assert C.sizeof(lib.DevSpectral) == 8 + 4 * 32 and C.sizeof(lib.DevPlanar) == 16 + 4 * 32


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point must fail; nothing silently runs on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from jpeg_b200 import host, lib
    h = C.c_void_p()
    assert lib.load().jpeg_sm100_create(0, C.byref(h)) == lib.ERR_CUDA
    with pytest.raises(lib.JpegSm100Error):
        host.default_context()


def test_product_never_imports_the_oracle():
    """Only tests/, smoke() and bench.py's baseline legs may touch oracle/."""
    for base, _, files in os.walk(os.path.join(ROOT, "jpeg_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(base, f)).read()
                code = "\n".join(l for l in text.splitlines() if "oracle" in l and not l.strip().startswith(("//", "#", "*", '"')))
                assert "import oracle" not in code and "from oracle" not in code and "jpeg_oracle" not in code, f


def test_bindings_name_only_declared_symbols():
    """The reference-side glue (swift/JPEGSM100Shim.swift, which cannot be compiled here), INTEGRATION.md and the C++ host
    mirror may only name entry points, types and constants the header declares."""
    header = open(os.path.join(ROOT, "include", "jpeg_sm100.h")).read()
    known = set(re.findall(r"\b(jpeg_sm100_[a-z0-9_]+)\b", header)) | set(re.findall(r"\b(JPEG_SM100_[A-Z0-9_]+)\b", header))
    for rel in ("swift/JPEGSM100Shim.swift", "INTEGRATION.md", "jpeg_b200/host/jpeg_host.cpp", "jpeg_b200/host/jpeg_host.hpp"):
        text = open(os.path.join(ROOT, rel)).read()
        used = set(re.findall(r"\b(jpeg_sm100_[a-z0-9_]+)\b", text)) | set(re.findall(r"\b(JPEG_SM100_[A-Z0-9_]+)\b", text))
        used -= {"jpeg_sm100_h"}
        # prose may abbreviate a family of entry points with a trailing underscore or wildcard (jpeg_sm100_unpack_*8)
        unknown = {u for u in used - known if not any(k.startswith(u) for k in known)}
        assert not unknown, (rel, sorted(unknown))


def _calls(text):
    """(name, number of arguments, line) of every jpeg_sm100_*( ... ) in `text`"""
    out = []
    for m in re.finditer(r"\b(jpeg_sm100_[a-z0-9_]+)\s*\(", text):
        i = j = m.end()
        depth = 1
        while depth and j < len(text):
            depth += {"(": 1, ")": -1}.get(text[j], 0)
            j += 1
        args, d, n = text[i:j - 1], 0, 0
        for ch in args:
            d += {"(": 1, "[": 1, "{": 1, ")": -1, "]": -1, "}": -1}.get(ch, 0)
            n += ch == "," and d == 0
        out.append((m.group(1), 0 if not args.strip() or args.strip() == "void" else n + 1, text[:m.start()].count("\n") + 1))
    return out


def test_bindings_pass_as_many_arguments_as_the_header_declares():
    """A compile check by other means for the Swift glue (no Swift toolchain here), and a cross-check for the C++ host."""
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "jpeg_sm100.h")).read(), flags=re.S)
    declared = {name: n for name, n, _ in _calls(header)}
    assert len(declared) >= 40
    for rel in ("swift/JPEGSM100Shim.swift", "jpeg_b200/host/jpeg_host.cpp"):
        text = re.sub(r"//.*", "", open(os.path.join(ROOT, rel)).read())
        used = [(name, n, line) for name, n, line in _calls(text) if name in declared]
        assert used, rel
        for name, n, line in used:
            assert n == declared[name], (rel, line, name, n, declared[name])


def test_ctypes_binding_matches_the_header_argument_for_argument():
    """jpeg_b200/lib.py declares restype/argtypes by hand: the number of arguments, and integer vs pointer vs 64-bit for each
    one, must agree with include/jpeg_sm100.h."""
    from jpeg_b200 import lib
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "jpeg_sm100.h")).read(), flags=re.S)

    def kind_of_c(decl):
        decl = decl.strip()
        if "*" in decl or "[" in decl:
            return "ptr"
        if re.search(r"\b(uint64_t|size_t|int64_t)\b", decl):
            return "i64"
        if re.search(r"\b(double)\b", decl):
            return "f64"
        return "i32"

    def kind_of_ctypes(t):
        if t in (C.c_uint64, C.c_int64, C.c_size_t):
            return "i64"
        if t in (C.c_int, C.c_uint32, C.c_int32):
            return "i32"
        if t is C.c_double:
            return "f64"
        return "ptr"  # c_void_p, c_char_p, POINTER(...)

    checked = 0
    for m in re.finditer(r"\b(jpeg_sm100_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S):
        name, params = m.group(1), m.group(2)
        if name not in lib.SYMBOLS:
            continue
        params = [] if params.strip() in ("", "void") else [p for p in params.split(",")]
        _, argtypes = lib.SYMBOLS[name]
        assert len(params) == len(argtypes), (name, len(params), len(argtypes))
        for k, (p, t) in enumerate(zip(params, argtypes)):
            assert kind_of_c(p) == kind_of_ctypes(t), (name, k, p.strip(), t)
        checked += 1
    assert checked == len(lib.SYMBOLS), (checked, len(lib.SYMBOLS))


def test_swift_check_maps_every_status_code_of_the_header():
    """swift/JPEGSM100Shim.swift cannot be compiled here: every status code the header defines must be named by a `case` of
    SM100.check(), so that callers matching on the reference's error enums never see a documented code as a CUDA failure."""
    header = open(os.path.join(ROOT, "include", "jpeg_sm100.h")).read()
    codes = {int(v) for v in re.findall(r"JPEG_SM100_(?:OK|ERR_[A-Z_]+)\s*=\s*(-?\d+)", header)}
    assert {0, -1, -8, -9, -11, -100} <= codes
    swift = open(os.path.join(ROOT, "swift", "JPEGSM100Shim.swift")).read()
    body = swift[swift.index("static func check(_ status:Int32) throws"):]
    body = body[:body.index("// MARK:")]
    handled = set()
    for m in re.finditer(r"case\s+([-0-9,\s]+):", body):
        handled |= {int(x) for x in m.group(1).replace(" ", "").split(",") if x}
    # -100 (CUDA) is the documented meaning of `default`
    assert codes - handled <= {-100}, sorted(codes - handled)
    from jpeg_b200 import lib
    assert {getattr(lib, k) for k in dir(lib) if k.startswith("ERR_") or k == "OK"} == codes
