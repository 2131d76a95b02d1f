"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/sa_b200.h declares;
the ctypes struct mirrors the header; argument validation errors come back as codes + text (no compute
calls here -- there is no GPU)."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from stableanimator_b200 import build, _lib
    build.build()
    return _lib.lib()


def test_every_declared_symbol_is_exported(lib):
    hdr = (ROOT / "include" / "sa_b200.h").read_text()
    names = sorted(set(re.findall(r"\b(sa_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n
    from stableanimator_b200 import _lib
    assert set(names) - {"sa_version", "sa_last_error"} == set(_lib.SIGNATURES)
    assert lib.sa_version() == 100


def test_gemm_args_struct_matches_header():
    from stableanimator_b200._lib import GemmArgs
    hdr = (ROOT / "include" / "sa_b200.h").read_text()
    body = hdr[hdr.index("typedef struct sa_gemm_args {"):hdr.index("} sa_gemm_args;")]
    fields = []
    for ln in body.splitlines()[1:]:
        ln = ln.split("/*")[0].strip().rstrip(";")
        if not ln:
            continue
        decl = ln.split()
        names = " ".join(decl[1:]) if decl[0] != "const" else " ".join(decl[2:])
        for nm in names.split(","):
            fields.append(nm.strip().lstrip("*").split("[")[0])
    assert fields == [f[0] for f in GemmArgs._fields_]
    assert C.sizeof(GemmArgs) % 8 == 0


def test_ff_args_struct_matches_header():
    from stableanimator_b200._lib import FfArgs
    hdr = (ROOT / "include" / "sa_b200.h").read_text()
    body = hdr[hdr.index("typedef struct sa_ff_args {"):hdr.index("} sa_ff_args;")]
    fields = []
    for ln in body.splitlines()[1:]:
        ln = ln.split("/*")[0].strip().rstrip(";")
        if not ln:
            continue
        decl = ln.split()
        names = " ".join(decl[1:]) if decl[0] != "const" else " ".join(decl[2:])
        for nm in names.split(","):
            fields.append(nm.strip().lstrip("*").split("[")[0])
    assert fields == [f[0] for f in FfArgs._fields_]


def test_argument_errors_are_reported_not_thrown(lib):
    from stableanimator_b200._lib import GemmArgs
    a = GemmArgs()
    assert lib.sa_gemm_tap(C.byref(a), None) == -1
    assert b"null pointer" in lib.sa_last_error()
    assert lib.sa_attn_temporal(C.c_void_p(8), 192, 1, 17, 4, 64, 1, C.c_void_p(8), 64, 0, None) == -1
    assert b"F=17" in lib.sa_last_error()
    assert lib.sa_layernorm(C.c_void_p(8), 60, 4, 60, None, 1, None, 0, C.c_void_p(8), C.c_void_p(8), 1e-5,
                            C.c_void_p(8), 60, 0, None) == -1
    from stableanimator_b200._lib import FfArgs
    f = FfArgs()
    assert lib.sa_ff_fused(C.byref(f), None) == -1 and b"null pointer" in lib.sa_last_error()
    f.x = f.gamma = f.beta = f.w1 = f.b1 = f.w2 = f.b2 = f.out = 8
    f.c, f.m, f.ldx, f.ldc = 192, 4, 192, 192
    assert lib.sa_ff_fused(C.byref(f), None) == -1 and b"C=192" in lib.sa_last_error()
    assert lib.sa_gemm_pick_tile(65536, 320, 0) == 160
    assert lib.sa_gemm_pick_tile(65536, 2560, 1) in (128, 256)


def test_missing_library_raises(monkeypatch, tmp_path):
    from stableanimator_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "libsa_b200.so")
    with pytest.raises(_lib.SaLibraryError):
        _lib.lib()


def test_ctypes_signatures_match_header_prototypes():
    """Every entry of _lib.SIGNATURES has the argument count and kinds (pointer / int32 / int64 / float) of the
    prototype in include/sa_b200.h -- a drifting ctypes table would corrupt arguments silently."""
    from stableanimator_b200 import _lib
    hdr = (ROOT / "include" / "sa_b200.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    protos = dict(re.findall(r"\bint\s+(sa_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", hdr))
    kinds = {C.c_void_p: "ptr", C.c_int32: "i32", C.c_int64: "i64", C.c_float: "f32"}
    for name, argtypes in _lib.SIGNATURES.items():
        assert name in protos, name
        params = [p.strip() for p in protos[name].split(",") if p.strip() and p.strip() != "void"]
        want = []
        for p in params:
            if "*" in p:
                want.append("ptr")
            elif p.startswith("int64_t"):
                want.append("i64")
            elif p.startswith("int32_t"):
                want.append("i32")
            elif p.startswith("float"):
                want.append("f32")
            else:
                raise AssertionError(f"{name}: unhandled parameter {p!r}")
        got = [kinds.get(t, "ptr") for t in argtypes]   # POINTER(...) types count as pointers
        assert got == want, f"{name}: ctypes {got} vs header {want}"
