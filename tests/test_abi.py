"""The C-ABI library loads and exports every symbol include/*.h declares (no compute calls)."""
import ctypes
import os
import re

from seevcn_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "seevcn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(seevcn_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    syms = declared_symbols()
    assert len(syms) >= 20
    L = ctypes.CDLL(_abi.LIB_PATH)
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/seevcn_b200.h but not exported"
        assert s in _abi.SIGNATURES, f"{s} has no ctypes signature in _abi.py"
    assert set(_abi.SIGNATURES) == set(syms)


def test_ctypes_signatures_match_header_arity():
    """every ctypes signature has as many arguments as the prototype in the header (a mismatch corrupts the call silently)"""
    text = open(os.path.join(ROOT, "include", "seevcn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    protos = re.findall(r"\b(seevcn_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S)
    assert len(protos) >= 20
    seen = set()
    for name, args in protos:
        if name in seen:
            continue
        seen.add(name)
        args = args.strip()
        n = 0 if args in ("", "void") else args.count(",") + 1
        assert n == len(_abi.SIGNATURES[name][1]), f"{name}: header has {n} parameters, _abi.py binds {len(_abi.SIGNATURES[name][1])}"
    assert seen == set(_abi.SIGNATURES)


def test_version_and_error_string():
    L = _abi.lib()
    assert L.seevcn_abi_version() == 2
    # argument validation happens before any CUDA call: safe without a GPU
    rc = L.seevcn_knn(1, 8, 8, 100, None, None, None, None, None)
    assert rc == 1 and b"outside [1,64]" in L.seevcn_last_error()
    rc = L.seevcn_points_in_boxes(-1, 0, 0, None, None, None, None)
    assert rc == 1
    assert L.seevcn_crop_workspace_bytes(2, 50, 180000) >= 2 * 176 * 50 * 4


def test_no_oracle_import_in_product():
    """The product package must never route through the oracle."""
    pkg = os.path.join(ROOT, "see-vcn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)
                assert "liboracle" not in src, os.path.join(dirpath, f)
