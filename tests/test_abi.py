"""CPU: the C-ABI library builds, loads, and exports exactly what include/b200seg.h declares."""
import ctypes
import os
import re

import pytest

import b200seg
from b200seg import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_protos():
    src = open(os.path.join(ROOT, "include", "b200seg.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|size_t|const char\*)\s+(b200seg_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        protos[m.group(1)] = n
    return protos


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        from b200seg.build import build
        build()
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    protos = _header_protos()
    assert len(protos) >= 18
    for name in protos:
        assert hasattr(lib, name), "header declares %s but the library does not export it" % name


def test_binding_matches_header_arity(lib):
    protos = _header_protos()
    assert set(protos) == set(_lib._PROTOS), set(protos) ^ set(_lib._PROTOS)
    for name, n in protos.items():
        assert len(_lib._PROTOS[name]) == n, "%s: header has %d params, binding %d" % (name, n, len(_lib._PROTOS[name]))


def test_version_and_strerror(lib):
    assert lib.b200seg_version() == 100
    assert lib.b200seg_strerror(0) == b"ok"
    assert b"fallback" in lib.b200seg_strerror(-4)


def test_no_torch_in_abi():
    """The boundary is plain C: the .so must not link against libtorch / libc10."""
    import subprocess
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    # match library NAMES only: ldd also prints load addresses, whose hex digits can spell "c10"
    assert not re.search(r"lib(torch|c10)[^ ]*\.so", out), out
