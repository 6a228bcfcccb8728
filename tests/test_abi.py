"""-m "not gpu": the C-ABI shared library loads and exports every symbol include/pyh_b200.h
declares (no compute calls), and the ctypes struct layouts match the header's."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pyh_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pyh_[a-z_0-9]+)\s*\(", src)))


def test_library_built_and_exports_header_symbols():
    from pyhype_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/pyh_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == syms, "ctypes binding and header disagree"
    lib.pyh_abi_version.restype = ctypes.c_int
    assert lib.pyh_abi_version() == _lib.PYH_ABI_VERSION


def test_struct_layouts_match_header():
    from pyhype_b200 import _lib

    # pyh_config: 10 int32 + 36 doubles + 2 doubles
    assert ctypes.sizeof(_lib.PyhConfig) == 10 * 4 + 38 * 8
    assert _lib.PyhConfig.tableau.offset == 40 and _lib.PyhConfig.gamma.offset == 40 + 36 * 8
    # pyh_block_desc: 14 int32 then 7 + 4 pointers
    assert _lib.PyhBlockDesc.nodes_x.offset == 56
    assert ctypes.sizeof(_lib.PyhBlockDesc) == 56 + 11 * 8


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from pyhype_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()
