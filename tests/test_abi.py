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


def test_stage_kernel_register_budget():
    """The one-quadrature-point stage kernels are built with launch bounds (160 threads, 3 blocks) so that explosion_multi's
    150-column blocks fit one strip, but the large-problem shape is 4 thread blocks of 128 threads per SM: that needs <= 128
    registers per thread, which nvcc's -maxrregcount cannot enforce on a kernel with launch bounds.  ptxas stays within 128
    today; this test fails the build the day an edit pushes it over (occupancy would silently drop to 3 blocks)."""
    import shutil
    import subprocess

    from pyhype_b200 import _lib

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump) or not os.path.exists(_lib.LIB_PATH):
        pytest.skip("cuobjdump or the library not available")
    out = subprocess.run([cuobjdump, "-res-usage", _lib.LIB_PATH], capture_output=True, text=True).stdout
    lines = out.splitlines()
    seen = 0
    for i, ln in enumerate(lines):
        m = re.search(r"Function _ZN3pyh13k_stage_marchILi(\d)ELi(\d)ELi(\d)ELi(\d)EE", ln)
        if not m or m.group(4) != "1":
            continue
        regs = int(re.search(r"REG:(\d+)", lines[i + 1]).group(1))
        seen += 1
        assert regs <= 128, f"k_stage_march<{m.group(1)},{m.group(2)},{m.group(3)},1> uses {regs} registers: 4 x 128 threads no longer fit an SM"
    assert seen == 24, seen
