"""The C-ABI library loads and exports every symbol include/sdtree.h declares (no compute
calls: there is no GPU in the CPU suite)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "sdtree.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sdt_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_by_cuda_library():
    from practical_path_guiding_lab_b200.build import build
    from practical_path_guiding_lab_b200 import _lib
    path = build()                                   # nvcc cross-compiles sm_100a without a GPU
    lib = ctypes.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 28
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sdtree.h but not exported"
    assert set(names) == set(_lib.SYMBOLS), "ctypes table and header disagree"


def test_sass_is_sm100a():
    import subprocess
    from practical_path_guiding_lab_b200.build import build
    out = subprocess.run(["cuobjdump", "-lelf", build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_missing_library_fails_loudly(tmp_path):
    import pytest
    from practical_path_guiding_lab_b200 import _lib
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load_library(str(tmp_path / "libsdtree.so"))
