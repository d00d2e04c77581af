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


def test_documented_tuning_keys_are_the_accepted_ones():
    """the key list in the sdt_set_tuning comment of include/sdtree.h == the keys the library accepts (checked on the
    host emulation build of the same sources: sdt_set_tuning computes nothing)"""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from hostemu.build_hostemu import build as build_hostemu
    from practical_path_guiding_lab_b200 import SDTree, SDTreeError
    hdr = open(os.path.join(ROOT, "include", "sdtree.h")).read()
    doc = hdr[hdr.index("tuning / introspection"):hdr.index("int sdt_set_tuning")]
    documented = set(re.findall(r'"([a-z_]+)"', doc))
    src = open(os.path.join(ROOT, "practical_path_guiding_lab_b200", "csrc", "sdt_io.inl")).read()
    accepted = set(re.findall(r'k == "([a-z_]+)"', src))
    assert documented == accepted, (documented ^ accepted)
    t = SDTree(lib_path=build_hostemu(), kd_capacity=16, quad_capacity=64)
    defaults = dict(query_block=768, splat_block=768, host_chunk=1 << 20)
    for k in sorted(accepted):
        t.set_tuning(k, defaults.get(k, 1))
    try:
        t.set_tuning("no_such_key", 1)
        assert False
    except SDTreeError:
        pass
