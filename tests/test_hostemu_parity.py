"""CPU suite: the kernels' per-lane / per-item logic (compiled by g++ with -DSDT_HOSTEMU,
tests/hostemu/build_hostemu.py -- test infrastructure, never loaded by the package) held
against the oracle through the same C ABI the GPU suite uses.  The GPU suite
(tests/test_gpu_parity.py) runs the very same cases on libsdtree.so."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sdt_cases as cases  # noqa: E402
import fuzz_cases  # noqa: E402
from hostemu.build_hostemu import build as build_hostemu  # noqa: E402

from practical_path_guiding_lab_b200 import SDTree  # noqa: E402


@pytest.fixture(scope="module")
def ctx():
    lib = build_hostemu()
    return cases.Ctx(make=lambda **kw: SDTree(lib_path=lib, **kw))


@pytest.mark.parametrize("case", cases.ALL_CASES + fuzz_cases.SUITE_CASES, ids=lambda c: c.__name__)
def test_case(ctx, case):
    case(ctx)


def test_npz_roundtrip(ctx, tmp_path):
    cases.case_npz_roundtrip(ctx, tmp_path)
