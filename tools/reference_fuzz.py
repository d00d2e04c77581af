#!/usr/bin/env python
"""TEST INFRASTRUCTURE (needs /root/reference; not runnable on the GPU box).  Randomised differential seeds with the
REFERENCE ITSELF as the expected-value source: src/common.py, quadtree.py, kdtree.py, path_guiding_integrator.py run
unmodified on the numpy Dr.Jit / Mitsuba stand-ins (oracle/refshim), against the shipped kernel sources compiled for the
host (tests/hostemu).  Same check as tests/test_reference_on_shim.py::test_fuzz_seed_against_reference, more seeds:
    python tools/reference_fuzz.py --seeds 300 --start 2000
"""
import argparse
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fuzz_cases  # noqa: E402
import sdt_cases as cases  # noqa: E402
from hostemu.build_hostemu import build as build_hostemu  # noqa: E402
from oracle import refshim  # noqa: E402
from practical_path_guiding_lab_b200 import SDTree  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seeds", type=int, default=100)
ap.add_argument("--start", type=int, default=2000)
a = ap.parse_args()
if not refshim.available():
    raise SystemExit("the reference tree is not here")
from oracle.refshim import as_oracle as ro  # noqa: E402
cases.so = ro
fuzz_cases.so = ro
lib = build_hostemu()
ctx = cases.Ctx(make=lambda **kw: SDTree(lib_path=lib, **kw))
bad = []
for seed in range(a.start, a.start + a.seeds):
    try:
        fuzz_cases.fuzz_one(ctx, seed)
    except Exception:
        traceback.print_exc()
        bad.append(seed)
    if (seed - a.start) % 25 == 24:
        print("seed", seed, "failed so far:", bad, flush=True)
print("reference-on-shim as expected values, seeds %d..%d: failed seeds: %s" % (a.start, a.start + a.seeds - 1, bad))
sys.exit(1 if bad else 0)
