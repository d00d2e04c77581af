#!/bin/bash
# TEST INFRASTRUCTURE.  compute-sanitizer is closed on the GPU pool, so the same kernel sources are compiled for the host
# (-DSDT_HOSTEMU: kernels as serial loops) with AddressSanitizer + UBSan and every parity case of tests/sdt_cases.py, the
# randomised suite cases and SEEDS more randomised differential seeds are run through the C ABI:
#   tools/hostemu_sanitize.sh [SEEDS]         (default 120; ~4 minutes)
# Covers indexing / arena / table logic of the lane functors and the refine, not device-only code (the single-pass scan
# kernel, warp-level primitives) and not races -- those are held by the bit-exact GPU cases.
set -eu
cd "$(dirname "$0")/.."
SEEDS=${1:-120}
OUT=/tmp/libsdtree_hostemu_asan.so
g++ -std=c++17 -O1 -g -ffp-contract=off -fPIC -shared -fsanitize=address,undefined -fno-omit-frame-pointer -DSDT_HOSTEMU \
    -x c++ practical_path_guiding_lab_b200/csrc/sdtree.cu -o $OUT
cat > /tmp/run_hostemu_asan.py <<PY
import sys
sys.path.insert(0, '$(pwd)'); sys.path.insert(0, '$(pwd)/tests')
import sdt_cases as cases, fuzz_cases
from practical_path_guiding_lab_b200 import SDTree
ctx = cases.Ctx(make=lambda **kw: SDTree(lib_path='$OUT', **kw))
bad = []
for c in cases.ALL_CASES + fuzz_cases.SUITE_CASES:
    try:
        c(ctx); print('ok', c.__name__, flush=True)
    except Exception as e:
        bad.append((c.__name__, repr(e)[:200]))
for seed in range(3000, 3000 + $SEEDS):
    try:
        fuzz_cases.fuzz_one(ctx, seed)
    except Exception as e:
        bad.append((seed, repr(e)[:200]))
print('failures:', bad)
sys.exit(1 if bad else 0)
PY
ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=halt_on_error=1:print_stacktrace=1 \
LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) python /tmp/run_hostemu_asan.py
