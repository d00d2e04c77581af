#!/bin/bash
# One GPU-box pass that produces everything profiles/make_summary.py needs for a tag:
#   gpurun --timeout 1500 -- 'bash profiles/capture.sh r02a'
# then, back in the container:
#   bash profiles/export.sh r02a
# Order matters: the plain bench first (its numbers are the reported ones), the ncu passes after it
# (numbers printed under ncu are never bench values).  One GPU only.
set -u
TAG=${1:?tag}
OUT=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -1 $OUT/${TAG}_gpu_tests.log
python bench.py > $OUT/${TAG}_bench_n1.log 2>&1; echo "bench rc=$?"
python bench.py --impl reference > $OUT/${TAG}_bench_ref.log 2>&1; echo "reference arm rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/${TAG}_launches.csv $B > $OUT/${TAG}_ncu1.log 2>&1; echo "ncu launch list rc=$?"
# 6 tree-build splats + 3 warm-up steps x 2 kernels precede the timed steps: skip 12, capture the step's two launches
# (fused sample + pdf, splat); the separate sample / pdf kernels are captured by name from their own timing loop
ncu --set full --clock-control none --import-source on -k regex:k_wavefront -s 12 -c 2 -o $OUT/prof_${TAG} $B > $OUT/${TAG}_ncu2.log 2>&1; echo "ncu full (step) rc=$?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_wavefront<(SampleLane<|PdfLane,)" -c 2 -o $OUT/prof_${TAG}_sep $B > $OUT/${TAG}_ncu3.log 2>&1; echo "ncu full (separate) rc=$?"
