#!/bin/bash
# Turns the files profiles/capture.sh left in gpurun_out/ into the committed evidence of a tag.
set -eu
TAG=${1:?tag}
cd "$(dirname "$0")/.."
ncu -i gpurun_out/prof_${TAG}.ncu-rep --page raw --csv > profiles/${TAG}_ncu_full_wavefront_raw.csv 2>/dev/null
if [ -f gpurun_out/prof_${TAG}_sep.ncu-rep ]; then ncu -i gpurun_out/prof_${TAG}_sep.ncu-rep --page raw --csv > profiles/${TAG}_ncu_full_separate_raw.csv 2>/dev/null; fi
# SASS of the shipped library's hot loops: instruction mix per wavefront kernel
cuobjdump -sass practical_path_guiding_lab_b200/libsdtree.so 2>/dev/null | python profiles/sass_counts.py > profiles/${TAG}_sass_instcounts.txt
cp gpurun_out/${TAG}_launches.csv profiles/${TAG}_launches_bench_steps2.csv
grep '^{' gpurun_out/${TAG}_bench_n1.log | tail -1 > profiles/${TAG}_bench_n1.json
grep '^{' gpurun_out/${TAG}_bench_ref.log | tail -1 > profiles/${TAG}_bench_reference_arm.json
cp gpurun_out/${TAG}_gpu_tests.log profiles/${TAG}_gpu_tests.log
python profiles/make_summary.py ${TAG}
sed -n 1,12p profiles/${TAG}_summary.md
