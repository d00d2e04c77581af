#!/bin/bash
# Multi-GPU evidence for a tag on N GPUs of one box:
#   gpurun --gpus N --timeout 1500 -- 'bash profiles/multi.sh r02x N'
set -u
TAG=${1:?tag}; N=${2:?gpus}
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$N" = "2" ]; then
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_handles" > $OUT/${TAG}_two_devices.log 2>&1; echo "two-device test rc=$?"; tail -1 $OUT/${TAG}_two_devices.log
fi
$TR bench.py --gpus $N > $OUT/${TAG}_bench_n$N.log 2>&1; echo "bench n=$N rc=$?"
$TR bench.py --gpus $N --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref_n$N.log 2>&1; echo "reference arm n=$N rc=$?"
G=tests/golden/cornell_box_tungsten_256.npy
D="-m practical_path_guiding_lab_b200.driver --res 1024 --budget 1020 --max-depth 13 --ground-truth $G"
for SH in ${SHARDS:-auto}; do
  $TR $D --shard $SH > $OUT/${TAG}_cornell1024_${SH}_n$N.log 2>&1; echo "cornell $SH n=$N rc=$?"; tail -1 $OUT/${TAG}_cornell1024_${SH}_n$N.log | cut -c1-200
done
