#!/bin/bash
# ncu --set full capture of the plan kernels (count + emit) of $1 -> gpurun_out/$2.ncu-rep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
W=${1:-cfg3}; OUT=${2:-r2_plan_$W}
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:plan_(count|emit)_kernel' -s 6 -c 2 \
  -o gpurun_out/$OUT -f python bench.py --workload $W --others none --steps 2 --warmup 3 --no-cpu-baseline --no-class > gpurun_out/$OUT.log 2>&1
python -c "import bench; print(bench.csrc_hash())" > gpurun_out/$OUT.hash
tail -3 gpurun_out/$OUT.log
