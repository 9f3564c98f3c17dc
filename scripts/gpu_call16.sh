#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f martini_b200/lib_var_old.so
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider -k "not full_size" 2>&1 | tail -3
WORKLOADS="cfg2 cfg3 cfg4" STEPS=5 bash scripts/ab_bench.sh
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_cfg3.csv \
    python bench.py --workload cfg3 --others none --steps 1 --warmup 3 --no-cpu-baseline --no-class > gpurun_out/r2b_launches_cfg3.log 2>&1
