#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/r2_bench_v2.json 2> gpurun_out/r2_bench_v2.err
tail -c 600 gpurun_out/r2_bench_v2.json; tail -3 gpurun_out/r2_bench_v2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_cfg3.csv python bench.py --workload cfg3 --others none --steps 2 --warmup 3 --no-cpu-baseline --no-class > gpurun_out/r2_launches_cfg3.log 2>&1
tail -2 gpurun_out/r2_launches_cfg3.log | cut -c1-300
