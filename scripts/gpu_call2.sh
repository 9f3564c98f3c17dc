#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider -k "full_size or cube_size" --durations=8 > $O/r2_pytest_fullsize.log 2>&1
tail -15 $O/r2_pytest_fullsize.log
nproc
timeout 900 python bench.py > $O/r2_bench_new.json 2> $O/r2_bench_new.err
tail -c 3000 $O/r2_bench_new.json; tail -5 $O/r2_bench_new.err
