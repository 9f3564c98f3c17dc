#!/bin/bash
# parity (incl. full-size configs) of the new sort / plan kernels / brick tile, then A/B against
# HEAD~2 (lib_var_old.so) on configs 2, 3, 4
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zfuzz.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
WORKLOADS="cfg2 cfg3 cfg4" STEPS=5 bash scripts/ab_bench.sh
