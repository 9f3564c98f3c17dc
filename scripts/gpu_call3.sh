#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider -k "not cfg4_full" > gpurun_out/r2_pytest_kB.log 2>&1
tail -5 gpurun_out/r2_pytest_kB.log
bash scripts/ab_bench.sh
