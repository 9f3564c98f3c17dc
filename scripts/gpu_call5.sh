#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zfuzz.py -x -q -m gpu -p no:cacheprovider --durations=5 > gpurun_out/r2_pytest_streams.log 2>&1
tail -12 gpurun_out/r2_pytest_streams.log
STEPS=5 bash scripts/ab_bench.sh
