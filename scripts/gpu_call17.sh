#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zfuzz.py -x -q -m gpu -p no:cacheprovider -k "not cfg4 and not cfg2_full" 2>&1 | tail -3
WORKLOADS="cfg3" STEPS=10 bash scripts/ab_bench.sh
