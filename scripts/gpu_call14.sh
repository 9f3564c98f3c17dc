#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider -k "not full_size" 2>&1 | tail -3
WORKLOADS="cfg2 cfg3" STEPS=5 bash scripts/ab_bench.sh
