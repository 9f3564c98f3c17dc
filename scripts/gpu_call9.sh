#!/bin/bash
# A/B of the 4 px x 8 ch register tile (default lib) against HEAD (lib_var_old.so), the brick
# parity tests, then the ncu capture of the plan kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider -k "not full_size" 2>&1 | tail -3
WORKLOADS="cfg2 cfg3" STEPS=5 bash scripts/ab_bench.sh
bash scripts/gpu_ncu_plan.sh cfg3 r2_plan_cfg3
