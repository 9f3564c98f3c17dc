#!/bin/bash
cd "$(dirname "$0")/.."
bash scripts/gpu_ncu.sh cfg2 r2b_cfg2_project_kernel project_kernel
bash scripts/gpu_ncu.sh cfg3 r2b_cfg3_column_kernel column_kernel
