#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import bench; print(bench.csrc_hash())"
python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -1
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/r2_pytest_gpu_final.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu_final.log
