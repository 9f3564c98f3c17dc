#!/bin/bash
# The round's evidence run (one B200): full GPU suite, the default bench line, launch lists,
# ncu --set full captures of the kernels that dominate each config (with source-hash sidecars),
# the beam-convolution timing and the full CPU run of config 2.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -1
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > $O/r2_pytest_gpu_final.log 2>&1; tail -3 $O/r2_pytest_gpu_final.log
timeout 900 python bench.py > $O/r2_bench_1gpu.json 2> $O/r2_bench_1gpu.err; tail -c 300 $O/r2_bench_1gpu.json; tail -2 $O/r2_bench_1gpu.err
for w in cfg3 cfg2 cfg4; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_$w.csv \
    python bench.py --workload $w --others none --steps 1 --warmup 3 --no-cpu-baseline --no-class > $O/r2_launches_$w.log 2>&1
done
bash scripts/gpu_ncu.sh cfg3 r2_cfg3_column_kernel column_kernel
bash scripts/gpu_ncu.sh cfg3 r2_cfg3_project_kernel project_kernel
bash scripts/gpu_ncu.sh cfg2 r2_cfg2_project_kernel project_kernel
bash scripts/gpu_ncu.sh cfg4 r2_cfg4_splat_kernel splat_kernel
bash scripts/gpu_ncu.sh cfg2c6 r2_cfg2c6_project_kernel project_kernel
timeout 300 python scripts/bench_convolve.py --out $O/r2_convolve.json | cut -c1-600
