#!/bin/bash
# The round's evidence run, part 2 (one B200; the GPU suite of the same sources ran in
# scripts/gpu_profiles.sh: profiles/r2_pytest_gpu_final.txt): the default bench line, launch
# lists, ncu --set full captures of the kernels that dominate each config -- summarised ON THE
# BOX (the reports are 12 MB each and gpurun brings back 64 MiB), two reports kept.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
export NCU_SUMMARY_OUT=$O/ncu_summaries_new.json
timeout 900 python bench.py > $O/r2_bench_1gpu.json 2> $O/r2_bench_1gpu.err; tail -c 300 $O/r2_bench_1gpu.json; tail -2 $O/r2_bench_1gpu.err
for w in cfg3 cfg2 cfg4; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_$w.csv \
    python bench.py --workload $w --others none --steps 1 --warmup 3 --no-cpu-baseline --no-class > $O/r2_launches_$w.log 2>&1
done
cap() {  # workload, name, kernel regex, mangled-name fragment for the per-line table
  bash scripts/gpu_ncu.sh $1 $2 $3 > /dev/null
  python profiles/summarize_ncu.py $O/$2.ncu-rep $O/$2.md > /dev/null 2>&1
  python profiles/make_ncu_summary.py $O/$2.ncu-rep $1 > /dev/null 2>&1
  python profiles/sass_by_line.py $O/$2.ncu-rep martini_b200/libmartini_b200.so 1.0 $4 > $O/$2.lines.txt 2>&1
  [ "$5" = keep ] || rm -f $O/$2.ncu-rep
  head -4 $O/$2.md | tail -2
}
cap cfg3 r2_cfg3_column_kernel column_kernel column_kernelILb0 keep
cap cfg3 r2_cfg3_project_kernel project_kernel project_kernelILb0ELi2E
cap cfg2 r2_cfg2_project_kernel project_kernel project_kernelILb0ELi0E keep
cap cfg4 r2_cfg4_splat_kernel splat_kernel splat_kernelILb0
cap cfg2c6 r2_cfg2c6_project_kernel project_kernel project_kernelILb0ELi1E
timeout 300 python scripts/bench_convolve.py --out $O/r2_convolve.json | cut -c1-300
du -sh $O
