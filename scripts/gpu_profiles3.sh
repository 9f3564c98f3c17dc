#!/bin/bash
# Evidence for a change of the brick kernel only (one B200, ~3.5 min): ncu --set full captures of
# the config-3 column / brick kernels and the config-2 brick kernel, summarised on the box and
# merged into the box's profiles/ncu_summaries.json, then the default bench line (which reads
# roofline.traffic from there).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
cap() {  # workload, name, kernel regex, mangled-name fragment for the per-line table
  bash scripts/gpu_ncu.sh $1 $2 $3 > /dev/null
  python profiles/summarize_ncu.py $O/$2.ncu-rep $O/$2.md > /dev/null 2>&1
  NCU_SUMMARY_OUT=$O/ncu_summaries_new.json python profiles/make_ncu_summary.py $O/$2.ncu-rep $1 > /dev/null 2>&1
  python profiles/make_ncu_summary.py $O/$2.ncu-rep $1 > /dev/null 2>&1
  python profiles/sass_by_line.py $O/$2.ncu-rep martini_b200/libmartini_b200.so 1.0 $4 > $O/$2.lines.txt 2>&1
  rm -f $O/$2.ncu-rep
  head -4 $O/$2.md | tail -2
}
rm -f $O/ncu_summaries_new.json
cap cfg3 r2_cfg3_project_kernel project_kernel project_kernelILb0ELi2E
cap cfg2 r2_cfg2_project_kernel project_kernel project_kernelILb0ELi0E
cap cfg3 r2_cfg3_column_kernel column_kernel column_kernelILb0
timeout 200 python bench.py > $O/r2_bench_1gpu.json 2> $O/r2_bench_1gpu.err; tail -c 200 $O/r2_bench_1gpu.json; tail -2 $O/r2_bench_1gpu.err
