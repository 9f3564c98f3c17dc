#!/bin/bash
# one more capture for the current sources: the config-4 splat kernel (summarised on the box)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/ncu_summaries_new.json
bash scripts/gpu_ncu.sh cfg4 r2_cfg4_splat_kernel splat_kernel > /dev/null
python profiles/summarize_ncu.py $O/r2_cfg4_splat_kernel.ncu-rep $O/r2_cfg4_splat_kernel.md > /dev/null 2>&1
NCU_SUMMARY_OUT=$O/ncu_summaries_new.json python profiles/make_ncu_summary.py $O/r2_cfg4_splat_kernel.ncu-rep cfg4 > /dev/null 2>&1
python profiles/sass_by_line.py $O/r2_cfg4_splat_kernel.ncu-rep martini_b200/libmartini_b200.so 1.0 splat_kernelILb0 > $O/r2_cfg4_splat_kernel.lines.txt 2>&1
rm -f $O/r2_cfg4_splat_kernel.ncu-rep
head -4 $O/r2_cfg4_splat_kernel.md | tail -2
