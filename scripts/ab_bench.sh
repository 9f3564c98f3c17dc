#!/bin/bash
# A/B on the GPU box: bench.py (device-timed legs only) for the default library and every
# prebuilt martini_b200/lib_var_*.so, on the workloads named in $WORKLOADS.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
show='import json,sys
d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["ms_per_step"],3), {k:round(v,3) for k,v in r["stage_ms"].items()}, "e2e", round(d["e2e"]["ms_per_step"],2), "frac %.4f"%r["frac"])'
for lib in default $(ls martini_b200/lib_var_*.so 2>/dev/null); do
  for w in ${WORKLOADS:-cfg2 cfg3 cfg4}; do
    echo "== $lib $w"
    if [ "$lib" = default ]; then unset MTN_B200_LIB; else export MTN_B200_LIB=$PWD/$lib; fi
    timeout 600 python bench.py --workload $w --others none --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline --no-class 2>&1 | tail -1 | python -c "$show"
  done
done 2>&1 | tee gpurun_out/ab_bench.log
