#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu -p no:cacheprovider --durations=5 > gpurun_out/r2_pytest_all.log 2>&1
tail -12 gpurun_out/r2_pytest_all.log
for w in cfg3 cfg2 cfg4; do
timeout 600 python bench.py --workload $w --others none --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c '
import json,sys
d=json.loads(sys.stdin.read()); r=d["roofline"]; print(sys.argv[1], round(d["ms_per_step"],3), {k:round(v,3) for k,v in r["stage_ms"].items()}, "e2e", round(d["e2e"]["ms_per_step"],2), "class", d.get("martini_class_wall_ms"), d.get("martini_class_to_host_ms"), d.get("martini_class_error"), "frac %.4f full %.4f"%(r["frac"], r["frac_full"]))' $w
done 2>&1 | tee gpurun_out/r2_call7_bench.log
