#!/bin/bash
# A/B: default projection kernel against the warp-specialised one (MTN_PROJECT=ws).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
show='import json,sys
d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["stage_ms"]["project"], d["e2e"]["ms_per_step"])'
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4
for mode in ${MODES:-classic ws classic}; do
  echo "== $mode"
  MTN_PROJECT=$mode timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "$show"
done
