#!/bin/bash
# A/B: warp-specialised projection kernel (default) against the classic one (MTN_PROJECT=classic).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
show='import json,sys
d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["stage_ms"]["project"], d["e2e"]["ms_per_step"])'
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4
for mode in ws classic ws; do
  echo "== $mode"
  MTN_PROJECT=$mode timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "$show"
done
