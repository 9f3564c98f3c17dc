#!/bin/bash
# Experiment helper: time bench.py with each prebuilt variant library (martini_b200/lib_var_*.so).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cp martini_b200/libmartini_b200.so /tmp/orig.so
for f in martini_b200/lib_var_*.so; do
  cp "$f" martini_b200/libmartini_b200.so
  echo "== $f"
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['stage_ms'], d['e2e']['value'])"
  case "$f" in *exp*|*base*) ;; *) timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -1;; esac
done
cp /tmp/orig.so martini_b200/libmartini_b200.so
