#!/bin/bash
# Shortest possible A/B of the prebuilt variant libraries: one bench.py run each, results
# appended to gpurun_out/variants.log as they arrive (so a cut-off call still returns some).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cp martini_b200/libmartini_b200.so /tmp/orig.so
for name in ${VARIANTS:-base footrec2 footrec gauss_sep}; do
  cp "martini_b200/lib_var_${name}.so" martini_b200/libmartini_b200.so
  timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > /tmp/line.json
  python - "$name" <<'PY' | tee -a gpurun_out/variants.log
import json, sys
try:
    d = json.loads(open("/tmp/line.json").read())
    print(sys.argv[1], d["ms_per_step"], d["roofline"]["stage_ms"], d["e2e"]["value"])
except Exception as e:
    print(sys.argv[1], "failed", e, open("/tmp/line.json").read()[-300:])
PY
done
cp /tmp/orig.so martini_b200/libmartini_b200.so
