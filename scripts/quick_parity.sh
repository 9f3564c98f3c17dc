#!/bin/bash
# GPU parity of one prebuilt variant library (array-level suite through the C ABI), then, time
# permitting, a config-4 (wide Gaussian footprints) A/B of base against gauss_sep.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
V=${VARIANT:-footrec2}
cp martini_b200/libmartini_b200.so /tmp/orig.so
cp "martini_b200/lib_var_${V}.so" martini_b200/libmartini_b200.so
timeout ${PT:-70} python -u -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider > gpurun_out/${V}_parity.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${V}_parity.log
for name in ${AB:-}; do
  cp "martini_b200/lib_var_${name}.so" martini_b200/libmartini_b200.so
  timeout 40 python bench.py --workload cfg4 --particles 300000 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > /tmp/line.json
  python - "$name" <<'PY' | tee -a gpurun_out/variants_cfg4.log
import json, sys
try:
    d = json.loads(open("/tmp/line.json").read())
    print(sys.argv[1], d["ms_per_step"], d["roofline"]["stage_ms"])
except Exception as e:
    print(sys.argv[1], "failed", e, open("/tmp/line.json").read()[-300:])
PY
done
cp /tmp/orig.so martini_b200/libmartini_b200.so
