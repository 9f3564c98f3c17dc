#!/bin/bash
# End-to-end leg of bench.py for several sub-slab counts (read-back overlapped with projection).
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "to_host" 2>&1 | tail -2
for s in ${SLABS:-1 2 3 4}; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-slabs $s 2>&1 | tail -1 | python -c "
import json, sys
d = json.loads(sys.stdin.read())
print('slabs', $s, 'resident ms', round(d['ms_per_step'], 3), 'e2e ms', round(d['e2e']['ms_per_step'], 3), 'e2e updates/s %.3e' % d['e2e']['value'])"
done
