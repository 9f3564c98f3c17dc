#!/bin/bash
# bench.py under torchrun on N GPUs of this box, as the driver launches it:  N=2 bash scripts/bench_n.sh
cd "$(dirname "$0")/.."
N=${N:-2}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 2>&1 | tail -2 | python -c "
import json, sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln)
        print('N', d['n_gpus'], 'ms', round(d['ms_per_step'], 3), 'value %.3e' % d['value'], 'e2e ms', round(d['e2e']['ms_per_step'], 3), 'e2e %.3e' % d['e2e']['value'])
    else:
        print(ln.strip()[:300])"
