#!/bin/bash
# bench.py under torchrun on N GPUs as the driver launches it; the JSON line goes to
# gpurun_out/r2_bench_${N}gpu.json:   N=8 bash scripts/bench_n_json.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-2}
timeout ${TMO:-900} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 $EXTRA > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
tail -c 1500 gpurun_out/r2_bench_${N}gpu.json; tail -3 gpurun_out/r2_bench_${N}gpu.err | cut -c1-300
