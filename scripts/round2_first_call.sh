#!/bin/bash
# First GPU call of round 2 (one B200, ~10 min): everything round 1 queued once its GPU budget
# was spent.  Build the variant libraries first, here in the container:
#   scripts/build_variants.sh base= gauss_sep=-DMTN_GAUSS_SEP=1 wtab_more0=-DMTN_WTAB_MORE=0
# then:  gpurun --timeout 900 -- 'bash scripts/round2_first_call.sh'
# Output: gpurun_out/r2_*  (copy what is to be judged into profiles/).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
show='import json,sys
d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["stage_ms"], d["e2e"]["ms_per_step"])'

# 1. parity of the default library (the class-level suite has not run on a GPU since the
#    footprint record became the default, nor have the Wendland C6 / quartic tables)
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > $O/r2_pytest_gpu.log 2>&1
tail -3 $O/r2_pytest_gpu.log

# 2. bench line + launch list + full capture of the projection kernel (v11)
timeout 300 python bench.py > $O/r2_bench_cfg2.json 2> $O/r2_bench_cfg2.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $O/r2_launches_ncu_raw.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2_ncu_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:project_kernel -c 1 -s 4 \
  -o $O/r2_project_kernel python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2_ncu_full.log 2>&1

# 3. the BASELINE configs that were only parity-tested at reduced size: full-size timings
for w in cfg3 cfg4; do
  echo "== $w"
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > $O/r2_bench_$w.json 2> $O/r2_bench_$w.err
  tail -1 $O/r2_bench_$w.json | python -c "$show" || tail -3 $O/r2_bench_$w.err
done

# 4. queued A/Bs: separable Gaussian kernel integrals on config 4, closed forms against tables
#    for Wendland C6 (config-2 geometry with the C6 kernel)
cp martini_b200/libmartini_b200.so /tmp/orig.so
for v in gauss_sep wtab_more0; do
  [ -f martini_b200/lib_var_$v.so ] || continue
  cp martini_b200/lib_var_$v.so martini_b200/libmartini_b200.so
  for w in cfg4 cfg2c6; do
    echo "== $v $w"
    timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "$show"
  done
done 2>&1 | tee $O/r2_variants.log
cp /tmp/orig.so martini_b200/libmartini_b200.so
