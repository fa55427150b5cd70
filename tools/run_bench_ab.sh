#!/bin/bash
# usage: VARIANTS="A=1|A=0" bash tools/run_bench_ab.sh   -- benches only (no tests), all variants on the same box
mkdir -p gpurun_out
summ() { grep '"metric"' $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['eval_forward']['ms_per_step'])
"; }
IFS='|' read -ra VS <<< "${VARIANTS:-X=0}"
i=0
for v in "${VS[@]}"; do
  i=$((i+1))
  env $v timeout 600 python bench.py --steps ${STEPS:-100} --warmup 10 --no-cpu-baseline ${BENCH_FLAGS:-} > gpurun_out/bench_ab$i.log 2>&1; rc=$?; echo "bench [$v] exit=$rc"
  if [ $rc -ne 0 ]; then tail -5 gpurun_out/bench_ab$i.log | cut -c1-300; continue; fi
  summ gpurun_out/bench_ab$i.log
done
