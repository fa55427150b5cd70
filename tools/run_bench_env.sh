#!/bin/bash
# usage: VARIANTS="A=1 B=2|A=0" bash tools/run_bench_env.sh   (variants separated by '|', each a space-separated env list)
mkdir -p gpurun_out
summ() { grep '"metric"' $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['eval_forward'])
"; }
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_xlnet.py -x -q 2>&1 | tail -3 | cut -c1-300
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "tests failed: skipping the benches"; exit 1; fi
IFS='|' read -ra VS <<< "${VARIANTS:-X=0}"
i=0
for v in "${VS[@]}"; do
  i=$((i+1))
  env $v timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_v$i.log 2>&1; rc=$?; echo "bench [$v] exit=$rc"
  if [ $rc -ne 0 ]; then tail -5 gpurun_out/bench_v$i.log | cut -c1-300; exit 1; fi
  summ gpurun_out/bench_v$i.log
done
