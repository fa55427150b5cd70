#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short 2>&1 | tail -8
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_tmp.log 2>&1; echo "bench exit=$?"
tail -1 gpurun_out/bench_tmp.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['eval_forward'])
for k in d['roofline']['per_shape']: print(k)
"
NCU=/usr/local/cuda/bin/ncu
timeout 900 $NCU --metrics gpu__time_duration.sum --clock-control none -s 1032 -c 258 --csv \
   --log-file gpurun_out/launches_tmp.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --profile-mode \
   > gpurun_out/ncu_launches.log 2>&1; echo "launch-list exit=$?"
