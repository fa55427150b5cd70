#!/bin/bash
# tests + bench with and without the side stream for weight-gradient GEMMs + a launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
summ() { grep '"metric"' $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['eval_forward'])
"; }
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_r1c.log 2>&1; echo "bench exit=$?"; summ gpurun_out/bench_r1c.log
MAGB200_FUSED_OPTIMIZER=0 timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_r1c_nofuse.log 2>&1; echo "bench(separate AdamW) exit=$?"; summ gpurun_out/bench_r1c_nofuse.log
MAGB200_FUSED_OPTIMIZER=0 MAGB200_SIDE_STREAM=0 timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_r1c_noside.log 2>&1; echo "bench(no side stream, separate AdamW) exit=$?"; summ gpurun_out/bench_r1c_noside.log
NCU=/usr/local/cuda/bin/ncu
MAGB200_FUSED_OPTIMIZER=0 MAGB200_SIDE_STREAM=0 timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none -s 980 -c 245 --csv \
   --log-file gpurun_out/launches_v4.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --profile-mode \
   > gpurun_out/ncu_launches.log 2>&1; echo "launch-list exit=$?"
python tools/launch_summary.py gpurun_out/launches_v4.csv | head -24
