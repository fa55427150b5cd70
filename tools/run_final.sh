#!/bin/bash
# round-end evidence: full GPU test suite, smoke, bench (with CPU baseline), ncu launch list and one --set full capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_final.log 2>&1; echo "bench exit=$?"
grep '"metric"' gpurun_out/bench_final.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['eval_forward'], d['cpu_baseline'])
print(d['roofline'])
"
NCU=/usr/local/cuda/bin/ncu
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none -s 984 -c 246 --csv \
   --log-file gpurun_out/launches_v5.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --profile-mode \
   > gpurun_out/ncu_launches.log 2>&1; echo "launch-list exit=$?"
python tools/launch_summary.py gpurun_out/launches_v5.csv | head -26
timeout 600 $NCU --set full --clock-control none --import-source on -k regex:"gemm2?p?_tcgen05" -s 300 -c 12 \
   -o gpurun_out/prof_gemm_r1c -f python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --profile-mode \
   > gpurun_out/ncu_gemm.log 2>&1; echo "gemm capture exit=$?"
