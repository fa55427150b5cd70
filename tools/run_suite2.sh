#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_r1b.log 2>&1; echo "bench exit=$?"
grep '"metric"' gpurun_out/bench_r1b.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['eval_forward'], d['cpu_baseline'])
for k in d['roofline']['per_shape']: print(k)
"
NCU=/usr/local/cuda/bin/ncu
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none -s 980 -c 245 --csv \
   --log-file gpurun_out/launches_v3.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --profile-mode \
   > gpurun_out/ncu_launches.log 2>&1; echo "launch-list exit=$?"
timeout 600 $NCU --set full --clock-control none --import-source on -k regex:"gemm2?_tcgen05" -s 300 -c 10 \
   -o gpurun_out/prof_gemm_r1b -f python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --profile-mode \
   > gpurun_out/ncu_gemm.log 2>&1; echo "gemm capture exit=$?"
timeout 600 $NCU --set full --clock-control none --import-source on \
   -k regex:"mag_tail_fwd|sdpa_fwd_mma|sdpa_bwd_mma|embed_ln_fwd|adamw" -s 20 -c 6 \
   -o gpurun_out/prof_mag_sdpa_r1b -f python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --profile-mode \
   > gpurun_out/ncu_rows.log 2>&1; echo "mag/sdpa capture exit=$?"
