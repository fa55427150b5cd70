#!/bin/bash
# ncu evidence for round 1: (1) launch list of one training step, (2) full-set capture of the GEMM, (3) of the row kernels
mkdir -p gpurun_out
NCU=/usr/local/cuda/bin/ncu
# one eager step = 258 launches; bench does 1 counted + 3 warm-up + 2 timed; skip the first 4 steps
timeout 900 $NCU --metrics gpu__time_duration.sum --clock-control none -s 1032 -c 258 --csv \
   --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --profile-mode \
   > gpurun_out/ncu_launches.log 2>&1; echo "launch-list exit=$?"
timeout 900 $NCU --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 300 -c 8 \
   -o gpurun_out/prof_gemm_r1 -f python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --profile-mode \
   > gpurun_out/ncu_gemm.log 2>&1; echo "gemm capture exit=$?"
timeout 900 $NCU --set full --clock-control none --import-source on \
   -k regex:"mag_tail_fwd|dropout_residual_ln_fwd|sdpa_fwd|ln_bwd_kernel|adamw|embed_ln_fwd" -s 60 -c 8 \
   -o gpurun_out/prof_rows_r1 -f python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --profile-mode \
   > gpurun_out/ncu_rows.log 2>&1; echo "rows capture exit=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r1.csv
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_r1.log 2>&1; echo "bench exit=$?"; tail -1 gpurun_out/bench_r1.log | cut -c1-1500
