#!/bin/bash
# TMA-reduce epilogue: tests, same-box A/B, the full default bench line, the other BASELINE workloads
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -30 | cut -c1-400 > gpurun_out/r2c_tests.log; tail -4 gpurun_out/r2c_tests.log
VARIANTS="MAGB200_TMA_REDUCE=0|MAGB200_TMA_REDUCE=1|MAGB200_TMA_REDUCE=1 MAGB200_STREAMK_FWD=1|MAGB200_TMA_REDUCE=1 MAGB200_FUSED_OPTIMIZER=1" STEPS=100 bash tools/run_bench_ab.sh 2>&1 | tee gpurun_out/r2c_ab.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/r2c_bench_default.log 2>&1; echo "bench default exit=$?"; grep '"metric"' gpurun_out/r2c_bench_default.log | cut -c1-400
timeout 600 python bench.py --workload s512 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_s512.log 2>&1; echo "bench s512 exit=$?"; grep '"metric"' gpurun_out/r2c_bench_s512.log | cut -c1-300; tail -2 gpurun_out/r2c_bench_s512.log | cut -c1-300
timeout 600 python bench.py --workload xlnet48 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench_xlnet.log 2>&1; echo "bench xlnet exit=$?"; grep '"metric"' gpurun_out/r2c_bench_xlnet.log | cut -c1-300; tail -2 gpurun_out/r2c_bench_xlnet.log | cut -c1-300
timeout 600 python bench.py --workload mosei128 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench_mosei.log 2>&1; echo "bench mosei exit=$?"; grep '"metric"' gpurun_out/r2c_bench_mosei.log | cut -c1-300
