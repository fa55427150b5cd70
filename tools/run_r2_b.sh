#!/bin/bash
# full GPU suite (no -x), stream-K forward A/B, ncu launch list of one eager step
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -40 | cut -c1-400 > gpurun_out/r2b_tests.log; tail -5 gpurun_out/r2b_tests.log
VARIANTS="MAGB200_STREAMK_FWD=0|MAGB200_STREAMK_FWD=1" STEPS=100 bash tools/run_bench_ab.sh 2>&1 | tee gpurun_out/r2b_ab.log
NCU=/usr/local/cuda/bin/ncu
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none -s 984 -c 246 --csv \
   --log-file gpurun_out/r02_launches_v1.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --profile-mode \
   > gpurun_out/ncu_launches.log 2>&1; echo "launch-list exit=$?"
python tools/launch_summary.py gpurun_out/r02_launches_v1.csv | head -30 | tee gpurun_out/r02_launches_v1.summary.txt
