#!/bin/bash
# stream-K on every cluster, hash RNG, 16-warp ln_bwd, LN epilogue with TMA residual: suite + same-box A/B
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -30 | cut -c1-400 > gpurun_out/r2d_tests.log; tail -4 gpurun_out/r2d_tests.log
V="X=0|MAGB200_LIB_PATH=$PWD/bert_multimodal_transformer_b200/libmagb200_philox.so|MAGB200_STREAMK_FWD=1|MAGB200_STREAMK_FWD=2|MAGB200_FUSED_LN=1|MAGB200_FUSED_MAG=1|MAGB200_FUSED_OPTIMIZER=1"
VARIANTS="$V" STEPS=100 bash tools/run_bench_ab.sh 2>&1 | tee gpurun_out/r2d_ab.log
NCU=/usr/local/cuda/bin/ncu
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none -s 984 -c 246 --csv \
   --log-file gpurun_out/r02_launches_v2.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-class-api --profile-mode \
   > gpurun_out/ncu_launches.log 2>&1; echo "launch-list exit=$?"
python tools/launch_summary.py gpurun_out/r02_launches_v2.csv | head -30 | tee gpurun_out/r02_launches_v2.summary.txt
