#!/bin/bash
# same-box A/B: stream-K grid (all clusters / tile-count bound / 37 clusters), RNG, forward stream-K, fusions; GEMM pipeline trace
mkdir -p gpurun_out
P="MAGB200_LIB_PATH=$PWD/bert_multimodal_transformer_b200/libmagb200_philox.so"
V="X=0|MAGB200_STREAMK_GRID=1|MAGB200_STREAMK_GRID=37|MAGB200_STREAMK_GRID=1 $P|MAGB200_STREAMK_FWD=1|MAGB200_STREAMK_FWD=2|MAGB200_STREAMK_GRID=37 MAGB200_STREAMK_FWD=1|MAGB200_STREAMK_GRID=1 MAGB200_FUSED_LN=1|MAGB200_STREAMK_GRID=1 MAGB200_FUSED_MAG=1|MAGB200_STREAMK_GRID=1 MAGB200_FUSED_OPTIMIZER=1|X=1"
VARIANTS="$V" STEPS=100 BENCH_FLAGS="--no-class-api" bash tools/run_bench_ab.sh 2>&1 | tee gpurun_out/r2e_ab.log
timeout 300 python tools/gemm_p_trace.py > gpurun_out/r02_gemm_persistent_trace.txt 2>&1; echo "trace exit=$?"; head -30 gpurun_out/r02_gemm_persistent_trace.txt | cut -c1-250
