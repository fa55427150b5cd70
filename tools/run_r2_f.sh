#!/bin/bash
# interleaved epilogue (MAGB200_GEMM_CFG=500): correctness, then same-box A/B with the stream-K minimum-cluster knob; peer op test
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "peer_optimizer" 2>&1 | tail -5 | cut -c1-300
MAGB200_GEMM_CFG=500 timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q -x -k "gemm or linear or golden or gradient or oracle" 2>&1 | tail -6 | cut -c1-300
V="X=0|MAGB200_GEMM_CFG=500|MAGB200_STREAMK_MIN=18|MAGB200_STREAMK_MIN=27|MAGB200_STREAMK_MIN=37|MAGB200_GEMM_CFG=500 MAGB200_STREAMK_MIN=27|X=1"
VARIANTS="$V" STEPS=100 BENCH_FLAGS="--no-class-api" bash tools/run_bench_ab.sh 2>&1 | tee gpurun_out/r2f_ab.log
TRACE_CFGS=9,509 timeout 300 python tools/gemm_p_trace.py > gpurun_out/r02_gemm_persistent_trace_interleaved.txt 2>&1; echo "trace exit=$?"; grep -E "^==|cta0|epi\(|sub-block" gpurun_out/r02_gemm_persistent_trace_interleaved.txt | head -24 | cut -c1-250
