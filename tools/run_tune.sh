#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short 2>&1 | tail -15
timeout 900 python tools/gemm_tune.py > gpurun_out/gemm_tune.log 2>&1; echo "exit=$?"; cat gpurun_out/gemm_tune.log
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short -x 2>&1 | tail -15
