#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q 2>&1 | tail -4 | cut -c1-300
timeout 420 python tools/gemm_p_check.py > gpurun_out/gemm_p_check.log 2>&1; echo "exit=$?"
grep -E "FAIL|====|error|Error|timed out|us " gpurun_out/gemm_p_check.log | cut -c1-200 | head -60
timeout 300 python tools/gemm_p_trace.py > gpurun_out/gemm_p_trace.log 2>&1; echo "exit=$?"
grep -E "^==|cta0|epi\(" gpurun_out/gemm_p_trace.log | cut -c1-220 | head -80
