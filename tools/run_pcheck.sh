#!/bin/bash
mkdir -p gpurun_out
timeout 420 python tools/gemm_p_check.py > gpurun_out/gemm_p_check.log 2>&1; echo "exit=$?"
grep -E "PASS|FAIL|====|error|Error|timed out|us " gpurun_out/gemm_p_check.log | cut -c1-200 | head -120
