#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gemm_p_trace.py > gpurun_out/gemm_p_trace.log 2>&1; echo "exit=$?"
grep -E "^==|cta0|epi\(|sub-block" gpurun_out/gemm_p_trace.log | cut -c1-220 | head -80
