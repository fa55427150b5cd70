#!/bin/bash
mkdir -p gpurun_out
bash tools/run_gemm_probe.sh > gpurun_out/gemm_summary.txt 2>&1
grep -c PASS gpurun_out/gemm_summary.txt; grep -E "FAIL|exit=[^0]" gpurun_out/gemm_summary.txt
timeout 600 python tools/model_probe.py > gpurun_out/model_probe.log 2>&1; echo "model_probe exit=$?"
tail -60 gpurun_out/model_probe.log
