#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gemm_modes.py > gpurun_out/gemm_modes.log 2>&1; echo "exit=$?"; cat gpurun_out/gemm_modes.log | cut -c1-330
