#!/bin/bash
# Runs every GEMM probe case in its own process (bounded), summary in gpurun_out/gemm_probe.log
mkdir -p gpurun_out
LOG=gpurun_out/gemm_probe.log
: > $LOG
nvidia-smi --query-gpu=name,driver_version --format=csv >> $LOG 2>&1
for c in $(python tools/gemm_probe.py list); do
  echo "== $c" >> $LOG
  timeout 120 python tools/gemm_probe.py $c >> $LOG 2>&1
  echo "   exit=$?" >> $LOG
done
grep -E "PASS|FAIL|exit=[^0]|Error|error|timed out" $LOG | head -100
