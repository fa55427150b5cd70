#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short > gpurun_out/pytest_gpu_ops.log 2>&1; echo "pytest ops exit=$?"
tail -30 gpurun_out/pytest_gpu_ops.log | cut -c1-250
timeout 1200 python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short > gpurun_out/pytest_gpu_model.log 2>&1; echo "pytest model exit=$?"
tail -40 gpurun_out/pytest_gpu_model.log | cut -c1-250
timeout 300 python tools/e2e_probe.py > gpurun_out/e2e_probe.log 2>&1; echo "e2e exit=$?"; tail -6 gpurun_out/e2e_probe.log
