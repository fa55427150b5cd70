#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_xlnet.py -m gpu -q --tb=short 2>&1 | tail -40 | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q --tb=short 2>&1 | tail -6
