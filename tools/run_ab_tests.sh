#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -x -q 2>&1 | tail -3 | cut -c1-300
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "tests failed"; exit 1; fi
bash tools/run_bench_ab.sh
