#!/bin/bash
# GPU tests, then same-box A/B of the stream-K switches
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | cut -c1-300 | tee gpurun_out/r2_tests.log
VARIANTS="MAGB200_STREAMK=0|MAGB200_STREAMK=1|MAGB200_STREAMK_FWD=0" STEPS=100 bash tools/run_bench_ab.sh 2>&1 | tee gpurun_out/r2_ab.log
