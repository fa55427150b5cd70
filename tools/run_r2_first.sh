#!/bin/bash
# round-2 opening call: full GPU suite on the merged tree (edge-case goldens, exact MAG), bench baseline for this round
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30 | cut -c1-400 | tee gpurun_out/r2_first_tests.log
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r2_first_bench.log 2>&1; echo "bench exit=$?"
grep '"metric"' gpurun_out/r2_first_bench.log | cut -c1-1500
