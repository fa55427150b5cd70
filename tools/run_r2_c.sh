#!/bin/bash
# new kernels first (short timeouts: a protocol bug must not hold the box), then the suite, same-box A/B, benches
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k "embed_mag_fused" 2>&1 | tail -12 | cut -c1-300 > gpurun_out/r2c_k1.log; K1=$?; tail -3 gpurun_out/r2c_k1.log
timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k "layernorm_epilogue" 2>&1 | tail -12 | cut -c1-300 > gpurun_out/r2c_ln.log; tail -3 gpurun_out/r2c_ln.log
K1OK=$(grep -c " passed" gpurun_out/r2c_k1.log); K1BAD=$(grep -c "failed\|error\|Error" gpurun_out/r2c_k1.log)
LNOK=$(grep -c " passed" gpurun_out/r2c_ln.log); LNBAD=$(grep -c "failed\|error\|Error" gpurun_out/r2c_ln.log)
echo "K1 ok=$K1OK bad=$K1BAD  LN ok=$LNOK bad=$LNBAD"
timeout 2400 python -m pytest tests -m gpu -q --deselect tests/test_gpu_ops.py::test_embed_mag_fused_kernel_equals_the_three_kernel_path --deselect tests/test_gpu_ops.py::test_linear_with_layernorm_epilogue 2>&1 | tail -30 | cut -c1-400 > gpurun_out/r2c_tests.log; tail -4 gpurun_out/r2c_tests.log
V="MAGB200_TMA_REDUCE=0|MAGB200_TMA_REDUCE=1|MAGB200_TMA_REDUCE=1 MAGB200_STREAMK_FWD=1|MAGB200_TMA_REDUCE=1 MAGB200_FUSED_OPTIMIZER=1"
if [ "$K1OK" -ge 1 ] && [ "$K1BAD" -eq 0 ]; then V="$V|MAGB200_FUSED_MAG=1"; fi
if [ "$LNOK" -ge 1 ] && [ "$LNBAD" -eq 0 ]; then V="$V|MAGB200_FUSED_LN=1"; fi
VARIANTS="$V" STEPS=100 bash tools/run_bench_ab.sh 2>&1 | tee gpurun_out/r2c_ab.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/r2c_bench_default.log 2>&1; echo "bench default exit=$?"; grep '"metric"' gpurun_out/r2c_bench_default.log | cut -c1-400
timeout 600 python bench.py --workload s512 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_s512.log 2>&1; echo "bench s512 exit=$?"; grep '"metric"' gpurun_out/r2c_bench_s512.log | cut -c1-300; tail -2 gpurun_out/r2c_bench_s512.log | cut -c1-300
timeout 600 python bench.py --workload xlnet48 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench_xlnet.log 2>&1; echo "bench xlnet exit=$?"; grep '"metric"' gpurun_out/r2c_bench_xlnet.log | cut -c1-300; tail -2 gpurun_out/r2c_bench_xlnet.log | cut -c1-300
timeout 600 python bench.py --workload mosei128 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench_mosei.log 2>&1; echo "bench mosei exit=$?"; grep '"metric"' gpurun_out/r2c_bench_mosei.log | cut -c1-300
