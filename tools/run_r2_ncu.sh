#!/bin/bash
# ncu --set full captures of the non-GEMM kernels and of forward / backward GEMM launches (small reports, several passes)
mkdir -p gpurun_out
NCU=/usr/local/cuda/bin/ncu
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --profile-mode"
cap() {  # name regex skip count
  timeout 500 $NCU --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o gpurun_out/r02_prof_$1 -f $CMD > gpurun_out/ncu_$1.log 2>&1
  echo "capture $1 exit=$? $(ls -la gpurun_out/r02_prof_$1.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
}
cap fwd_rows "embed_ln_fwd|mag_tail_fwd|sdpa_fwd|dropout_residual" 0 6
cap bwd_rows "sdpa_bwd|ln_bwd" 0 3
cap adamw "adamw" 0 1
cap gemm_fwd "gemm" 1 8
cap gemm_bwd "gemm" 52 10
