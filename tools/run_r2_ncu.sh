#!/bin/bash
# ncu --set full captures of the non-GEMM kernels (reports kept: small) and of forward / backward GEMM launches (raw-page
# CSV only: the reports are ~40 MB each and gpurun brings back at most 64 MiB)
mkdir -p gpurun_out
NCU=/usr/local/cuda/bin/ncu
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-class-api --profile-mode"
cap() {  # name regex skip count keep_report extra_flags
  timeout 500 $NCU --set full --clock-control none $6 -k regex:"$2" -s $3 -c $4 -o /tmp/r02_prof_$1 -f $CMD > gpurun_out/ncu_$1.log 2>&1
  echo "capture $1 exit=$? $(ls -la /tmp/r02_prof_$1.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
  $NCU -i /tmp/r02_prof_$1.ncu-rep --page raw --csv > gpurun_out/r02_prof_$1.raw.csv 2>/dev/null
  if [ "$5" = "keep" ]; then cp /tmp/r02_prof_$1.ncu-rep gpurun_out/; fi
}
cap fwd_rows "embed_ln_fwd|mag_tail_fwd|sdpa_fwd|dropout_residual" 0 6 keep "--import-source on"
cap bwd_rows "sdpa_bwd|ln_bwd" 0 3 keep "--import-source on"
cap adamw "adamw" 0 1 keep ""
cap gemm_fwd "gemm" 1 8 drop ""
cap gemm_bwd "gemm" 52 10 drop ""
du -sh gpurun_out
