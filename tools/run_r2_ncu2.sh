#!/bin/bash
# ncu --set full of K1 (fused embedding + MAG kernel, opt-in path) and of the flash-style attention kernels at S = 512;
# raw-page CSV + summary, reports kept (small: 1-3 launches each)
mkdir -p gpurun_out
NCU=/usr/local/cuda/bin/ncu
cap() {  # name regex skip count env workload
  env $5 timeout 400 $NCU --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o /tmp/r02_$1 -f \
      python bench.py --workload $6 --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-class-api --profile-mode > gpurun_out/ncu_$1.log 2>&1
  echo "capture $1 exit=$? $(ls -la /tmp/r02_$1.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
  $NCU -i /tmp/r02_$1.ncu-rep --page raw --csv > gpurun_out/r02_$1.raw.csv 2>/dev/null
  cp /tmp/r02_$1.ncu-rep gpurun_out/ 2>/dev/null
}
cap k1_mag_fused "mag_fused" 1 1 MAGB200_FUSED_MAG=1 mosi48
cap flash_s512 "flash" 0 3 X=0 s512
cap gemm_ffn_in_src "gemm2p_tcgen05_kernel" 2 1 X=0 mosi48
python tools/ncu_summary.py gpurun_out/r02_k1_mag_fused.raw.csv gpurun_out/r02_flash_s512.raw.csv gpurun_out/r02_gemm_ffn_in_src.raw.csv > gpurun_out/r02_ncu_summary_k1_flash.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_ncu_summary_k1_flash.json'))
for f,rs in d.items():
    for r in rs: print(f, r['kernel'][:40], r['grid'], r.get('duration_us'), 'us dram', r.get('dram_bytes',0)/1e6, 'MB', r.get('dram_GBps'), 'GB/s tensor', r.get('tensor_pipe_pct'), 'issue', r.get('issue_active_pct'), 'regs', r.get('registers'))
PY
du -sh gpurun_out
