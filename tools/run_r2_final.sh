#!/bin/bash
# round-2 end evidence in one call: full GPU suite, smoke, default bench (CPU baseline + class-API arm), the other workloads,
# the reference arm, ncu launch list, --set full captures (raw CSV + one small report with source)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 | cut -c1-300 | tee gpurun_out/r02_pytest_gpu_final.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_final.log 2>&1; echo "bench exit=$?"
grep '"metric"' gpurun_out/r02_bench_final.log > gpurun_out/r02_bench_final.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_final.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d.get('e2e_class_api',{}).get('value'), d['eval_forward'], d['cpu_baseline']['value'])
r=d['roofline']; print({k:r[k] for k in ('kernel','achieved','peak','frac','traffic')}); print(r.get('gemm_family'), r.get('whole_step'))
for g in r.get('per_shape',[]): print(g['shape'], g['us'], g['frac'])
PY
for w in s512 xlnet48 mosei128; do
  timeout 600 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline --no-class-api > gpurun_out/r02_bench_$w.log 2>&1; echo "bench $w exit=$?"
  grep '"metric"' gpurun_out/r02_bench_$w.log > gpurun_out/r02_bench_$w.json
  python -c "
import json; d=json.load(open('gpurun_out/r02_bench_$w.json')); print('$w', {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'])"
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.log 2>&1; echo "reference arm exit=$?"; grep '"metric"' gpurun_out/r02_bench_reference.log | cut -c1-400
NCU=/usr/local/cuda/bin/ncu
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-class-api --profile-mode"
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none -s 984 -c 246 --csv \
   --log-file gpurun_out/r02_launches_v3.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-class-api --profile-mode \
   > gpurun_out/ncu_launches.log 2>&1; echo "launch-list exit=$?"
python tools/launch_summary.py gpurun_out/r02_launches_v3.csv | head -30 | tee gpurun_out/r02_launches_v3.summary.txt
cap() {  # name regex skip count keep_report extra_flags
  timeout 500 $NCU --set full --clock-control none $6 -k regex:"$2" -s $3 -c $4 -o /tmp/r02_final_$1 -f $CMD > gpurun_out/ncu_$1.log 2>&1
  echo "capture $1 exit=$? $(ls -la /tmp/r02_final_$1.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
  $NCU -i /tmp/r02_final_$1.ncu-rep --page raw --csv > gpurun_out/r02_final_$1.raw.csv 2>/dev/null
  if [ "$5" = "keep" ]; then cp /tmp/r02_final_$1.ncu-rep gpurun_out/; fi
}
cap gemm_fwd "gemm" 1 8 drop ""
cap gemm_bwd "gemm" 52 10 drop ""
cap gemm_ffn_in "gemm2p_tcgen05_kernel<0, 0, 256" 2 1 keep "--import-source on"
cap rows "sdpa_fwd|sdpa_bwd|ln_bwd|dropout_residual|adamw" 0 5 drop ""
python tools/ncu_summary.py gpurun_out/r02_final_*.raw.csv > gpurun_out/r02_ncu_summary_final.json
du -sh gpurun_out
