#!/bin/bash
# N GPUs (meant for 4 / 8): piece timings of the peer step, equivalence check, then NCCL schedule vs peer-memory step on the same
# box for BASELINE configs[1] sharding (bs 48 / GPU) and configs[2] (MOSEI shapes, bs 128 / GPU)
mkdir -p gpurun_out
N=${1:-8}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29590 tools/peer_probe.py > gpurun_out/peer_probe_$N.log 2>&1; echo "peer_probe exit=$?"
grep -A10 "^world" gpurun_out/peer_probe_$N.log | cut -c1-200; grep -E "Error|error|Traceback|timed out" gpurun_out/peer_probe_$N.log | head -5 | cut -c1-300
MAGB200_PEER_ADAMW=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29601 tools/ddp_check.py > gpurun_out/ddp_check_peer_${N}.log 2>&1; echo "ddp_check(peer) exit=$?"
grep -E "DDP_CHECK|rel_first|spread|Error|error|Traceback|timed out" gpurun_out/ddp_check_peer_${N}.log | cut -c1-300 | tail -8
i=0
IFS='|' read -ra VS <<< "${VARIANTS:-MAGB200_PEER_ADAMW=0|MAGB200_PEER_ADAMW=1|MAGB200_PEER_ADAMW=1 WL=mosei128}"
for v in "${VS[@]}"; do
  i=$((i+1))
  wl=mosi48; case "$v" in *WL=mosei128*) wl=mosei128;; esac
  env $v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29610+i)) bench.py --gpus $N --workload $wl --steps ${STEPS:-60} --warmup 5 --no-cpu-baseline --no-class-api > gpurun_out/bench_n${N}_$i.log 2>&1; echo "bench [$v] exit=$?"
  grep '"metric"' gpurun_out/bench_n${N}_$i.log > gpurun_out/r02_bench_n${N}_$i.json
  python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/r02_bench_n${N}_$i.json')); print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value'], d['config'].get('workload'))
except Exception as e: print('parse fail', e)
"
  grep -E "Error|error|Traceback|warn|timed out" gpurun_out/bench_n${N}_$i.log | head -5 | cut -c1-300
done
