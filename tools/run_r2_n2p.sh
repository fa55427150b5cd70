#!/bin/bash
# N GPUs: peer-memory optimizer step (MAGB200_PEER_ADAMW=1): equivalence check, then same-box A/B against the NCCL schedule
mkdir -p gpurun_out
N=${1:-2}
MAGB200_PEER_ADAMW=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29601 tools/ddp_check.py > gpurun_out/ddp_check_peer_$N.log 2>&1; echo "ddp_check(peer) exit=$?"
grep -E "DDP_CHECK|rel_first|spread|Error|error|peer-memory|Traceback" gpurun_out/ddp_check_peer_$N.log | cut -c1-400 | tail -12
i=0
IFS='|' read -ra VS <<< "${VARIANTS:-MAGB200_PEER_ADAMW=0|MAGB200_PEER_ADAMW=1}"
for v in "${VS[@]}"; do
  i=$((i+1))
  env $v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29610+i)) bench.py --gpus $N --steps ${STEPS:-60} --warmup 5 --no-cpu-baseline > gpurun_out/bench_peer_n${N}_$i.log 2>&1; echo "bench [$v] exit=$?"
  grep '"metric"' gpurun_out/bench_peer_n${N}_$i.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value'], d['eval_forward']['value'])
except Exception as e: print('parse fail', e)
"
  grep -E "Error|error|Traceback|warn" gpurun_out/bench_peer_n${N}_$i.log | head -5 | cut -c1-300
done
