#!/bin/bash
# N GPUs: peer-memory optimizer step (default form: not overlapped, early table pass, gradients cleared under the next forward):
# piece timings, equivalence checks, same-box A/B
mkdir -p gpurun_out
N=${1:-2}
for mc in 0 1; do
MAGB200_PEER_MULTICAST=$mc timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29590+mc)) tools/peer_probe.py > gpurun_out/peer_probe_${N}_mc$mc.log 2>&1; echo "peer_probe (multicast $mc) exit=$?"
grep -A12 "^world" gpurun_out/peer_probe_${N}_mc$mc.log | cut -c1-200; grep -E "Error|error|Traceback|timed out" gpurun_out/peer_probe_${N}_mc$mc.log | head -5 | cut -c1-300
done
MAGB200_PEER_MULTICAST=1 MAGB200_PEER_ADAMW=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29607 tools/ddp_check.py > gpurun_out/ddp_check_peer_${N}_mc.log 2>&1; echo "ddp_check(peer, multicast) exit=$?"
grep -E "DDP_CHECK|rel_first|spread|Error|error|Traceback|timed out|multicast" gpurun_out/ddp_check_peer_${N}_mc.log | cut -c1-300 | tail -8
for steps in 3; do
  DDP_STEPS=$steps MAGB200_PEER_ADAMW=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+steps)) tools/ddp_check.py > gpurun_out/ddp_check_peer_${N}_s$steps.log 2>&1; echo "ddp_check(peer, $steps steps) exit=$?"
  grep -E "DDP_CHECK|rel_first|spread|Error|error|Traceback|timed out" gpurun_out/ddp_check_peer_${N}_s$steps.log | cut -c1-300 | tail -8
done
i=0
IFS='|' read -ra VS <<< "${VARIANTS:-MAGB200_PEER_ADAMW=0|MAGB200_PEER_ADAMW=1|MAGB200_PEER_ADAMW=1 MAGB200_PEER_MULTICAST=1|MAGB200_PEER_ADAMW=1 MAGB200_PEER_EARLY_BLOCKS=8}"
for v in "${VS[@]}"; do
  i=$((i+1))
  env $v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29610+i)) bench.py --gpus $N --steps ${STEPS:-60} --warmup 5 --no-cpu-baseline --no-class-api > gpurun_out/bench_peer_n${N}_$i.log 2>&1; echo "bench [$v] exit=$?"
  grep '"metric"' gpurun_out/bench_peer_n${N}_$i.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value'], d['eval_forward']['value'])
except Exception as e: print('parse fail', e)
"
  grep -E "Error|error|Traceback|warn|timed out" gpurun_out/bench_peer_n${N}_$i.log | head -5 | cut -c1-300
done
