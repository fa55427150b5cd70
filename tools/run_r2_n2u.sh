#!/bin/bash
# N GPUs: weak vs relaxed.sys peer accesses in the peer-memory optimizer kernel: piece timings, equivalence, bench
mkdir -p gpurun_out
N=${1:-2}
for strong in 1 0; do
MAGB200_PEER_STRONG=$strong timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29590+strong)) tools/peer_probe.py > gpurun_out/peer_probe_${N}_strong$strong.log 2>&1; echo "peer_probe (strong $strong) exit=$?"
grep -A9 "^world" gpurun_out/peer_probe_${N}_strong$strong.log | cut -c1-200; grep -E "Error|error|Traceback|timed out" gpurun_out/peer_probe_${N}_strong$strong.log | head -5 | cut -c1-300
done
DDP_STEPS=3 MAGB200_PEER_ADAMW=1 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29607 tools/ddp_check.py > gpurun_out/ddp_check_peer_${N}_weak.log 2>&1; echo "ddp_check(peer, weak accesses, 3 steps) exit=$?"
grep -E "DDP_CHECK|rel_first|spread|Error|error|Traceback|timed out" gpurun_out/ddp_check_peer_${N}_weak.log | cut -c1-300 | tail -6
MAGB200_PEER_ADAMW=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 60 --warmup 5 --no-cpu-baseline --no-class-api > gpurun_out/bench_peer_n${N}_weak.log 2>&1; echo "bench exit=$?"
grep '"metric"' gpurun_out/bench_peer_n${N}_weak.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value'])"
