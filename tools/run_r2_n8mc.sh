#!/bin/bash
# 8 GPUs, NVSwitch multicast form of the peer-memory optimizer step: equivalence check, piece timings, bench (tight timeouts)
mkdir -p gpurun_out
N=${1:-8}
export MAGB200_PEER_MULTICAST=1
timeout 75 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29601 tools/ddp_check.py > gpurun_out/ddp_check_peer_${N}_mc.log 2>&1; echo "ddp_check(peer, multicast) exit=$?"
grep -E "multicast|DDP_CHECK|rel_first|spread|Error|error|Traceback|timed out" gpurun_out/ddp_check_peer_${N}_mc.log | cut -c1-300 | tail -8
timeout 75 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29590 tools/peer_probe.py > gpurun_out/peer_probe_${N}_mc.log 2>&1; echo "peer_probe exit=$?"
grep -A9 "^world" gpurun_out/peer_probe_${N}_mc.log | cut -c1-200; grep -E "Error|error|Traceback|timed out" gpurun_out/peer_probe_${N}_mc.log | head -5 | cut -c1-300
timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 60 --warmup 5 --no-cpu-baseline --no-class-api > gpurun_out/bench_n${N}_mc.log 2>&1; echo "bench exit=$?"
grep '"metric"' gpurun_out/bench_n${N}_mc.log > gpurun_out/r02_bench_n${N}_peer_multicast.json
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_n${N}_peer_multicast.json')); print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value'], d.get('last_loss'))"
grep -E "Error|error|Traceback|warn|timed out" gpurun_out/bench_n${N}_mc.log | head -5 | cut -c1-300
