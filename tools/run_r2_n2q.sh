#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
MAGB200_PEER_ADAMW=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29601 tools/ddp_check.py > gpurun_out/ddp_check_peer_$N.log 2>&1; echo "ddp_check(peer) exit=$?"
grep -E "ddp_check rank|DDP_CHECK|rel_first|spread|Error|error|File \"/root|line " gpurun_out/ddp_check_peer_$N.log | cut -c1-300 | tail -40
