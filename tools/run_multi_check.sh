#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29601 tools/ddp_check.py > gpurun_out/ddp_check_$N.log 2>&1; echo "ddp_check exit=$?"
grep -E "^\{|DDP_CHECK|spread" gpurun_out/ddp_check_$N.log | cut -c1-600 | tail -5
