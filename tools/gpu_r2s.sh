#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2s}
for g in 1 0; do
  B200POSE_PEER_GATHER=$g timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((21000 + g)) tools/prof_peer_step.py 10000 1000000 2>&1 | grep -v "Warn\|warn\|^\*\|OMP" | cut -c1-170
done | tee $OUT/${TAG}_prof_peer_step.log
B200POSE_PEER_GATHER=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 21007 tools/prof_peer_step.py 100000 10000000 2>&1 | grep -v "Warn\|warn\|^\*\|OMP" | cut -c1-170 | tee -a $OUT/${TAG}_prof_peer_step.log
