#!/bin/bash
# N GPUs: sharded == unsharded check + cfg5 timing through the peer exchange (CUDA graph)
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR scripts/bench_sharded.py --check --exchange peer --graph --steps 30 --warmup 3 > gpurun_out/shard${N}_peer_graph.log 2>&1; echo "rc=$?" >> gpurun_out/shard${N}_peer_graph.log
grep -v "^W1018\|OMP_NUM\|^\*\*\*" gpurun_out/shard${N}_peer_graph.log | tail -6
