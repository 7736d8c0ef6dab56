#!/bin/bash
# N GPUs: sharded == unsharded check + cfg5 timing through the peer exchange (CUDA graph), pull and push forward
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for pf in ${2:-pull push}; do
timeout 400 $TR scripts/bench_sharded.py --check --exchange peer --peer-forward $pf --graph --steps 30 --warmup 3 > gpurun_out/shard${N}_peer_${pf}.log 2>&1; echo "rc=$?" >> gpurun_out/shard${N}_peer_${pf}.log
grep -v "^W1018\|OMP_NUM\|^\*\*\*" gpurun_out/shard${N}_peer_${pf}.log | tail -4
done
