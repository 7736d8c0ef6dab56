#!/bin/bash
# N GPUs: peer-memory probe, sharded == unsharded check, cfg5 timing per exchange
N=${1:-2}
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 180 $TR scripts/peer_probe.py > gpurun_out/probe_n$N.log 2>&1; echo "probe rc=$?" >> gpurun_out/probe_n$N.log
grep -v "^W1018\|OMP_NUM\|^\*\*\*" gpurun_out/probe_n$N.log | tail -20
timeout 300 $TR scripts/bench_sharded.py --check --exchange peer --steps 20 --warmup 3 > gpurun_out/shard${N}_peer.log 2>&1; echo "rc=$?" >> gpurun_out/shard${N}_peer.log; tail -4 gpurun_out/shard${N}_peer.log
RECEMB_PHASES=1 timeout 300 $TR scripts/bench_sharded.py --exchange peer --graph --steps 20 --warmup 3 > gpurun_out/shard${N}_peer_graph.log 2>&1; echo "rc=$?" >> gpurun_out/shard${N}_peer_graph.log; tail -3 gpurun_out/shard${N}_peer_graph.log
timeout 300 $TR scripts/bench_sharded.py --exchange route --steps 20 --warmup 3 > gpurun_out/shard${N}_route.log 2>&1; tail -2 gpurun_out/shard${N}_route.log
