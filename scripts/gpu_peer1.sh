#!/bin/bash
# 1 GPU: peer-exchange kernels (W emulated), then cfg5 W=1 through each exchange
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_peer.py -x -q > gpurun_out/pytest_peer.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_peer.log
tail -15 gpurun_out/pytest_peer.log
for ex in route peer; do
  timeout 300 python scripts/bench_sharded.py --exchange $ex --steps 20 --warmup 3 > gpurun_out/shard1_$ex.log 2>&1; tail -2 gpurun_out/shard1_$ex.log
done
RECEMB_PHASES=1 timeout 300 python scripts/bench_sharded.py --exchange peer --graph --steps 20 --warmup 3 > gpurun_out/shard1_peer_graph.log 2>&1; tail -2 gpurun_out/shard1_peer_graph.log
