#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_peer.py tests/test_gpu_sharding.py -x -q 2>&1 | tail -3
timeout 300 python scripts/bench_sharded.py --exchange peer --graph --steps 30 --warmup 3 2>&1 | tail -1 > gpurun_out/shard1_peer_graph.log; cat gpurun_out/shard1_peer_graph.log
