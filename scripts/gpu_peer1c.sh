#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_peer.py tests/test_gpu_sharding.py -x -q 2>&1 | tail -5
for pf in pull push; do
timeout 300 python scripts/bench_sharded.py --exchange peer --peer-forward $pf --graph --steps 30 --warmup 3 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$pf', round(d['ms_per_step'],4))"
done | tee gpurun_out/shard1_pushpull.log
