#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharding.py -q -k "pool or shard or fused" 2>&1 | tail -2
CMD="python scripts/bench_sharded.py --steps 5 --warmup 2"
$CMD > gpurun_out/shard_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/shard_launches.csv $CMD > gpurun_out/shard_ncu.log 2>&1
echo "exit $?"; tail -1 gpurun_out/shard_plain.log | cut -c1-200
