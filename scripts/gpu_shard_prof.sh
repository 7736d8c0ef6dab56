#!/bin/bash
mkdir -p gpurun_out
CMD="python scripts/bench_sharded.py --steps 5 --warmup 2 --exchange route"
$CMD > gpurun_out/shard_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/shard_launches.csv $CMD > gpurun_out/shard_ncu.log 2>&1
echo "exit $?"; tail -1 gpurun_out/shard_plain.log | cut -c1-200
