#!/bin/bash
mkdir -p gpurun_out
CMD="python scripts/bench_sharded.py --steps 3 --warmup 1 --exchange route"
$CMD > gpurun_out/shard_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:seg_kernel -s 10 -c 1 -o gpurun_out/prof_shard -f $CMD > gpurun_out/shard_ncu_full.log 2>&1
echo "exit $?"
