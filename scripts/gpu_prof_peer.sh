#!/bin/bash
mkdir -p gpurun_out
CMD="python scripts/bench_sharded.py --steps 3 --warmup 3 --exchange peer"
$CMD > gpurun_out/peer1_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/peer1_launches.csv $CMD > gpurun_out/peer1_ncu.log 2>&1
echo "ncu rc=$?"
tail -2 gpurun_out/peer1_plain.log
