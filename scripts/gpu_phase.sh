#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR scripts/bench_sharded.py --phase-bench > gpurun_out/phase_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/phase_n$N.log
grep -v "^W1018\|OMP_NUM\|^\*\*\*" gpurun_out/phase_n$N.log | tail -4
