#!/bin/bash
# seg_kernel prefetch-distance sweep on cfg5 W=1 (all-unique rows, 51 GB table)
mkdir -p gpurun_out
for t in "2,1,4" "2,2,4" "2,4,4" "2,8,4" "4,1,4" "2,4,5" "1,4,5"; do
  RECEMB_SEG_TUNE=$t timeout 200 python scripts/bench_sharded.py --exchange peer --graph --steps 30 --warmup 3 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$t', round(d['ms_per_step'],4))" | tee -a gpurun_out/tune4.log
done
