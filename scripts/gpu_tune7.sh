#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharding.py tests/test_gpu_peer.py -x -q 2>&1 | tail -2
for t in "" "2,0,4" "4,0,3"; do
  echo "TUNE=$t" | tee -a gpurun_out/tune7.log
  RECEMB_SEG_TUNE=$t timeout 300 python scripts/bench_configs.py cfg3 2>&1 | grep "pooled bwd" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['name'][:40], d['ms'], d['frac_of_measured_hbm'])" | tee -a gpurun_out/tune7.log
done
for t in "" "2,0,4"; do
RECEMB_SEG_TUNE=$t timeout 200 python scripts/bench_sharded.py --exchange peer --graph --steps 30 --warmup 3 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('cfg5 W=1 tune=$t', round(d['ms_per_step'],4))" | tee -a gpurun_out/tune7.log
done
