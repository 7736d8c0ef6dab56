#!/bin/bash
mkdir -p gpurun_out
RECEMB_SEG_PRE=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_pre1.log
run() {
  env "$@" timeout 200 python scripts/bench_sharded.py --exchange peer --graph --steps 30 --warmup 3 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$*', round(d['ms_per_step'],4))" | tee -a gpurun_out/tune6.log
}
run RECEMB_SEG_PRE=0
run RECEMB_SEG_PRE=1
run RECEMB_SEG_PRE=1 RECEMB_SEG_TUNE=2,1,3
run RECEMB_SEG_PRE=1 RECEMB_SEG_TUNE=4,1,3
run RECEMB_SEG_PRE=1 RECEMB_SEG_TUNE=2,0,4
run RECEMB_SEG_PRE=1 RECEMB_SEG_TUNE=2,2,4
for pre in 0 1; do
  echo "PRE=$pre" | tee -a gpurun_out/tune6.log
  RECEMB_SEG_PRE=$pre timeout 300 python scripts/bench_configs.py cfg3 cfg4 2>&1 | grep "bwd" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['name'][:40], d['ms'], d['frac_of_measured_hbm'])" | tee -a gpurun_out/tune6.log
done
