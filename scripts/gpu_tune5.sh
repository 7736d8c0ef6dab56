#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_peer.py tests/test_gpu_sharding.py -x -q 2>&1 | tail -3
run() {
  env "$@" timeout 200 python scripts/bench_sharded.py --exchange peer --graph --steps 30 --warmup 3 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$*', round(d['ms_per_step'],4))" | tee -a gpurun_out/tune5.log
}
run RECEMB_SEG_FIT=0 RECEMB_SEG_PF_BULK=0
run RECEMB_SEG_FIT=1 RECEMB_SEG_PF_BULK=0
run RECEMB_SEG_FIT=1 RECEMB_SEG_PF_BULK=1
run RECEMB_SEG_FIT=1 RECEMB_SEG_PF_BULK=1 RECEMB_SEG_TUNE=2,4,4
run RECEMB_SEG_FIT=1 RECEMB_SEG_PF_BULK=1 RECEMB_SEG_TUNE=1,4,5
run RECEMB_SEG_FIT=1 RECEMB_SEG_PF_BULK=0 RECEMB_SEG_TUNE=1,4,5
run RECEMB_SEG_FIT=1 RECEMB_SEG_PF_BULK=1 RECEMB_SEG_TUNE=2,8,4
