#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
timeout 300 python scripts/bench_configs.py kshift 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['name'][:34], d['ms'], d['frac_of_measured_hbm'])" | tee gpurun_out/tune9.log
