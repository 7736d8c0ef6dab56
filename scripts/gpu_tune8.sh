#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "kshift or gather or flat" 2>&1 | tail -2
for v in 0 1; do
RECEMB_KSHIFT_L1=$v timeout 300 python scripts/bench_configs.py kshift 2>&1 | grep "fwd" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('KSHIFT_L1=$v', d['name'][:30], d['ms'], d['frac_of_measured_hbm'])" | tee -a gpurun_out/tune8.log
RECEMB_GATHER_L1=$v timeout 300 python scripts/bench_configs.py cfg4 2>&1 | grep "cfg4 fwd:" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('GATHER_L1=$v', d['name'][:30], d['ms'], d['frac_of_measured_hbm'])" | tee -a gpurun_out/tune8.log
done
RECEMB_GATHER_L1=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('cfg2 GATHER_L1=1 ms', round(d['ms_per_step'],4), 'gather', round(r['gather_kernel']['frac'],3))" | tee -a gpurun_out/tune8.log
