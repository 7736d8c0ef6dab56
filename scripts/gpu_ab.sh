#!/bin/bash
mkdir -p gpurun_out
for f in "" "--no-overlap-plan"; do
python bench.py --steps 10 --warmup 3 --no-cpu-baseline $f 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$f', 'ms', round(d['ms_per_step'],4), 'seg', round(r['frac'],3), 'gather', round(r['gather_kernel']['frac'],3), 'step', round(r['whole_step']['frac'],3), 'e2e', round(d['e2e']['ms_per_step'],3))"
done | tee gpurun_out/ab_overlap.log
