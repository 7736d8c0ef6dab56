#!/bin/bash
for pf in 0 1 2 4 8; do
  echo "== RECEMB_GATHER_PREFETCH=$pf"
  RECEMB_GATHER_PREFETCH=$pf timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('cfg2 ms/step',round(d['ms_per_step'],3),'gather ms',round(r['gather_kernel']['avg_launch_ms'],4),'frac',round(r['gather_kernel']['frac'],3))"
done
