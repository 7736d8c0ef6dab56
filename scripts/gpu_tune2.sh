#!/bin/bash
for tune in "" "2,2,4" "2,3,4" "2,4,4" "2,8,4" "2,4,5" "1,4,5"; do
  echo "== RECEMB_SEG_TUNE=$tune"
  RECEMB_SEG_TUNE=$tune timeout 300 python scripts/bench_sharded.py --steps 10 --exchange route 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cfg5 route W=1 ms/step', round(d['ms_per_step'],4))"
  RECEMB_SEG_TUNE=$tune timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('cfg2 ms/step',round(d['ms_per_step'],3),'seg ms',round(r['avg_launch_ms'],4),'frac',round(r['frac'],3))"
done
