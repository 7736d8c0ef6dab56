#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
for tune in "" "2,1,4" "2,3,4" "2,2,5" "2,4,5" "2,2,6" "4,1,4" "1,4,5"; do
  echo "== RECEMB_SEG_TUNE=$tune"
  RECEMB_SEG_TUNE=$tune timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('ms/step',round(d['ms_per_step'],3),'G lookups/s',round(d['value']/1e9,3),'seg ms',round(r['avg_launch_ms'],4),'frac',round(r['frac'],3),'gather ms',round(r['gather_kernel']['avg_launch_ms'],4),'frac',round(r['gather_kernel']['frac'],3),'whole frac',round(r['whole_step']['frac'],3),'e2e ms',round(d['e2e']['ms_per_step'],3), 'launches', d['gpu_launches'])"
done
timeout 600 python scripts/bench_sharded.py --check --steps 10 2>&1 | tail -2 | cut -c1-400
