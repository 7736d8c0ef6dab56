#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharding.py tests/test_gpu_peer.py -x -q 2>&1 | tail -2
timeout 300 python scripts/bench_configs.py kshift 2>&1 | grep "bwd:" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['name'][:40], d['ms'], d['frac_of_measured_hbm'])"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('cfg2 ms', round(d['ms_per_step'],4), 'seg', round(r['frac'],3), 'seg ms', round(r['avg_launch_ms'],4), 'step', round(r['whole_step']['frac'],3))"
