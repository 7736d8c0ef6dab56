#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharding.py tests/test_gpu_peer.py -x -q 2>&1 | tail -2
timeout 300 python scripts/bench_configs.py cfg3 cfg4 2>&1 | grep "bwd:" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['name'][:40], d['ms'], d['frac_of_measured_hbm'])"
timeout 200 python scripts/bench_sharded.py --exchange peer --graph --steps 30 --warmup 3 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('cfg5 W=1', round(d['ms_per_step'],4))"
