#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_interaction.py -q 2>&1 | tail -3
timeout 300 python scripts/bench_configs.py cfg3 2>&1 | grep "interaction" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['name'][:40], d['ms'], d['frac_of_measured_hbm'])"
