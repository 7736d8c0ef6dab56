#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
RECEMB_PLAN_IN_FORWARD=0 RECEMB_SEG_PRE=0 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_alt.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench.json')); r=d['roofline']; print('ms', round(d['ms_per_step'],4), 'seg', round(r['frac'],3), 'gather', round(r['gather_kernel']['frac'],3), 'step', round(r['whole_step']['frac'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d['e2e']['value'])"
