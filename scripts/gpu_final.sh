#!/bin/bash
# round-end check on one GPU: full GPU test suite, smoke, headline bench + reference arm, the other configs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/smoke.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; echo "ref rc=$?"
python scripts/bench_configs.py > gpurun_out/configs.jsonl 2>&1; echo "configs rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json')); r=d['roofline']
print('bench ms', round(d['ms_per_step'],4), 'value', d['value'], 'seg', round(r['frac'],3), 'gather', round(r['gather_kernel']['frac'],3), 'step', round(r['whole_step']['frac'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], d['clocks'])
print(open('gpurun_out/bench_ref.json').read()[:300])
for l in open('gpurun_out/configs.jsonl'):
    try:
        x=json.loads(l); print(x['name'][:60], x['ms'], x['frac_of_measured_hbm'], x.get('G_lookups_per_s'))
    except Exception: print(l[:200])
PY
