export RECEMB_PEER_BARRIER_TIMEOUT_S=20
python -m pytest tests/ -x -q -m gpu > gpurun_out/r2_full_gpu.log 2>&1
tail -6 gpurun_out/r2_full_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke2.log 2>&1; tail -3 gpurun_out/r2_smoke2.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; tail -2 gpurun_out/r2_bench5.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench5_ref.json 2>> gpurun_out/r2_bench5.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench5.json'))
print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline'])
print(d['sharded_cfg5']['ms_per_step'] if d.get('sharded_cfg5') else None)
for c in d.get('configs') or []: print(c['name'][:80], c['ms'], c['frac_of_measured_hbm'])
"
cat gpurun_out/r2_bench5_ref.json | head -c 600
