#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_n$N.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1]); r=d['roofline']
print('value', d['value'], 'ms', round(d['ms_per_step'],4), 'step', round(r['whole_step']['frac'],3), 'e2e', round(d['e2e']['ms_per_step'],3))
print('sharded', json.dumps(d['sharded_cfg5'])[:900])"
