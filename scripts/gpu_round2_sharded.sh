export RECEMB_PEER_BARRIER_TIMEOUT_S=20
W=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29511 scripts/bench_sharded.py --check --exchange peer --graph --steps 30 --warmup 3 > gpurun_out/r2_final_s$W.log 2>&1
tail -1 gpurun_out/r2_final_s$W.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('W',d['n_gpus'],'rowwise fused',d['ms_per_step'],d['gpu_launches'],d['nvlink']['frac'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29511 scripts/bench_sharded.py --exchange peer --graph --steps 30 --warmup 3 --partition table > gpurun_out/r2_final_tw$W.log 2>&1
tail -1 gpurun_out/r2_final_tw$W.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('W',d['n_gpus'],'tablewise',d['ms_per_step'],d['gpu_launches'],d['nvlink']['frac'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $W --steps 20 --warmup 3 > gpurun_out/r2_bench_n$W.json 2> gpurun_out/r2_bench_n$W.err
echo "rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_n$W.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['value'], d.get('host'))
s=d['sharded_cfg5']; print({k: s.get(k) for k in ('ms_per_step','value','w1_anchor_ms_per_step','efficiency_vs_w1','error')}); print(s.get('tablewise'))
"
