export RECEMB_PEER_BARRIER_TIMEOUT_S=20
W=${1:-2}
for f in 1 0 1 0; do
RECEMB_PEER_FUSED_PUSH=$f timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29511 scripts/bench_sharded.py --exchange peer --graph --steps 30 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('W',d['n_gpus'],'fused=$f',d['ms_per_step'],d['gpu_launches'],d['nvlink']['frac'])"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29511 scripts/bench_sharded.py --exchange peer --graph --steps 30 --warmup 3 --partition table 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('W',d['n_gpus'],'tablewise',d['ms_per_step'],d['gpu_launches'],d['nvlink']['frac'])"
