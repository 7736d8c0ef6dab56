export RECEMB_PEER_BARRIER_TIMEOUT_S=20
W=${1:-2}
for g in 1 2 4; do
RECEMB_PEER_GROUPS=$g timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29511 scripts/bench_sharded.py --check --exchange peer --graph --steps 30 --warmup 3 > gpurun_out/r2_s${W}_g$g.log 2>&1
echo "rc=$?"
tail -1 gpurun_out/r2_s${W}_g$g.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('W',d['n_gpus'],'groups',d['config']['pipeline_groups'],d['ms_per_step'],d['gpu_launches'],d['nvlink']['frac'])"
done
