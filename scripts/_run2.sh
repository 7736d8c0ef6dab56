export RECEMB_PEER_BARRIER_TIMEOUT_S=20
W=${1:-2}
for f in 1 0; do
RECEMB_PEER_FUSED_PUSH=$f timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29511 scripts/bench_sharded.py --check --exchange peer --graph --steps 30 --warmup 3 > gpurun_out/r2_s${W}_f$f.log 2>&1
echo "rc=$?"
tail -1 gpurun_out/r2_s${W}_f$f.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('W',d['n_gpus'],'fused=$f groups',d['config']['pipeline_groups'],d['ms_per_step'],d['gpu_launches'],d['nvlink']['frac'])"
done
for c in 16 64; do
RECEMB_PEER_PUSH_CTAS=$c timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29511 scripts/bench_sharded.py --exchange peer --graph --steps 30 --warmup 3 > gpurun_out/r2_s${W}_c$c.log 2>&1
tail -1 gpurun_out/r2_s${W}_c$c.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('W',d['n_gpus'],'fused ctas=$c',d['ms_per_step'],d['gpu_launches'],d['nvlink']['frac'])"
done
RECEMB_PEER_GROUPS=2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29511 scripts/bench_sharded.py --exchange peer --graph --steps 30 --warmup 3 > gpurun_out/r2_s${W}_g2.log 2>&1
tail -1 gpurun_out/r2_s${W}_g2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('W',d['n_gpus'],'groups=2',d['ms_per_step'],d['gpu_launches'],d['nvlink']['frac'])"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29511 scripts/bench_sharded.py --phase-bench > gpurun_out/r2_phase${W}.log 2>&1
tail -1 gpurun_out/r2_phase${W}.log
