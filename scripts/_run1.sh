python -m pytest tests/test_gpu_peer.py tests/test_gpu_peer_multiproc.py tests/test_gpu_sharding.py -x -q -m gpu > gpurun_out/r2_t8.log 2>&1
tail -5 gpurun_out/r2_t8.log
for g in 1 2 4 8; do
RECEMB_PEER_GROUPS=$g python scripts/bench_sharded.py --exchange peer --peer-forward push --graph --steps 30 --warmup 3 > gpurun_out/r2_s1_g$g.log 2>&1
tail -1 gpurun_out/r2_s1_g$g.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('W1 groups',d['config']['pipeline_groups'],d['ms_per_step'],d['gpu_launches'])"
done
