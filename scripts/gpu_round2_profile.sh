export RECEMB_PEER_BARRIER_TIMEOUT_S=20
# launch list of the default bench (cfg 2 headline + configs + sharded W=1)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_ncu_list.log 2>&1
# full captures: pool kernel (cfg3), fused push+apply (forced on one rank)
ncu --set full --clock-control none --import-source on -k regex:pool_kernel -s 4 -c 2 -o gpurun_out/r2_pool -f python scripts/bench_configs.py cfg3 > gpurun_out/r2_ncu_pool.log 2>&1
RECEMB_PEER_FUSED_PUSH=force ncu --set full --clock-control none --import-source on -k regex:seg_pre_kernel -s 3 -c 2 -o gpurun_out/r2_fused -f python scripts/bench_sharded.py --exchange peer --peer-forward push --steps 4 --warmup 2 > gpurun_out/r2_ncu_fused.log 2>&1
ls -la gpurun_out/*.ncu-rep
