export RECEMB_PEER_BARRIER_TIMEOUT_S=15
for i in 1 2 3; do
timeout 600 python -m pytest tests/test_gpu_peer.py tests/test_gpu_peer_multiproc.py -q -m gpu -x > gpurun_out/r2_t16_$i.log 2>&1
tail -2 gpurun_out/r2_t16_$i.log
done
