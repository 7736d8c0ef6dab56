export RECEMB_PEER_BARRIER_TIMEOUT_S=10
for d in 0 2; do
echo "debug=$d"; RECEMB_GATE_DEBUG=$d timeout 200 python scripts/bench_sharded.py --phase-bench 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read())['phase_ms']; print({k:v for k,v in d.items() if 'apply' in k})"
done
python -m pytest tests/test_gpu_peer.py -q -m gpu -x -k "fused_push" 2>&1 | tail -3
