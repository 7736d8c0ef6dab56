export RECEMB_PEER_BARRIER_TIMEOUT_S=10
python -m pytest tests/test_gpu_interaction.py tests/test_gpu_collection.py tests/test_gpu_peer.py tests/test_gpu_sharding.py -q -m gpu -x 2>&1 | tail -5
timeout 200 python scripts/bench_sharded.py --phase-bench 2>&1 | tail -1
python scripts/bench_configs.py cfg3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['name'][:70], d['ms'], d['frac_of_measured_hbm'], d.get('ms_lookup_cat_interaction'))"
