#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharding.py -x -q > gpurun_out/pytest_route.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_route.log
timeout 300 python scripts/bench_sharded.py --check --steps 10 2>&1 | tail -2 | cut -c1-200
