#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 300 python scripts/bench_configs.py kshift 2>&1 | tee gpurun_out/kshift.log
