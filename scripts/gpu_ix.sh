#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_interaction.py -x -q > gpurun_out/pytest_ix.log 2>&1; echo "ix pytest exit $?"; tail -30 gpurun_out/pytest_ix.log
