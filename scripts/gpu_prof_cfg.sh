#!/bin/bash
mkdir -p gpurun_out
python scripts/bench_configs.py kshift cfg3 cfg4 > gpurun_out/configs_plain.log 2>&1 || exit 1
cat gpurun_out/configs_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/configs_launches.csv python scripts/bench_configs.py kshift cfg3 cfg4 > gpurun_out/configs_ncu.log 2>&1
echo "ncu rc=$?"
