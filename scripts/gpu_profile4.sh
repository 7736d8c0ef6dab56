#!/bin/bash
mkdir -p gpurun_out
R=/tmp/ncu_reports; mkdir -p $R
python scripts/bench_configs.py kshift > gpurun_out/kshift_plain.log 2>&1 || exit 1
cat gpurun_out/kshift_plain.log | cut -c1-220
ncu --set full --clock-control none -k regex:'kshift_kernel|seg_kernel' -s 30 -c 3 -o $R/prof_ks2 -f python scripts/bench_configs.py kshift > gpurun_out/ncu_full_ks2.log 2>&1
echo "capture exit $?"
python scripts/ncu_summary.py raw $R/prof_ks2.ncu-rep > gpurun_out/ncu_summary_ks2.txt 2>&1
