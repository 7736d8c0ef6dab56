#!/bin/bash
# full ncu captures of the cfg3 / k-shift kernels (interaction: tensor-pipe utilisation), reduced on the box
mkdir -p gpurun_out
R=/tmp/ncu_reports; mkdir -p $R
CMD="python scripts/bench_configs.py kshift cfg3"
$CMD > gpurun_out/configs_plain2.log 2>&1 || exit 1
ncu --set full --clock-control none -k regex:'dot_fwd_kernel|dot_bwd_kernel' -s 8 -c 2 -o $R/prof_ix -f $CMD > gpurun_out/ncu_full_ix.log 2>&1
echo "ix capture exit $?"
ncu --set full --clock-control none -k regex:'kshift_kernel|pool_kernel|seg_pre_kernel' -s 6 -c 1 -o $R/prof_ks -f python scripts/bench_configs.py kshift > gpurun_out/ncu_full_ks.log 2>&1
echo "kshift capture exit $?"
ncu --set full --clock-control none -k regex:'pool_kernel|seg_pre_kernel' -s 6 -c 12 -o $R/prof_c3 -f python scripts/bench_configs.py cfg3 > gpurun_out/ncu_full_c3.log 2>&1
echo "cfg3 capture exit $?"
for n in ix ks c3; do
  python scripts/ncu_summary.py raw $R/prof_$n.ncu-rep > gpurun_out/ncu_summary_$n.txt 2>&1
done
cat gpurun_out/configs_plain2.log | cut -c1-200
