#!/bin/bash
# Launch list + full ncu capture of the two hot kernels, each only after the plain run exited 0.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'seg_kernel' -c 1 \
    -o gpurun_out/prof_seg -f $CMD > gpurun_out/ncu_full_seg.log 2>&1
echo "seg capture exit $?"
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gather_kernel' -c 1 \
    -o gpurun_out/prof_gather -f $CMD > gpurun_out/ncu_full_gather.log 2>&1
echo "gather capture exit $?"
ls -la gpurun_out | head -30
