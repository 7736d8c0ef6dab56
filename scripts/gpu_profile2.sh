#!/bin/bash
# launch list + full captures (cfg2 seg / gather, cfg5 peer pool / seg_pre), each after a plain run exited 0;
# the reports are reduced to their raw-page CSV on the box (the pull-back limit is 64 MiB)
mkdir -p gpurun_out
R=/tmp/ncu_reports; mkdir -p $R
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:'seg_kernel' -c 1 -o $R/prof_seg -f $CMD > gpurun_out/ncu_full_seg.log 2>&1
echo "seg capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:'gather_kernel' -c 1 -o $R/prof_gather -f $CMD > gpurun_out/ncu_full_gather.log 2>&1
echo "gather capture exit $?"
CMD5="python scripts/bench_sharded.py --steps 3 --warmup 3 --exchange peer"
$CMD5 > gpurun_out/peer1_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:'seg_pre_kernel|pool_kernel' -s 6 -c 2 -o $R/prof_peer -f $CMD5 > gpurun_out/ncu_full_peer.log 2>&1
echo "peer capture exit $?"
for n in seg gather peer; do
  ncu -i $R/prof_$n.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$n.csv 2>/dev/null
  python scripts/ncu_summary.py raw $R/prof_$n.ncu-rep > gpurun_out/ncu_summary_$n.txt 2>&1
done
cp $R/prof_seg.ncu-rep gpurun_out/
du -sh gpurun_out
