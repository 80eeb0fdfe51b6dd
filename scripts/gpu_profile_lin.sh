#!/bin/bash
# launch list of one step + --set full on every linear launch (raw page) + source page of one of them (skip N)
tag=$1; skip=${2:-5}
mkdir -p gpurun_out
CMD="python bench.py --workload c2 --warmup 3 --ncu-window 1"
$CMD > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_$tag.log; exit 1; }
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_list_$tag.log 2>&1
echo "launch list exit $?"
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:linear_tc -f -o gpurun_out/lin_${tag} $CMD > gpurun_out/ncu_lin_$tag.log 2>&1
echo "capture exit $?"
ncu -i gpurun_out/lin_${tag}.ncu-rep --page raw --csv > gpurun_out/lin_${tag}_raw.csv 2> /dev/null
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:linear_tc -s $skip -c 1 -f -o gpurun_out/lin1_${tag} $CMD > gpurun_out/ncu_lin1_$tag.log 2>&1
ncu -i gpurun_out/lin1_${tag}.ncu-rep --page source --csv > gpurun_out/lin1_${tag}_src.csv 2> /dev/null
rm -f gpurun_out/lin_${tag}.ncu-rep
ls -la gpurun_out | grep $tag
