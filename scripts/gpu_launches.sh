#!/bin/bash
# ncu launch list (device time per launch, serialised, cold caches) of one steady-state step
tag=$1; wl=${2:-c2}
mkdir -p gpurun_out
CMD="python bench.py --workload $wl --warmup 3 --ncu-window 1"
$CMD > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_$tag.log; exit 1; }
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_list_$tag.log 2>&1
echo "launch list exit $?"
