#!/bin/bash
# ncu evidence: (1) every launch of two bench steps with its device time, (2) --set full on the top kernels.
# Usage: bash scripts/gpu_profile.sh <tag> [kernel-regex]
tag=${1:-r01}; kre=${2:-rvq_tc_search}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_$tag.log; exit 1; }
tail -1 gpurun_out/plain_$tag.log | cut -c1-600
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_list_$tag.log 2>&1
echo "launch list exit $?"; tail -3 gpurun_out/ncu_list_$tag.log
ncu --set full --clock-control none --import-source on -k regex:$kre -s 6 -c 2 -f -o gpurun_out/prof_${tag} $CMD > gpurun_out/ncu_full_$tag.log 2>&1
echo "full capture exit $?"; tail -3 gpurun_out/ncu_full_$tag.log; ls -la gpurun_out/
