#!/bin/bash
# ncu --set full on ONE launch of the steady-state bench step: kernel regex + how many matching launches to skip.
# Usage: bash scripts/gpu_profile_one.sh <tag> <kernel-regex> <skip> [workload]
tag=$1; kre=$2; skip=${3:-0}; wl=${4:-c2}
mkdir -p gpurun_out
CMD="python bench.py --workload $wl --warmup 3 --ncu-window 1"
$CMD > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_$tag.log; exit 1; }
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$kre -s $skip -c 1 -f -o gpurun_out/one_${tag} $CMD > gpurun_out/ncu_one_$tag.log 2>&1
echo "capture exit $?"; tail -2 gpurun_out/ncu_one_$tag.log | cut -c1-300
ncu -i gpurun_out/one_${tag}.ncu-rep --page raw --csv > gpurun_out/one_${tag}_raw.csv 2> /dev/null
ncu -i gpurun_out/one_${tag}.ncu-rep --page source --csv > gpurun_out/one_${tag}_src.csv 2> /dev/null
ls -la gpurun_out/ | grep one_$tag
