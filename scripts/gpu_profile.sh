#!/bin/bash
# ncu evidence for one steady-state step of the bench workload (bench.py --ncu-window brackets it with
# cudaProfilerStart/Stop):  (1) every launch with its device time, (2) --set full on the kernels matching a regex.
# The .ncu-rep is exported to CSV on the box (raw page + per-kernel source page) and dropped when it is too large
# to travel back (gpurun_out/ is capped at 64 MiB).
# Usage: bash scripts/gpu_profile.sh <tag> [workload] [kernel-regex for the full capture]
tag=${1:-r01}; wl=${2:-c2}; kre=${3:-.}
mkdir -p gpurun_out
CMD="python bench.py --workload $wl --warmup 3 --ncu-window 1"
$CMD > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_$tag.log; exit 1; }
tail -1 gpurun_out/plain_$tag.log | cut -c1-300
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_list_$tag.log 2>&1
echo "launch list exit $?"; tail -2 gpurun_out/ncu_list_$tag.log | cut -c1-300
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$kre -f -o gpurun_out/prof_${tag} $CMD > gpurun_out/ncu_full_$tag.log 2>&1
echo "full capture exit $?"; tail -2 gpurun_out/ncu_full_$tag.log | cut -c1-300
ncu -i gpurun_out/prof_${tag}.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2> /dev/null
for k in rvq_tc2_search linear_tc_pair linear_tc_kernel topology gather_mean face_sum face_bins rescore scatter_mean; do
  ncu -i gpurun_out/prof_${tag}.ncu-rep --page source --csv -k regex:$k -c 1 > gpurun_out/prof_${tag}_src_$k.csv 2> /dev/null
done
sz=$(stat -c %s gpurun_out/prof_${tag}.ncu-rep); if [ "$sz" -gt 30000000 ]; then rm gpurun_out/prof_${tag}.ncu-rep; echo "dropped ${sz}-byte report (CSV exports kept)"; fi
ls -la gpurun_out/ | grep $tag
