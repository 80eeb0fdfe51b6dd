#!/bin/bash
# source-level ncu capture of the non-tensor-core kernels of one bench step
tag=${1:-r01f}; wl=${2:-c2}
mkdir -p gpurun_out
CMD="python bench.py --workload $wl --warmup 3 --ncu-window 1"
$CMD > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_$tag.log; exit 1; }
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'topology|face_embed|gather_mean|row_norm|rescore|scatter_mean|rvq_update' -f -o gpurun_out/prof_${tag} $CMD > gpurun_out/ncu_full_$tag.log 2>&1
echo "full capture exit $?"; tail -2 gpurun_out/ncu_full_$tag.log | cut -c1-300
ncu -i gpurun_out/prof_${tag}.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2> /dev/null
for k in topology face_embed gather_mean row_norm rescore; do
  ncu -i gpurun_out/prof_${tag}.ncu-rep --page source --csv -k regex:$k -c 1 > gpurun_out/prof_${tag}_src_$k.csv 2> /dev/null
done
sz=$(stat -c %s gpurun_out/prof_${tag}.ncu-rep); if [ "$sz" -gt 30000000 ]; then rm gpurun_out/prof_${tag}.ncu-rep; echo "dropped ${sz}-byte report (CSV exports kept)"; fi
ls -la gpurun_out/ | grep $tag
