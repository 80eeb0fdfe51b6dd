#!/bin/bash
# ncu --set full on the two front-end kernels (one eager step between cudaProfilerStart / Stop); CSV export of the raw page only.
tag=${1:-fe}
mkdir -p gpurun_out
timeout 300 ncu --profile-from-start off --set full --clock-control none -f -o gpurun_out/prof_fe_$tag \
  python scripts/gpu_frontend_bench.py 64 profile > gpurun_out/ncu_full_fe_$tag.log 2>&1
echo "full capture exit $?"; tail -2 gpurun_out/ncu_full_fe_$tag.log | cut -c1-300
ncu -i gpurun_out/prof_fe_$tag.ncu-rep --page raw --csv > gpurun_out/prof_fe_${tag}_raw.csv 2> /dev/null
rm -f gpurun_out/prof_fe_$tag.ncu-rep
python scripts/summarize_ncu.py full gpurun_out/prof_fe_${tag}_raw.csv | head -60
