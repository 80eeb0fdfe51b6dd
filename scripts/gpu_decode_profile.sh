#!/bin/bash
# ncu --set full on every launch of ONE eager decode step (scripts/gpu_decode_bench.py <meshes> profile brackets it with
# cudaProfilerStart/Stop); only the CSV export of the raw page travels back.
# Usage: bash scripts/gpu_decode_profile.sh <tag> [meshes]
tag=${1:-dec}; B=${2:-64}
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --set full --clock-control none -f -o gpurun_out/prof_dec_$tag \
  python scripts/gpu_decode_bench.py $B profile > gpurun_out/ncu_full_dec_$tag.log 2>&1
echo "full capture exit $?"; tail -2 gpurun_out/ncu_full_dec_$tag.log | cut -c1-300
ncu -i gpurun_out/prof_dec_$tag.ncu-rep --page raw --csv > gpurun_out/prof_dec_${tag}_raw.csv 2> /dev/null
rm -f gpurun_out/prof_dec_$tag.ncu-rep
ls -la gpurun_out | grep dec_$tag
