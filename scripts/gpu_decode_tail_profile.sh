#!/bin/bash
# ncu --set full on the last four launches of linear_tc_pair_tail_kernel of ONE eager decode step: three 256-wide convolutions with the Block
# tail (PixelNorm + SiLU) in the epilogue and to_coor_logits with the arg-max in the epilogue.  CSV export of the raw page only.
tag=${1:-tail}
mkdir -p gpurun_out
timeout 500 ncu --profile-from-start off --set full --clock-control none -k regex:linear_tc_pair_tail_kernel --launch-skip 10 --launch-count 4 \
  -f -o gpurun_out/prof_dectail_$tag python scripts/gpu_decode_bench.py 64 profile > gpurun_out/ncu_full_dectail_$tag.log 2>&1
echo "full capture exit $?"; tail -2 gpurun_out/ncu_full_dectail_$tag.log | cut -c1-300
ncu -i gpurun_out/prof_dectail_$tag.ncu-rep --page raw --csv > gpurun_out/prof_dectail_${tag}_raw.csv 2> /dev/null
rm -f gpurun_out/prof_dectail_$tag.ncu-rep
python scripts/summarize_ncu.py full gpurun_out/prof_dectail_${tag}_raw.csv | head -120
