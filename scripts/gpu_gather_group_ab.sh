#!/bin/bash
# neighbour aggregation: one CTA per (mesh, group of slices) with double-buffered x slices (MESHTOK_GATHER_GROUP=1) against the default kernel
mkdir -p gpurun_out
rm -f /tmp/gg.pt
timeout 60 python scripts/gpu_gather_group_check.py /tmp/gg.pt 2>&1 | tail -1
MESHTOK_GATHER_GROUP=1 timeout 60 python scripts/gpu_gather_group_check.py /tmp/gg.pt 2>&1 | tail -2 | tee gpurun_out/gather_group_check.log
for gq in 0 1; do
  MESHTOK_GATHER_GROUP=$gq timeout 60 python bench.py --workload c5 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gg${gq}_c5.log 2>&1
  echo "GROUP=$gq c5: $(python scripts/summarize_bench.py gpurun_out/bench_gg${gq}_c5.log | grep -E 'value|sage_gather_mean' | tr '\n' ' ' | cut -c1-200)" | tee -a gpurun_out/gather_group_ab.log
done
