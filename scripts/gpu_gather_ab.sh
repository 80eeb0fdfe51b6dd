#!/bin/bash
# A/B of the two neighbour-aggregation kernels (shared-memory slices = default, MESHTOK_GATHER_L2=1 = warp per row from L2)
mkdir -p gpurun_out
for v in 0 1; do
  for w in c2 c5 c3; do
    MESHTOK_GATHER_L2=$v timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g${v}_$w.log 2>&1
    echo "MESHTOK_GATHER_L2=$v $w: $(python scripts/summarize_bench.py gpurun_out/bench_g${v}_$w.log | grep -E 'value|gather_mean' | tr '\n' ' ')"
  done
done
