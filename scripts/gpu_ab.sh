#!/bin/bash
# A/B of environment switches on ONE box: scripts/gpu_ab.sh "<bench args>" "VAR=a" "VAR=b" ...   (each variant = one bench.py process)
mkdir -p gpurun_out
ARGS=$1; shift
i=0
for v in "$@"; do
  i=$((i+1))
  echo "=== variant $i: $v"
  env $v timeout 600 python bench.py --workloads none --no-cpu-baseline $ARGS > gpurun_out/ab_$i.log 2>&1 || tail -5 gpurun_out/ab_$i.log
  python scripts/summarize_bench.py gpurun_out/ab_$i.log | head -${ABLINES:-14}
done
