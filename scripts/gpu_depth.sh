#!/bin/bash
# e2e throughput of the host pipeline against the number of batches in flight
mkdir -p gpurun_out
for d in 1 2 3 4; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --pipeline-depth $d > gpurun_out/bench_d$d.log 2>&1
  echo "depth $d: $(python scripts/summarize_bench.py gpurun_out/bench_d$d.log | head -1)"
done
