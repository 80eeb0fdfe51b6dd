#!/bin/bash
# every bench workload once (c1 README, c3 quads/LFQ, c4 ragged, c5 = 1024 meshes per launch)
mkdir -p gpurun_out
for w in c1 c3 c4 c5; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.log 2>&1; echo "bench $w exit $?"; tail -n 2 gpurun_out/bench_$w.log | cut -c1-200 | grep -v "^{" ; python scripts/summarize_bench.py gpurun_out/bench_$w.log
done
timeout 600 python bench.py --workload c4 --precomputed-edges --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4pe.log 2>&1; echo "bench c4 precomputed exit $?"; python scripts/summarize_bench.py gpurun_out/bench_c4pe.log | head -3
