#!/bin/bash
# A/B: a layer's output linear and the next layer's lin in one launch (default) against separate launches (MESHTOK_LINEAR_B2B=0)
mkdir -p gpurun_out
for v in 1 0; do
  export MESHTOK_LINEAR_B2B=$v
  echo "=== MESHTOK_LINEAR_B2B=$v"
  timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -3
  for w in c2 c5; do
    timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b2b${v}_$w.log 2>&1
    echo "$w: $(python scripts/summarize_bench.py gpurun_out/bench_b2b${v}_$w.log | grep -E 'value|sage_linear' | tr '\n' ' ')"
  done
done
