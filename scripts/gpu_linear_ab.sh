#!/bin/bash
# A/B of the tcgen05 linear variants (0 = CTA pair, 16 epilogue warps; 2 = CTA pair, 8; 1 = single CTA): tests, then the bench.
mkdir -p gpurun_out
for v in ${VARIANTS:-0 2 1}; do
  export MESHTOK_LINEAR_TC=$v
  echo "=== MESHTOK_LINEAR_TC=$v"
  timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "linear or encoder or encode_and or golden or config_matrix or rvq" 2>&1 | tail -3
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_lin$v.log 2>&1
  python scripts/summarize_bench.py gpurun_out/bench_lin$v.log | grep -E "value|linear|norm|search"
done
