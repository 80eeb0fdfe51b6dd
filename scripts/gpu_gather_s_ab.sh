#!/bin/bash
# neighbour aggregation: 16-column slices with 256 threads (3 CTAs per SM, MESHTOK_GATHER_S=16) against 32-column slices with 512 (2 per SM)
mkdir -p gpurun_out
MESHTOK_GATHER_S=16 timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "encoder or golden or full_model or config_matrix" 2>&1 | tail -3
for s in 32 16 32 16; do
  for w in c2 c5; do
    MESHTOK_GATHER_S=$s timeout 200 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gs${s}_$w.log 2>&1
    echo "S=$s $w: $(python scripts/summarize_bench.py gpurun_out/bench_gs${s}_$w.log | grep -E 'value|sage_gather_mean' | tr '\n' ' ' | cut -c1-200)"
  done
done
