#!/bin/bash
# decode path: arg-max in the epilogue of to_coor_logits (MESHTOK_DEC_ARGMAX=1, default) against the separate arg-max launch (=0)
tag=${1:-am}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decode.py -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/t_dec_$tag.log
for e in 1 0 1 0; do
  MESHTOK_DEC_ARGMAX=$e timeout 300 python scripts/gpu_decode_bench.py 64 fusedonly 2>&1 | tail -2 | sed "s/^/ARGMAX=$e /" | tee -a gpurun_out/dec_bench_$tag.log
done
