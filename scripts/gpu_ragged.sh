#!/bin/bash
# ragged input layout: GPU parity tests against the padded layout + the C4 A/B
tag=${1:-rag}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ragged.py -q -m gpu -x 2>&1 | tail -25 | tee gpurun_out/t_ragged_$tag.log
timeout 300 python scripts/gpu_ragged_bench.py 256 2>&1 | tail -6 | tee gpurun_out/ragged_bench_$tag.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ragged_$tag.csv \
  python scripts/gpu_ragged_bench.py 256 profile > gpurun_out/ncu_ragged_$tag.log 2>&1
echo "launch list exit $?"
