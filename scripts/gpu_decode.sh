#!/bin/bash
# decode path: GPU parity tests, throughput (fused and staged schedules), ncu launch list of one eager step
tag=${1:-dec}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decode.py -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/t_dec_$tag.log
timeout 300 python scripts/gpu_decode_bench.py 64 2>&1 | tee gpurun_out/dec_bench_$tag.log | tail -4
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_dec_$tag.csv \
  python scripts/gpu_decode_bench.py 64 profile > gpurun_out/ncu_dec_$tag.log 2>&1
echo "launch list exit $?"
python scripts/summarize_ncu.py launches gpurun_out/launches_dec_$tag.csv
