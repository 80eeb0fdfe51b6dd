#!/bin/bash
# end-of-round confirmation: every GPU test, smoke(), the default bench line, the decode launch list of the current schedule
tag=${1:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/t_all_$tag.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke_$tag.log
timeout 600 python scripts/gpu_robustness.py 2>&1 | grep "^\[" | cut -c1-240 | tee gpurun_out/robustness_$tag.log
timeout 600 python bench.py > gpurun_out/bench_$tag.log 2>&1; echo "bench exit $?"
python scripts/summarize_bench.py gpurun_out/bench_$tag.log
timeout 300 python scripts/gpu_decode_bench.py 64 fusedonly 2>&1 | tail -1 | tee gpurun_out/dec_bench_$tag.log | cut -c1-400
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_dec_$tag.csv \
  python scripts/gpu_decode_bench.py 64 profile > gpurun_out/ncu_dec_$tag.log 2>&1
echo "launch list exit $?"
python scripts/summarize_ncu.py launches gpurun_out/launches_dec_$tag.csv | head -30
