#!/bin/bash
# quick loop: the whole GPU parity file in one process, the bench, and the ncu launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
python scripts/summarize_bench.py gpurun_out/bench.log
bash scripts/gpu_launches.sh ${1:-quick} > /dev/null 2>&1
python scripts/summarize_ncu.py launches gpurun_out/launches_${1:-quick}.csv
