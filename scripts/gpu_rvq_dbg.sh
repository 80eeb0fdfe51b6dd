#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "rvq or golden or full_size or variants" -x > gpurun_out/t_rvq2.log 2>&1; echo "rvq tests (pair) exit $?"; tail -n 12 gpurun_out/t_rvq2.log | cut -c1-300
MESHTOK_RVQ_TC=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "rvq" -x > gpurun_out/t_rvq1.log 2>&1; echo "rvq tests (single) exit $?"; tail -n 3 gpurun_out/t_rvq1.log | cut -c1-300
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pair.log 2>&1; echo "bench (pair) exit $?"; python scripts/summarize_bench.py gpurun_out/bench_pair.log | grep -E "value|rvq|rechecked"
