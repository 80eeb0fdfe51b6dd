#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl.log 2>&1; echo "bench (PDL) exit $?"; python scripts/summarize_bench.py gpurun_out/bench_pdl.log | grep -E "value"
MESHTOK_NO_PDL=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nopdl.log 2>&1; echo "bench (no PDL) exit $?"; python scripts/summarize_bench.py gpurun_out/bench_nopdl.log | grep -E "value"
