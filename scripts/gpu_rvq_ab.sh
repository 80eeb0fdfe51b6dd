#!/bin/bash
# RVQ screening kernels: parity tests, the bench with the default single-CTA kernel and with the CTA-pair kernel
# (MESHTOK_RVQ_TC=2), and two timing experiments on the single-CTA kernel (MESHTOK_RVQ_DEBUG; results are garbage,
# only the rvq_search_tc time means anything): 1 = epilogue never reads TMEM, 2 = reads TMEM but does no arithmetic.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "rvq" -x > gpurun_out/t_rvq.log 2>&1; echo "rvq tests exit $?"; tail -n 3 gpurun_out/t_rvq.log | cut -c1-300
MESHTOK_RVQ_TC=2 timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "rvq" -x > gpurun_out/t_rvq2.log 2>&1; echo "rvq tests (pair) exit $?"; tail -n 3 gpurun_out/t_rvq2.log | cut -c1-300
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_single.log 2>&1; echo "bench (single) exit $?"; python scripts/summarize_bench.py gpurun_out/bench_single.log
MESHTOK_RVQ_TC=2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pair.log 2>&1; echo "bench (pair) exit $?"; python scripts/summarize_bench.py gpurun_out/bench_pair.log | grep -E "value|rvq"
for d in 1 2; do
MESHTOK_RVQ_DEBUG=$d timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dbg$d.log 2>&1; echo "bench (debug $d) exit $?"; python scripts/summarize_bench.py gpurun_out/bench_dbg$d.log | grep -E "rvq_search"
done
