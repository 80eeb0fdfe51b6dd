#!/bin/bash
# round 2: all GPU tests (one process per file so a faulting kernel poisons only its own file), smoke, short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n 8 gpurun_out/$name.log | cut -c1-600; }
run t_parity   python -m pytest tests/test_gpu_parity.py -q -m gpu -x
run t_ragged   python -m pytest tests/test_gpu_ragged.py -q -m gpu -x
run t_rest     python -m pytest tests -q -m gpu -x --ignore tests/test_gpu_parity.py --ignore tests/test_gpu_ragged.py
run smoke      python __graft_entry__.py smoke
run bench      python bench.py --steps 10 --warmup 3 ${BENCH_ARGS}
python scripts/summarize_bench.py gpurun_out/bench.log
