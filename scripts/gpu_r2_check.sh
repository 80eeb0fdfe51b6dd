#!/bin/bash
# round 2: all GPU tests (one process per file so a faulting kernel poisons only its own file), smoke, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-8} gpurun_out/$name.log | cut -c1-700; }
if [ -z "$SKIP_TESTS" ]; then
run t_parity   python -m pytest tests/test_gpu_parity.py -q -m gpu -x
run t_fullsize python -m pytest tests/test_gpu_fullsize.py -q -m gpu -x -s
run t_ragged   python -m pytest tests/test_gpu_ragged.py -q -m gpu -x
run t_rest     python -m pytest tests -q -m gpu -x --ignore tests/test_gpu_parity.py --ignore tests/test_gpu_ragged.py --ignore tests/test_gpu_fullsize.py
run smoke      python __graft_entry__.py smoke
fi
TAILN=2 run bench      python bench.py ${BENCH_ARGS}
python scripts/summarize_bench.py gpurun_out/bench.log
if [ -n "$WITH_REF" ]; then TAILN=2 run bench_ref python bench.py --impl reference --steps 3 --warmup 1; fi
