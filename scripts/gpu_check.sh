#!/bin/bash
# Runs the GPU parity tests group by group (separate processes: a faulting kernel poisons only its own group),
# then a short bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n 6 gpurun_out/$name.log | cut -c1-400; }
run t_edges   python -m pytest tests/test_gpu_parity.py -q -m gpu -k "face_edges"
run t_linear  python -m pytest tests/test_gpu_parity.py -q -m gpu -k "linear"
run t_quant   python -m pytest tests/test_gpu_parity.py -q -m gpu -k "lfq or rvq"
run t_encoder python -m pytest tests/test_gpu_parity.py -q -m gpu -k "encoder or encode_and"
run t_e2e     python -m pytest tests/test_gpu_parity.py -q -m gpu -k "golden or variants or histogram or full_size or full_model"
run t_misc    python -m pytest tests/test_gpu_parity.py -q -m gpu -k "pipeline or cache or precomputed or config_matrix or dense_user or fallback"
run smoke     python __graft_entry__.py smoke
run bench     python bench.py --steps 10 --warmup 3
python scripts/summarize_bench.py gpurun_out/bench.log
