#!/bin/bash
# parity tests, then a tuning sweep of the persistent traversal kernels on C2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_v2.log 2>&1
echo "pytest exit: $?" | tee -a gpurun_out/pytest_gpu_v2.log
grep -E "passed|failed|FAILED|Error|error|replays [1-9]" gpurun_out/pytest_gpu_v2.log | grep -v Loader | head -20
run() { echo "== $*"; env "$@" timeout 120 python tools/render_once.py 2 16 4194304 fast 2 1 2>&1 | tail -1; }
run TRT_FAST_THREADS=512
run TRT_FAST_THREADS=512 TRT_SMEM_NODES=0
run TRT_FAST_THREADS=768
run TRT_FAST_THREADS=768 TRT_SMEM_NODES=0
run TRT_FAST_THREADS=1024
run TRT_FAST_THREADS=1024 TRT_SMEM_NODES=0
run TRT_FAST_THREADS=768 TRT_REFILL=16
run TRT_FAST_THREADS=768 TRT_REFILL=30
