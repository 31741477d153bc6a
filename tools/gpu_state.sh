#!/bin/bash
# current state: full GPU parity suite + timing split of C2 at 32 spp (pool 4 Mi); extra env sets as arguments
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_state.log 2>&1
echo "pytest exit: $?" | tee -a gpurun_out/pytest_gpu_state.log
grep -E "passed|failed|Error|assert" gpurun_out/pytest_gpu_state.log | tail -5
run() { echo "== $*"; env $* timeout 200 python tools/render_once.py 2 32 4194304 fast 2 1 2>&1 | tail -1; env $* timeout 200 python tools/render_once.py 2 32 4194304 fast 2 0 2>&1 | tail -1; }
if [ $# -eq 0 ]; then run X=0; fi
for cfg in "$@"; do run $cfg; done
