#!/bin/bash
# full GPU parity suite, then timing + counters of C2/C3/C4
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_state.log 2>&1
echo "pytest exit: $?"; grep -E "passed|failed|Error|assert" gpurun_out/pytest_gpu_state.log | tail -5
timeout 200 python tools/render_once.py 2 32 4194304 fast 2 1 2>&1 | tail -1
timeout 200 python tools/render_once.py 2 64 4194304 fast 2 0 2>&1 | tail -1
for c in 2 3 4; do TRT_COUNT=1 timeout 300 python tools/render_once.py $c 4 4194304 fast 1 0 2>&1 | tail -1 | sed 's/.*closest:/closest:/'; done
