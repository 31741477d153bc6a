#!/bin/bash
# parity tests + one timing line per environment given as arguments ("A=1 B=2" strings)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_quick.log 2>&1
echo "pytest exit: $?" | tee -a gpurun_out/pytest_gpu_quick.log
grep -E "passed|failed|FAILED|Error|error|assert" gpurun_out/pytest_gpu_quick.log | grep -v Loader | head -20
run() { echo "== $*"; env $* timeout 120 python tools/render_once.py 2 16 4194304 fast 2 1 2>&1 | tail -1; }
if [ $# -eq 0 ]; then run X=0; fi
for cfg in "$@"; do run $cfg; done
