#!/bin/bash
# full GPU parity suite, then C2 timings at 16 / 64 spp and C1 for each environment given ("A=1 B=2" strings)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_state.log 2>&1
echo "pytest exit: $?"; grep -E "passed|failed|Error|assert" gpurun_out/pytest_gpu_state.log | tail -5
if [ $# -eq 0 ]; then set -- "X=0"; fi
for cfg in "$@"; do
  for sp in 16 64; do echo "== $cfg spp $sp"; env $cfg timeout 300 python tools/render_once.py 2 $sp 0 fast 2 1 2>&1 | tail -1 | cut -c1-150; done
  echo "== $cfg spp 64 untimed"; env $cfg timeout 300 python tools/render_once.py 2 64 0 fast 2 0 2>&1 | tail -1 | cut -c1-60
  echo "== $cfg C1"; env $cfg timeout 300 python tools/render_once.py 1 16 0 fast 2 0 2>&1 | tail -1 | cut -c1-100
done
