#!/bin/bash
# usage: gpu_ab.sh [-t] "ENV=1;ENV2=2" ...   C2 at 64 spp (auto pool) per environment: untimed ms/spp, then the per-kernel split; -t = run the GPU tests first
mkdir -p gpurun_out
if [ "$1" = "-t" ]; then shift; timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_state.log 2>&1; echo "pytest exit: $?"; grep -E "passed|failed|Error|assert" gpurun_out/pytest_gpu_state.log | tail -5; fi
for cfg in "$@"; do
  e=$(echo $cfg | tr ';' ' ')
  echo "== $cfg"
  env $e timeout 300 python tools/render_once.py 2 64 0 fast 2 0 2>&1 | tail -1 | cut -c1-60
  env $e timeout 300 python tools/render_once.py 2 64 0 fast 2 1 2>&1 | tail -1 | cut -c1-150
  env $e timeout 300 python tools/render_once.py 1 16 0 fast 2 0 2>&1 | tail -1 | cut -c1-50
done
