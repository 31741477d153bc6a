#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_tests.sh 2>&1 | tail -4
for b in 4 5 6 8; do echo "blocks/SM $b"; TRT_FAST_BLOCKS=$b timeout 300 python tools/render_once.py 2 8 2097152 fast 2>&1 | tail -1; done
bash tools/gpu_prof.sh r1c "k_extend_fast|k_shadow_fast" 20 2
