#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_tests.sh 2>&1 | tail -3
for v in 4 6 8; do for p in 1048576 4194304; do echo "variant $v pool $p"; TRT_FAST_VARIANT=$v timeout 300 python tools/render_once.py 2 8 $p fast 2>&1 | tail -1; done; done
TRT_FAST_VARIANT=8 bash tools/gpu_prof.sh r1d "k_extend_fast|k_shadow_fast" 20 2
