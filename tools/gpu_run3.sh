#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_tests.sh
timeout 600 python tools/gpu_probe.py 2 16 1048576,4194304 2>&1 | grep -v -E "Loader|Vertices|Renderer\]|BVH\]" > gpurun_out/probe_c2.log
cat gpurun_out/probe_c2.log
for b in 6 8 12 16; do echo "blocks/SM $b"; TRT_FAST_BLOCKS=$b timeout 300 python tools/render_once.py 2 8 2097152 fast 2>&1 | tail -1; done
