#!/bin/bash
# first GPU contact: parity tests + probe
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x -s 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/gpu_probe.py 2 16 262144,1048576,4194304 > gpurun_out/probe_c2.log 2>&1
echo "probe exit: $?" >> gpurun_out/probe_c2.log
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/probe_c2.log | grep -v Loader | grep -v Vertices | tail -30
