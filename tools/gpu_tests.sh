#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s "$@" > gpurun_out/pytest_gpu_full.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu_full.log
grep -E "mismatch|PSNR|passed|failed|FAILED|Error|error|replay" gpurun_out/pytest_gpu_full.log | grep -v Loader | head -80
