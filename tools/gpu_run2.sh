#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python tools/render_once.py 2 2 1048576 fast > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1a.csv python tools/render_once.py 2 2 1048576 fast > gpurun_out/ncu1.log 2>&1
python tools/render_once.py 2 1 1048576 fast 0 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shadow|k_shade|k_regen' -s 40 -c 8 -o gpurun_out/prof_r1a python tools/render_once.py 2 1 1048576 fast 0 > gpurun_out/ncu2.log 2>&1
cat gpurun_out/plain.log | tail -2; tail -3 gpurun_out/ncu1.log; tail -3 gpurun_out/ncu2.log
