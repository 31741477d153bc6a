mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_state.log 2>&1; echo "pytest exit: $?"; grep -E "passed|failed|Error|assert" gpurun_out/pytest_gpu_state.log | tail -5
python tools/c5_profile.py > gpurun_out/c5_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_trace_fast|k_shade' -s 6 -c 3 -o gpurun_out/prof_c5_r2 python tools/c5_profile.py > gpurun_out/ncu_c5.log 2>&1; tail -2 gpurun_out/ncu_c5.log
TRT_COUNT=1 TRT_TRAV_STATS=1 TRT_GRID=40 timeout 300 python tools/render_once.py 5 4 0 fast 1 0 2>&1 | tail -4 | cut -c1-400
