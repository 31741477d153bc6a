mkdir -p gpurun_out
bash tools/gpu_check.sh "TRT_MERGED_TRACE=1"
TRT_ITER_LOG=gpurun_out/iterlog_r2c.txt timeout 300 python tools/render_once.py 2 64 0 fast 2 1 | tail -1 | cut -c1-200
timeout 300 python tools/render_once.py 2 8 0 fast 0 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_refill' -s 3 -c 2 -o gpurun_out/prof_refill python tools/render_once.py 2 8 0 fast 0 > gpurun_out/ncu_refill.log 2>&1; tail -2 gpurun_out/ncu_refill.log
