bash tools/gpu_ab.sh -t "TRT_MERGED_TRACE=0" "TRT_FINISH_BELOW=0" "TRT_FINISH_BELOW=0;TRT_SHADOW_PAIR=1" "X=1" "TRT_FINISH_BELOW=1048576"
TRT_FINISH_BELOW=1048576 TRT_ITER_LOG=gpurun_out/iterlog_r2d.txt timeout 300 python tools/render_once.py 2 64 0 fast 2 1 | tail -1 | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:'k_refill' -s 12 -c 2 -o gpurun_out/prof_refill2 python tools/render_once.py 2 32 0 fast 0 > gpurun_out/ncu_refill2.log 2>&1; tail -2 gpurun_out/ncu_refill2.log
