mkdir -p gpurun_out
timeout 900 python tools/all_configs.py > gpurun_out/r2_all_configs.json 2> gpurun_out/all_configs.err; tail -c 1500 gpurun_out/r2_all_configs.json
timeout 900 python tools/c5_bench.py 40 4 1 2>/dev/null | tail -1 > gpurun_out/r2_c5_device_build.json; cat gpurun_out/r2_c5_device_build.json
