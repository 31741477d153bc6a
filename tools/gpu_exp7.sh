bash tools/gpu_ab.sh -t "X=1" "TRT_MERGED_TRACE=0"
TRT_ITER_LOG=gpurun_out/iterlog_r2f.txt timeout 300 python tools/render_once.py 2 64 0 fast 2 1 | tail -1 | cut -c1-200
echo "== compressed nodes forced on the small scenes"
TRT_COMPRESSED=1 timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_device_bvh.py tests/test_gpu_materials.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -4
TRT_COMPRESSED=1 timeout 300 python tools/render_once.py 2 64 0 fast 2 0 2>&1 | tail -1 | cut -c1-80
