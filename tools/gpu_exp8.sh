bash tools/gpu_ab.sh -t "X=1"
TRT_ITER_LOG=gpurun_out/iterlog_r2g.txt timeout 300 python tools/render_once.py 2 64 0 fast 2 1 | tail -1 | cut -c1-200
echo "== C5 10.1M tris, compressed nodes (default) vs uncompressed"
timeout 900 python tools/c5_bench.py 40 4 1 2>&1 | tail -1
TRT_COMPRESSED=0 timeout 600 python tools/c5_bench.py 40 4 0 2>&1 | tail -1
