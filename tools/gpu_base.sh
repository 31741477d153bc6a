mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2a.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu_r2a.log
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_r2a.json 2> gpurun_out/bench_ref_r2a.err; tail -c 600 gpurun_out/bench_ref_r2a.json
timeout 900 python bench.py > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench exit $?"; tail -c 3000 gpurun_out/bench_r2a.json
TRT_ITER_LOG=gpurun_out/iterlog_r2a.txt timeout 300 python tools/render_once.py 2 64 0 fast 2 1 | tail -1
TRT_COUNT=1 TRT_TRAV_STATS=1 timeout 300 python tools/render_once.py 2 64 0 fast 2 0 2>&1 | tail -8
