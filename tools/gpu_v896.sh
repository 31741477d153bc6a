mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_state.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu_state.log
for c in "1 16" "2 1" "2 16" "2 64" "3 64" "4 64"; do set -- $c; echo "== C$1 $2 spp"; timeout 300 python tools/render_once.py $1 $2 0 fast 2 0 2>&1 | tail -1 | cut -c1-70; done
timeout 300 python tools/c5_quick.py 40 4 2>&1 | tail -1 | cut -c1-110
bash tools/gpu_compressed_parity.sh
