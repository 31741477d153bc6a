mkdir -p gpurun_out
echo "== dual ctx k=2"; timeout 300 python tools/dual_ctx.py 2 64 2 2>&1 | tail -1
echo "== dual ctx k=2 512thr"; TRT_FAST_THREADS=512 timeout 300 python tools/dual_ctx.py 2 64 2 2>&1 | tail -1
echo "== dual ctx k=4"; timeout 300 python tools/dual_ctx.py 2 64 4 2>&1 | tail -1
bash tools/gpu_prof.sh r2b 'k_extend_fast|k_shadow_fast|k_shade|k_regen' 40 8 8 X=0
