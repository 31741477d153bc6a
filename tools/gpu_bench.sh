#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v -E "Loader|Vertices|\[BVH|Renderer\]" | tail -5
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | grep -v -E "Loader|Vertices|\[BVH|Renderer\]" | tail -2 | tee gpurun_out/bench_ref.json
python bench.py --steps 3 --warmup 3 2>&1 | grep -v -E "Loader|Vertices|\[BVH|Renderer\]" | tail -3 | tee gpurun_out/bench_ours.json
