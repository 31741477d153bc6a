#!/bin/bash
# host SAH vs device LBVH: render time of C2 (1080p) and C5 at a 13x13 grid (4K)
for b in 1 2; do
  echo "== C2 builder $b"; TRT_BUILDER=$b timeout 300 python tools/render_once.py 2 32 4194304 fast 2 1 2>&1 | tail -1
  echo "== C5-13 builder $b"; TRT_BUILDER=$b TRT_GRID=13 timeout 600 python tools/render_once.py 5 4 4194304 fast 1 1 2>&1 | tail -1
done
