#!/bin/bash
# host SAH vs device LBVH (with / without the SAH top levels): render time of C2 (1080p) and C5 at a 13x13 grid (4K)
timeout 900 python -m pytest tests/test_gpu_device_bvh.py -m gpu -x -q 2>&1 | tail -2
for cfg in "TRT_BUILDER=1" "TRT_BUILDER=2 TRT_TOP_SAH=0" "TRT_BUILDER=2 TRT_TOP_SAH=1"; do
  echo "== C2 $cfg"; env $cfg TRT_COUNT=0 timeout 300 python tools/render_once.py 2 32 4194304 fast 2 1 2>&1 | tail -1
  echo "== C5-13 $cfg"; env $cfg TRT_GRID=13 timeout 600 python tools/render_once.py 5 4 4194304 fast 1 1 2>&1 | tail -1
  echo "== C5-13 counts $cfg"; env $cfg TRT_GRID=13 TRT_COUNT=1 timeout 600 python tools/render_once.py 5 2 4194304 fast 1 0 2>&1 | tail -1 | sed 's/.*closest:/closest:/'
done
