#!/bin/bash
# The parity files with the compressed 64-byte nodes forced on the small scenes (by default only trees above 16 MB
# are compressed): ids, d_min, shadow bits, radiance gates and goldens must pass unchanged.
TRT_COMPRESSED=1 timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_device_bvh.py tests/test_gpu_materials.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -4
TRT_COMPRESSED=1 timeout 300 python tools/render_once.py 2 64 0 fast 2 0 2>&1 | tail -1 | cut -c1-80
