c2() { env $1 timeout 300 python tools/render_once.py 2 32 0 fast 2 1 2>&1 | tail -1 | cut -c1-130; }
c5() { env $1 TRT_GRID=40 timeout 300 python tools/render_once.py 5 4 0 fast 1 0 2>&1 | tail -1 | cut -c1-60; }
echo "== C2 st256 variants (0 none, 1 shade, 2 refill, 4 rest)"
for m in 0 1 2 4; do echo "TRT_EXP=$m"; c2 TRT_EXP=$m; done
echo "== C5 prefetch on / off, phases"
c5 TRT_EXP=0; c5 TRT_EXP=8
for ph in "4,12,8" "4,8,16" "6,8,16" "8,8,24" "3,16,12"; do echo "phases $ph"; c5 TRT_PHASES=$ph; done
