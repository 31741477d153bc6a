for g in 0 4 8 12 16; do echo "== tri gate $g"; TRT_TRI_GATE=$g timeout 300 python tools/c5_quick.py 40 4 2>&1 | tail -1 | cut -c1-110; done
echo "== C2 gate 0 / 6"
for g in 0 6; do TRT_TRI_GATE=$g timeout 300 python tools/render_once.py 2 32 0 fast 2 0 2>&1 | tail -1 | cut -c1-60; done
TRT_TRI_GATE=8 TRT_COUNT=1 TRT_TRAV_STATS=1 timeout 300 python tools/c5_quick.py 40 4 1 2>&1 | grep trav | cut -c1-330
