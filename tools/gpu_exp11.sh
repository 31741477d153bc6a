c2() { env $1 timeout 300 python tools/render_once.py 2 64 ${2:-0} fast 2 0 2>&1 | tail -1 | cut -c1-60; }
echo "pool auto"; c2 X=1
echo "pool 16Mi"; c2 X=1 16777216
echo "pool 8Mi"; c2 X=1 8388608
echo "quarters 2"; c2 TRT_COMPACT_QUARTERS=2
echo "phases 2,12,8,3,12,8"; c2 TRT_PHASES=2,12,8,3,12,8
echo "phases 2,12,8,2,8,8"; c2 TRT_PHASES=2,12,8,2,8,8
echo "phases 3,12,8,2,12,8"; c2 TRT_PHASES=3,12,8,2,12,8
echo "phases 2,16,8,2,16,8"; c2 TRT_PHASES=2,16,8,2,16,8
echo "finish 64K"; c2 TRT_FINISH_BELOW=65536
echo "finish 256K"; c2 TRT_FINISH_BELOW=262144
echo "C3 256spp"; timeout 300 python tools/render_once.py 3 256 0 fast 2 0 2>&1 | tail -1 | cut -c1-60
echo "C4 64spp"; timeout 300 python tools/render_once.py 4 64 0 fast 2 0 2>&1 | tail -1 | cut -c1-60
