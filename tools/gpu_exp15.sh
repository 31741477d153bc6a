echo "== C2 1 spp per call, pool sweep"
for p in 0 262144 524288 1048576 2097152 4194304; do echo "pool $p"; timeout 120 python tools/render_once.py 2 1 $p fast 1 0 2>&1 | tail -1 | cut -c1-70; done
echo "== C1 16 spp, pool sweep"
for p in 0 1048576 2097152 4194304 8388608; do echo "pool $p"; timeout 120 python tools/render_once.py 1 16 $p fast 2 0 2>&1 | tail -1 | cut -c1-70; done
echo "== C2 4 spp and 16 spp, pool sweep"
for p in 0 2097152 4194304 8388608; do echo "4spp pool $p"; timeout 120 python tools/render_once.py 2 4 $p fast 2 0 2>&1 | tail -1 | cut -c1-70; done
for p in 0 8388608 16777216 33554432; do echo "16spp pool $p"; timeout 120 python tools/render_once.py 2 16 $p fast 2 0 2>&1 | tail -1 | cut -c1-70; done
