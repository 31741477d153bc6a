python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for t in 768 1024 512; do echo "== C5 threads $t"; TRT_FAST_THREADS=$t timeout 300 python tools/c5_quick.py 40 4 2>&1 | tail -1 | cut -c1-110; done
echo "== C5 refill 24 / phases 6,12,8"; TRT_REFILL=24 timeout 300 python tools/c5_quick.py 40 4 2>&1 | tail -1 | cut -c1-80
