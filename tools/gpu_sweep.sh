#!/bin/bash
# timing of C2 (32 spp, pool 4 Mi) for each environment given as an argument ("A=1,B=2" -> A=1 B=2)
mkdir -p gpurun_out
for cfg in "$@"; do
  echo "== $cfg"
  env $(echo $cfg | tr ';' ' ') timeout 200 python tools/render_once.py 2 32 0 fast 2 1 2>&1 | tail -1
done
