#!/bin/bash
for p in 4194304 6291456 8388608; do echo "== pool $p"; timeout 200 python tools/render_once.py 2 64 $p fast 2 1 2>&1 | tail -1; done
