#!/bin/bash
# usage: gpu_prof.sh <tag> <kernel-regex> <skip> <count> <spp> [env...]
mkdir -p gpurun_out
TAG=$1; RX=$2; SKIP=${3:-40}; CNT=${4:-6}; SPP=${5:-8}; shift 5
env "$@" python tools/render_once.py 2 $SPP 0 fast 0 > gpurun_out/plain_$TAG.log 2>&1 && \
env "$@" ncu --set full --clock-control none --import-source on -k regex:"$RX" -s $SKIP -c $CNT -o gpurun_out/prof_$TAG python tools/render_once.py 2 $SPP 0 fast 0 > gpurun_out/ncu_$TAG.log 2>&1
tail -1 gpurun_out/plain_$TAG.log; tail -2 gpurun_out/ncu_$TAG.log
