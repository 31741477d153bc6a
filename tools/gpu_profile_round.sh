#!/bin/bash
# Round profile: launch list of the bench command + full capture of the dominant kernel.
# usage: gpu_profile_round.sh <tag> [matching launches to skip before the full capture]
TAG=${1:-r1}; SKIP=${2:-60}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/bench_plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
tail -1 gpurun_out/bench_plain_$TAG.log | cut -c1-300
$CMD > gpurun_out/bench_plain2_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_trace_fast|k_extend_fast|k_shadow_fast|k_shade|k_refill|k_regen' -s $SKIP -c 8 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
