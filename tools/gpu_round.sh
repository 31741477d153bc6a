#!/bin/bash
# Round evidence: GPU tests, both bench arms, launch list + full ncu capture of the bench command, iteration log.
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; tail -c 300 gpurun_out/bench_ref_$TAG.json
timeout 1200 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"]["value"],"ms",d["ms_per_step"],"frac",d["roofline"]["frac"],d["roofline"]["bound"],"launches",d["gpu_launches"])
print("cadence",d["dropin_cadence"]); print("strong",d["strong_c4"])
PY
TRT_ITER_LOG=gpurun_out/iterlog_$TAG.txt timeout 300 python tools/render_once.py 2 64 0 fast 2 1 | tail -1 | cut -c1-200
bash tools/gpu_profile_round.sh $TAG 40
