mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu_final.log
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_final.json 2>/dev/null
timeout 1200 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?"
python - <<PY
import json
r=json.loads(open("gpurun_out/bench_ref_final.json").read().strip().splitlines()[-1])
d=json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
print("ref",r["value"],"value",d["value"],"e2e",d["e2e"]["value"],"ratio",d["e2e"]["value"]/r["value"],"ms",d["ms_per_step"],"frac",d["roofline"]["frac"],"cadence",d["dropin_cadence"]["mrays_per_s"])
PY
