bash tools/gpu_ab.sh -t "X=1"
timeout 900 python bench.py --no-strong > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2e.json").read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"]["value"],"ms",d["ms_per_step"],"frac",d["roofline"]["frac"],d["roofline"]["bound"],"cadence",d["dropin_cadence"])
PY
TRT_ITER_LOG=gpurun_out/iterlog_r2e.txt timeout 300 python tools/render_once.py 2 64 0 fast 2 1 | tail -1 | cut -c1-200
