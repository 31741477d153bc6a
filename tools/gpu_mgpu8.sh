mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_n8_r2.json 2> gpurun_out/bench_n8_r2.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_n8_r2.json").read().strip().splitlines()[-1])
print("N=8 value",d["value"],"e2e",d["e2e"]["value"],"ms",d["ms_per_step"])
print("strong",json.dumps(d["strong_c4"])[:700])
PY
