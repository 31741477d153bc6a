mkdir -p gpurun_out
nvidia-smi -L | head -4
TRT_EXPECT_GPUS=2 timeout 900 python -m pytest tests/test_mgpu.py tests/test_dropin.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2_r2.json 2> gpurun_out/bench_n2_r2.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_n2_r2.json").read().strip().splitlines()[-1])
print("N=2 value",d["value"],"e2e",d["e2e"]["value"],"ms",d["ms_per_step"])
print("strong",json.dumps(d["strong_c4"])[:900])
PY
tail -3 gpurun_out/bench_n2_r2.err
