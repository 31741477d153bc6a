"""Experiment: one render split over K contexts of the same GPU, driven by K host threads.
Usage: dual_ctx.py config spp k [pool]
The frames of the job are split by sample index (frame_stride = k) exactly like the multi-GPU split, so the
union of the k parts is the single-context seed set.  Prints wall/device time of the split render and of the
single-context render, and the largest relative difference of the two images."""
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import tryraytrace_b200 as trt

config, spp, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
pool = int(sys.argv[4]) if len(sys.argv) > 4 else 0
sc = trt.HostScene.from_config(config)
cam, w, h = trt.config_camera(config)
ctxs = [trt.Context(0) for _ in range(k)]
for c in ctxs:
    c.upload(sc)
accs = [torch.zeros(w * h * 4, device="cuda") for _ in range(k)]
one = torch.zeros(w * h * 4, device="cuda")
torch.cuda.synchronize()
o = trt.default_opts(pool_paths=pool)


def part(i, first):
    ctxs[i].render(accs[i], w, h, first + i, spp // k, cam, o, frame_stride=k)
    ctxs[i].synchronize()


def split(first):
    th = [threading.Thread(target=part, args=(i, first)) for i in range(k)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    return (time.perf_counter() - t0) * 1e3


def single(first):
    t0 = time.perf_counter()
    ctxs[0].render(one, w, h, first, spp, cam, o)
    ctxs[0].synchronize()
    return (time.perf_counter() - t0) * 1e3


for a in accs:
    a.zero_()
torch.cuda.synchronize()
split(1)
single(1)  # warm both shapes
res = []
for rep in range(3):
    for a in accs:
        a.zero_()
    one.zero_()
    torch.cuda.synchronize()
    ts = split(1)
    t1 = single(1)
    res.append((ts, t1))
total = sum(accs)
d = float((total - one).abs().max() / one.abs().max())
print(f"config {config} spp {spp} k {k}: split wall ms {[round(r[0], 2) for r in res]} single wall ms {[round(r[1], 2) for r in res]} "
      f"ms/spp split {min(r[0] for r in res) / spp:.3f} single {min(r[1] for r in res) / spp:.3f} max rel diff {d:.2e}")
