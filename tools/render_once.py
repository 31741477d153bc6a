"""Render one config once (for profiling).  Usage: render_once.py config spp pool fast|ref [warm_spp]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import tryraytrace_b200 as trt
config, spp, pool, mode = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
warm = int(sys.argv[5]) if len(sys.argv) > 5 else 1
sc = trt.HostScene.from_config(config)
cam, w, h = trt.config_camera(config)
ctx = trt.Context(0)
ctx.upload(sc)
acc = torch.zeros(w * h * 4, device="cuda")
o = trt.default_opts(traversal=trt.TRAVERSE_REF if mode == "ref" else trt.TRAVERSE_FAST, pool_paths=pool)
if warm:
    ctx.render(acc, w, h, 1, warm, cam, o); ctx.synchronize()
ctx.render(acc, w, h, 1, spp, cam, o); ctx.synchronize()
print("ms/spp", ctx.last_render_ms() / spp, "mean", float(acc.view(-1, 4)[:, :3].mean()) / (spp + warm))
