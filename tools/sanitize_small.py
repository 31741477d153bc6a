"""compute-sanitizer target: a small render + the parity entry points in FAST mode."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import tryraytrace_b200 as trt
config = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sc = trt.HostScene.from_config(config)
cam, w, h = trt.config_camera(config, 256, 144)
ctx = trt.Context(0)
ctx.upload(sc)
acc = torch.zeros(w * h * 4, device="cuda")
ctx.render(acc, w, h, 1, 3, cam, trt.default_opts(pool_paths=1 << 15)); ctx.synchronize()
ids = torch.zeros(w * h, dtype=torch.int32, device="cuda")
ctx.trace_primary(w, h, 1, cam, trt.TRAVERSE_FAST, d_id=ids)
print("ok mean", float(acc.view(-1, 4)[:, :3].mean()) / 3, "hit frac", float((ids >= 0).float().mean()))
