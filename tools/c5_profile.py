"""ncu target: the 10 M-triangle field (C5), standalone device build, 2 spp -- for a --set full capture of the
traversal kernels on a scene far larger than L2."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import tryraytrace_b200 as trt
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 40
objs, tex = trt.create_scene(5, grid=grid)
cam, w, h = trt.config_camera(5)
ctx = trt.Context(0)
ctx.init_scene_data(objs, [], None, trt.collect_lights(objs), builder=trt.BUILD_DEVICE_LBVH)
acc = torch.zeros(w * h * 4, device="cuda")
torch.cuda.synchronize()
ctx.render(acc, w, h, 1, 2, cam, trt.default_opts(pool_paths=4 << 20)); ctx.synchronize()
print("ok", float(acc.view(-1, 4)[:, :3].mean()) / 2)
