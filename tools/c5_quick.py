"""C5 (10.1 M triangles) through the instanced upload (mesh parsed once, scene and tree built on the device), 4K.
usage: c5_quick.py [grid=40] [spp=4] [count=0]   -- prints ms/spp and Mrays/s; count=1 adds traversal statistics"""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import tryraytrace_b200 as trt
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 40
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 4
count = int(sys.argv[3]) if len(sys.argv) > 3 else 0
t0 = time.time()
unit = trt.load_obj(str(trt.ASSET_DIR / "teapot.obj"))
extra, _ = trt.create_scene(5, grid=0 if False else 1)  # floor + light come first in the factory's array
extra = extra[:2]
inst = np.array([(-175.0 + 9.0 * ix, 0.0, 60.0 - 9.0 * iz, 1.2) for iz in range(grid) for ix in range(grid)], dtype=np.float32)
host_s = time.time() - t0
cam, w, h = trt.config_camera(5)
ctx = trt.Context(0)
t0 = time.time()
ctx.upload_instanced(extra, unit, inst, [1])
ctx.synchronize()
up_s = time.time() - t0
acc = torch.zeros(w * h * 4, device="cuda")
torch.cuda.synchronize()
ctx.render(acc, w, h, 1, 1, cam, trt.default_opts()); ctx.synchronize()
ctx.reset_counters()
ctx.render(acc, w, h, 2, spp, cam, trt.default_opts(count_rays=count)); ctx.synchronize()
ms = ctx.last_render_ms()
c = ctx.counters()
rays = c["closest_rays"] + c["shadow_rays"]
info = ctx.scene_info()
print(f"C5 grid {grid}: ms/spp {ms / spp:.3f} Mrays/s {rays / ms / 1e3:.0f} rays/sample {rays / max(c['samples'], 1):.3f} iters {c['iterations']} "
      f"| host prepare {host_s:.3f} s, upload+build wall {up_s:.3f} s (build {info['build_ms']:.1f} ms), {info['n_objects']} objects, "
      f"{info['n_wide_nodes']} nodes of {info['wide_node_bytes']} B")
