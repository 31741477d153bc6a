"""First-contact probe for a GPU box: times the reference renderer and this library on a
config, prints ray counts and parity summary.  Usage: python tools/gpu_probe.py [config] [spp]"""
import sys, time, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
import tryraytrace_b200 as trt
import reflib as ref

config = int(sys.argv[1]) if len(sys.argv) > 1 else 2
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
pools = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1 << 20]
sc = trt.HostScene.from_config(config)
cam, w, h = trt.config_camera(config)
print(f"config {config}: {len(sc.objects)} tris, {len(sc.nodes)} nodes, {w}x{h}, spp {spp}", flush=True)
ctx = trt.Context(0)
t0 = time.time(); ctx.upload(sc); print("upload+relayout s", round(time.time() - t0, 4), ctx.scene_info(), flush=True)
ref.init_scene(sc)
n = w * h
acc_ref = torch.zeros(n * 4, device="cuda"); stage = torch.zeros(n * 4, device="cuda")
ref.render_frames(acc_ref, stage, w, h, 1, 2, cam, 0)  # warm-up
acc_ref.zero_()
ms_k = ref.render_frames(acc_ref, stage, w, h, 1, spp, cam, 0)
acc_ref2 = torch.zeros(n * 4, device="cuda")
ms_c = ref.render_frames(acc_ref2, stage, w, h, 1, spp, cam, 1)
counts = ref.full_counts(None, w, h, 1, min(spp, 4), cam)
rays_per_sample = (counts["closest_rays"] + counts["shadow_rays"]) / (n * min(spp, 4))
print("reference: kernel-only ms/spp", ms_k / spp, "cadence ms/spp", ms_c / spp, "rays/sample", rays_per_sample,
      "Mrays/s kernel-only", rays_per_sample * n * spp / ms_k / 1e3, counts, flush=True)
res = {}
for mode, name in ((trt.TRAVERSE_REF, "ref"), (trt.TRAVERSE_FAST, "fast")):
    for pool in pools:
        acc = torch.zeros(n * 4, device="cuda")
        o = trt.default_opts(traversal=mode, pool_paths=pool)
        ctx.render(acc, w, h, 1, 2, cam, o); ctx.synchronize(); acc.zero_()
        t0 = time.time(); ctx.render(acc, w, h, 1, spp, cam, o); ctx.synchronize(); wall = (time.time() - t0) * 1e3
        ms = ctx.last_render_ms()
        a, b = acc.cpu().numpy(), acc_ref.cpu().numpy()
        ia, ib = ref.tonemap(a, spp), ref.tonemap(b, spp)
        from gpu_common import psnr_8bit
        print(f"mine {name} pool {pool}: ms/spp {ms / spp:.4f} (wall {wall / spp:.4f}) speedup vs ref kernel {ms_k / ms:.2f}x "
              f"Mrays/s {rays_per_sample * n * spp / ms / 1e3:.1f} PSNR {psnr_8bit(ia, ib):.2f} mean {a.reshape(-1,4)[:, :3].mean():.6f} vs {b.reshape(-1,4)[:, :3].mean():.6f}", flush=True)
ctx.reset_counters()
acc = torch.zeros(n * 4, device="cuda")
ctx.render(acc, w, h, 1, min(spp, 4), cam, trt.default_opts(count_rays=1)); print("counters fast", ctx.counters(), flush=True)
ctx.reset_counters()
ctx.render(acc, w, h, 1, min(spp, 4), cam, trt.default_opts(count_rays=1, traversal=trt.TRAVERSE_REF)); print("counters ref ", ctx.counters(), flush=True)
