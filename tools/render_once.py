"""Render one config once (for profiling / tuning).
Usage: render_once.py config spp pool fast|ref [warm_spp] [time_kernels]
Prints ms per sample-per-pixel, Mrays/s and (with time_kernels=1) the per-kernel-family split."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import tryraytrace_b200 as trt
config, spp, pool, mode = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
warm = int(sys.argv[5]) if len(sys.argv) > 5 else 1
timed = int(sys.argv[6]) if len(sys.argv) > 6 else 0
import os, time
t0 = time.time()
sc = trt.HostScene.from_config(config, grid=int(os.environ.get('TRT_GRID', '0')))
host_s = time.time() - t0
cam, w, h = trt.config_camera(config)
ctx = trt.Context(0)
ctx.upload(sc)
acc = torch.zeros(w * h * 4, device="cuda")
torch.cuda.synchronize()  # the library works on its own stream
trav = trt.TRAVERSE_REF if mode == "ref" else trt.TRAVERSE_FAST
o = trt.default_opts(traversal=trav, pool_paths=pool)
if warm:
    ctx.render(acc, w, h, 1, warm, cam, o); ctx.synchronize()
ctx.reset_counters()
cnt = int(os.environ.get('TRT_COUNT', '0'))
ot = trt.default_opts(traversal=trav, pool_paths=pool, time_kernels=timed, count_rays=cnt)
ctx.render(acc, w, h, 1 + warm, spp, cam, ot); ctx.synchronize()
ms = ctx.last_render_ms()
c = ctx.counters()
rays = c["closest_rays"] + c["shadow_rays"]
line = f"ms/spp {ms / spp:.3f} Mrays/s {rays / ms / 1e3:.0f} rays/sample {rays / max(c['samples'], 1):.3f} iters {c['iterations']} replays {c['replays']}"
if timed:
    k = ctx.kernel_times()
    line += " | " + " ".join(f"{n[:-3]} {k[n]:.1f}" for n in ("regen_ms", "extend_ms", "shade_ms", "shadow_ms"))
if cnt:
    line += (f" | closest: nodes/ray {c['nodes_closest'] / max(c['closest_rays'], 1):.2f} tris/ray {c['tris_closest'] / max(c['closest_rays'], 1):.2f}"
             f" shadow: nodes/ray {(c['nodes_fetched'] - c['nodes_closest']) / max(c['shadow_rays'], 1):.2f} tris/ray {(c['tris_tested'] - c['tris_closest']) / max(c['shadow_rays'], 1):.2f}"
             f" rays c/s {c['closest_rays']}/{c['shadow_rays']} tree fraction c/s {c['tree_closest'] / max(c['closest_rays'], 1):.3f}/{c['tree_shadow'] / max(c['shadow_rays'], 1):.3f}")
inf = ctx.scene_info()
line += f" | builder {inf['builder']} build_ms {inf['build_ms']:.1f} wide_nodes {inf['n_wide_nodes']} depth {inf['wide_depth']} host_prep_s {host_s:.1f}"
print(line, "mean", float(acc.view(-1, 4)[:, :3].mean()) / (spp + warm))
