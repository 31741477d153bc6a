"""Throughput of every named configuration (BASELINE.json configs C1..C4; C5 has its own tool) on one
GPU: this library (device-resident, CUDA events) next to the unmodified reference kernel with its
own cadence, same scene, camera and seeds.  Prints one JSON object.
Usage: all_configs.py [spp_cap=64]"""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import tryraytrace_b200 as trt
import reflib

cap = int(sys.argv[1]) if len(sys.argv) > 1 else 64
out = {}
ctx = trt.Context(0)
for config in (1, 2, 3, 4):
    sc = trt.HostScene.from_config(config)
    cam, w, h = trt.config_camera(config)
    spp = min(trt.CONFIGS[config]["spp"], cap)
    ctx.upload(sc)
    acc = torch.zeros(w * h * 4, device="cuda")
    torch.cuda.synchronize()  # the library works on its own stream
    o = trt.default_opts()  # pool sized to the job by the library
    ctx.render(acc, w, h, 1, 4, cam, o); ctx.synchronize()
    ctx.reset_counters()
    ctx.render(acc, w, h, 5, spp, cam, o); ctx.synchronize()
    ms = ctx.last_render_ms()
    c = ctx.counters()
    rays = c["closest_rays"] + c["shadow_rays"]
    row = {"scene": trt.CONFIGS[config]["name"], "triangles": int(len(sc.objects)), "width": w, "height": h, "spp": spp,
           "rays_per_sample": round(rays / c["samples"], 3), "ours_ms_per_spp": round(ms / spp, 3),
           "ours_mrays_per_s": round(rays / ms / 1e3, 1), "ours_samples_per_s": round(c["samples"] / ms * 1e3)}
    if reflib.available():
        reflib.init_scene(sc)
        a2, st = torch.zeros_like(acc), torch.zeros_like(acc)
        torch.cuda.synchronize()
        rspp = min(spp, 16)
        reflib.render_frames(a2, st, w, h, 1, 2, cam, 1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reflib.render_frames(a2, st, w, h, 5, rspp, cam, 1)
        torch.cuda.synchronize()
        rms = (time.perf_counter() - t0) * 1e3
        row.update(reference_ms_per_spp=round(rms / rspp, 3), reference_mrays_per_s=round(rays / spp * rspp / rms / 1e3, 1),
                   speedup=round((rms / rspp) / (ms / spp), 2), reference_spp=rspp)
    out[f"C{config}"] = row
print(json.dumps(out))
