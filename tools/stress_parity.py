"""Wider parity sweep than the test suite runs: first-hit ids and d_min of the fast traversal against the
unmodified reference kernel at the full resolution of every config for many frame seeds, and
random secondary / shadow rays against the reference-order traversal.  Prints one line per case.
Usage: stress_parity.py [frames=16] [rays=4000000]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import torch
import tryraytrace_b200 as trt
import reflib

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n_rays = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
ctx = trt.Context(0)
bad_total = 0
for config in (1, 2, 3, 4):
    sc = trt.HostScene.from_config(config)
    cam, w, h = trt.config_camera(config)
    for builder in (trt.BUILD_HOST_SAH, trt.BUILD_DEVICE_LBVH):
        ctx.upload(sc, builder=builder)
        reflib.init_scene(sc)
        n = w * h
        bad_id = bad_t = replays = 0
        for frame in range(1, frames + 1):
            want = torch.from_numpy(reflib.first_hit_ids(w, h, frame, cam)).cuda()
            rt = torch.zeros(n, device="cuda")
            reflib.primary_counts(w, h, frame, cam, None, rt)
            mid = torch.zeros(n, dtype=torch.int32, device="cuda")
            mt = torch.zeros(n, device="cuda")
            amb = torch.zeros(n, dtype=torch.int32, device="cuda")
            torch.cuda.synchronize()
            ctx.trace_primary(w, h, frame, cam, trt.TRAVERSE_FAST, d_id=mid, d_t=mt, d_entered=amb)
            bad_id += int((mid != want).sum()); bad_t += int((mt.view(torch.int32) != rt.view(torch.int32)).sum())
            replays += int(amb.sum())
        g = torch.Generator(device="cuda").manual_seed(100 + config)
        rays = torch.zeros(n_rays, 8, device="cuda")
        rays[:, 0] = torch.rand(n_rays, generator=g, device="cuda") * 100
        rays[:, 1] = torch.rand(n_rays, generator=g, device="cuda") * 100
        rays[:, 2] = torch.rand(n_rays, generator=g, device="cuda") * 300
        d = torch.randn(n_rays, 3, generator=g, device="cuda")
        rays[:, 3:6] = d / d.norm(dim=1, keepdim=True)
        rays[:, 6] = torch.rand(n_rays, generator=g, device="cuda") * 150 + 1
        rays[: n_rays // 100, 4] = 0.0
        torch.cuda.synchronize()  # the library works on its own stream
        out = {}
        for mode in (trt.TRAVERSE_REF, trt.TRAVERSE_FAST):
            i = torch.zeros(n_rays, dtype=torch.int32, device="cuda"); t = torch.zeros(n_rays, device="cuda")
            o = torch.zeros(n_rays, dtype=torch.int32, device="cuda")
            torch.cuda.synchronize()
            ctx.trace_closest(rays, n_rays, mode, i, t); ctx.trace_shadow(rays, n_rays, mode, o)
            out[mode] = (i, t, o)
        a, b = out[trt.TRAVERSE_REF], out[trt.TRAVERSE_FAST]
        sec_id = int((a[0] != b[0]).sum()); sec_t = int((a[1].view(torch.int32) != b[1].view(torch.int32)).sum())
        sec_o = int((a[2] != b[2]).sum())
        bad_total += bad_id + bad_t + sec_id + sec_t + sec_o
        print(f"C{config} {w}x{h} builder {builder}: {frames} frames x {n} pixels: id mismatches {bad_id}, d_min mismatches {bad_t}, "
              f"replays {replays}; {n_rays} random rays: closest id/t mismatches {sec_id}/{sec_t}, shadow mismatches {sec_o}", flush=True)
print("TOTAL MISMATCHES", bad_total)
