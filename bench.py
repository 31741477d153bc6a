#!/usr/bin/env python
"""bench.py -- headline benchmark of the path-tracing hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], SURVEY 8d "C2"): room + teapot.obj (6 327 triangles),
1920x1080, 64 samples per pixel per step, diffuse + emissive, reference constants
(MAX_DEPTH 30, Russian roulette after depth 3), RNG streams of the reference
(seed 1984+frame, subsequence = pixel).  Scene and camera are synthetic-but-fixed; no
dataset is involved.  A "step" is one progressive pass: 64 frames of samples added to the
accumulation buffer.  With N GPUs every rank renders 64 frames of its own (frame seeds
interleaved with stride N, the sample split of SURVEY 8e) and the pass ends with ONE
all-reduce of the accumulation buffer over NCCL -- weak scaling.

Metric: Mrays/s = (closest-hit + shadow BVH queries) / time; samples/s is reported beside it.

  value   device-resident: accumulation buffer stays in HBM, timed with CUDA events on the
          stream the kernels are launched on (torch's current stream, handed to the library).
  e2e     the same pass through the host-buffer entry point trt_render_to_host(): per step the
          camera/options come from host memory, the device buffer is cleared, and the 33 MB
          result is copied to pinned host memory inside the timed region.
  roofline  the dominant kernel (closest-hit traversal): algorithmic bytes per launch / its
          average launch time, against the measured HBM copy bandwidth (MEASURED_PEAKS.json).
  cpu_baseline  the CPU restatement of the reference kernel (oracle/, OpenMP over pixels) on a
          bounded sample of the same workload.  (The reference has no CPU renderer.)

--impl reference runs the UNMODIFIED reference renderer (oracle/_ref, reference
src/renderer.cu compiled for sm_100 with its own flags) through its own entry points and
main-loop cadence (launch + D2D snapshot + cudaDeviceSynchronize per sample, reference
src/main.cpp:181-192), same scene, camera, seeds, metric and config.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

CONFIG = 2
SPP = 64


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--spp", type=int, default=SPP)
    ap.add_argument("--pool", type=int, default=0, help="wavefront pool size (paths in flight); 0 = the library's rule (a quarter of the step's samples, 256 Ki .. 32 Mi)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """Samples nvidia-smi during the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def committed_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, or None."""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get("k_extend_fast_dram_bytes_per_launch")
        except Exception:
            return None
    return None


def cpu_baseline(trt, scene, cam, w, h, target_s=12.0):
    """CPU restatement (oracle/cpu_oracle.cpp) on a bounded sample: 1 spp over as many image rows
    as fit ~target_s of all host cores."""
    import numpy as np
    import reflib
    if not reflib.cpu_available():
        return {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "port", "sample": "oracle library not built"}
    L = reflib.cpu()
    li = np.ascontiguousarray(scene.lights, dtype=np.int32)

    def run(row0, row1, frames=1):
        acc = np.zeros(w * h * 4, dtype=np.float32)
        tot = np.zeros(5, dtype=np.uint64)
        t0 = time.perf_counter()
        L.oracle_render(scene.objects.ctypes.data_as(C.c_void_p), scene.nodes.ctypes.data_as(C.c_void_p),
                        li.ctypes.data_as(C.c_void_p), len(li), cam.ctypes.data_as(C.c_void_p), w, h, 1984, 1, frames,
                        30, 3, None, None, None, 0, row0, row1, acc.ctypes.data_as(C.c_void_p),
                        tot.ctypes.data_as(C.c_void_p), 0)
        return time.perf_counter() - t0, int(tot[0] + tot[1])

    run(0, 4)  # builds the skip-ahead matrices
    mid = h // 2
    dt, rays = run(mid, mid + 16)
    rows = int(max(16, min(h, 16 * target_s / max(dt, 1e-3))))
    frames = 1 if rows < h else int(max(1, min(16, target_s / max(dt * h / 16, 1e-3))))
    row0 = max(0, mid - rows // 2)
    dt, rays = run(row0, min(h, row0 + rows), frames)
    return {"value": rays / dt / 1e6, "unit": "Mrays/s", "cores": int(L.oracle_max_threads()), "kind": "port",
            "sample": f"{frames} spp, image rows {row0}..{min(h, row0 + rows)} of {h} ({rays} rays, {dt:.1f} s), "
                      f"CPU restatement of reference renderer.cu (the reference has no CPU renderer)"}


def main():
    args = parse_args()
    import numpy as np
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = max(args.gpus, world)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod

    import tryraytrace_b200 as trt

    scene = trt.HostScene.from_config(CONFIG)
    cam, w, h = trt.config_camera(CONFIG)
    pixels = w * h
    workload = (f"C2 room+teapot.obj ({len(scene.objects)} triangles) {w}x{h}, {args.spp} spp per step, diffuse+emissive, "
                f"MAX_DEPTH 30, RR after depth 3")
    config = {"workload": workload, "spp_per_step": args.spp, "width": w, "height": h,
              "sharding": f"sample index, stride {world}, one all-reduce of the accumulation buffer per step" if world > 1
              else "single GPU", "l2": "no explicit L2 flush: each step streams the wavefront pool "
              "(132 B of path state per slot: 4.4 GB at the 32 Mi slots the library picks for this step) and the 33 MB "
              "accumulation buffer, both larger than or comparable to the 126 MB L2"}

    # ---------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        import reflib
        if not reflib.available():
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libtrt_ref.so was not built"}))
            return
        torch.cuda.set_device(0)
        reflib.init_scene(scene)
        acc = torch.zeros(pixels * 4, device="cuda")
        stage = torch.zeros(pixels * 4, device="cuda")
        host = torch.zeros(pixels * 4).pin_memory()
        sampler = ClockSampler(0)
        for s in range(args.warmup):
            reflib.render_frames(acc, stage, w, h, 1 + s * args.spp, args.spp, cam, 1)
        torch.cuda.synchronize()
        sampler.start()
        t0 = time.perf_counter()
        ms_dev = 0.0
        for s in range(args.steps):
            first = 1 + (args.warmup + s) * args.spp
            ms_dev += reflib.render_frames(acc, stage, w, h, first, args.spp, cam, 1)
            host.copy_(stage, non_blocking=True)  # what the display worker does (reference src/pipeline.cpp:45)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        clocks = sampler.stop()
        rays = 0
        for s in range(args.steps):  # instrumented restatement, untimed, same seeds
            c = reflib.full_counts(None, w, h, 1 + (args.warmup + s) * args.spp, args.spp, cam)
            rays += c["closest_rays"] + c["shadow_rays"]
        v = rays / wall / 1e6
        line = {"impl": "reference", "metric": "Mrays/s", "value": v, "unit": "Mrays/s", "n_gpus": 1,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall * 1e3 / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "samples_per_s": pixels * args.spp * args.steps / wall,
                "kernel_only_mrays_per_s": rays / (ms_dev * 1e-3) / 1e6,
                "device": "gpu: unmodified reference src/renderer.cu, nvcc -O3 -arch=sm_100 --use_fast_math, "
                          "main-loop cadence (launch + D2D snapshot + device sync per sample)",
                "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": 0, "kind": "reference",
                                 "sample": "whole workload on the GPU: the reference's implementation of this path is "
                                           "a CUDA kernel, it has no CPU renderer"},
                "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "clocks": clocks}
        print(json.dumps(line))
        return

    # ---------------------------------------------------------------- our arm
    torch.cuda.set_device(local)
    if dist is not None:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = trt.Context(local)
    ctx.upload(scene)
    stream = torch.cuda.Stream()  # kernels, the all-reduce and the timing events share this stream
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    opts = trt.default_opts(pool_paths=args.pool)
    # timed region: events only around the dominant kernel (roofline); the per-kernel split of a step
    # comes from one extra step with all marks on, outside the timed region
    timed_opts = trt.default_opts(pool_paths=args.pool, time_kernels=2)
    split_opts = trt.default_opts(pool_paths=args.pool, time_kernels=1)
    acc = torch.zeros(pixels * 4, device="cuda")

    from tryraytrace_b200.sharding import render_pass_sharded

    def step(first_seed, o):
        # weak scaling: the pass has spp*world frames, rank r renders first+r, first+r+world, ...
        acc.zero_()
        render_pass_sharded(ctx, acc, w, h, first_seed, args.spp * world, cam, o, dist)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for s in range(args.warmup):
        step(1 + s * args.spp * world, opts)
    barrier()
    ctx.reset_counters()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kt = {"extend_ms": 0.0, "iterations": 0}
    barrier()
    ev0.record(stream)
    for s in range(args.steps):
        step(1 + (args.warmup + s) * args.spp * world, timed_opts)
        k = ctx.kernel_times()
        for key in kt:
            kt[key] += k[key]
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    cnt = ctx.counters()
    rays_local = cnt["closest_rays"] + cnt["shadow_rays"]
    t = torch.tensor([ms, float(rays_local), float(cnt["samples"]), float(cnt["kernel_launches"])], device="cuda",
                     dtype=torch.float64)
    if dist is not None:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms = float(tmax[0])
    rays, samples, launches = float(t[1]), float(t[2]), int(t[3])
    value = rays / (ms * 1e-3) / 1e6

    # per-kernel split of one more step (all marks on; regenerate overlaps the shadow kernel)
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record(stream)
    step(1 + (args.warmup + args.steps) * args.spp * world, split_opts)
    ev3.record(stream)
    torch.cuda.synchronize()
    split = ctx.kernel_times()
    split_ms = ev2.elapsed_time(ev3)

    # ---- end to end through the host-buffer entry point (rank-local render + D2H)
    host = torch.zeros(pixels * 4).pin_memory()
    ctx.render_to_host(host, w, h, 1 + rank, args.spp, cam, opts, frame_stride=world)  # warm
    barrier()
    ctx.reset_counters()
    t0 = time.perf_counter()
    for s in range(args.steps):
        ctx.render_to_host(host, w, h, 1 + (args.warmup + s) * args.spp * world + rank, args.spp, cam, opts,
                           frame_stride=world)
        if dist is not None:  # the reduced image is what a consumer reads; reduce the device copy, then fetch
            dist.barrier()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    c2 = ctx.counters()
    e2e_rays = torch.tensor([float(c2["closest_rays"] + c2["shadow_rays"]), e2e_s], device="cuda", dtype=torch.float64)
    if dist is not None:
        mx = e2e_rays.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_rays, op=dist.ReduceOp.SUM)
        e2e_s = float(mx[1])
    e2e_value = float(e2e_rays[0]) / e2e_s / 1e6

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel: one counting step with the same seeds (deterministic)
    ctx.reset_counters()
    acc.zero_()
    ctx.render(acc, w, h, 1 + args.warmup * args.spp * world + rank, args.spp, cam,
               trt.default_opts(pool_paths=args.pool, count_rays=1), frame_stride=world)
    cc = ctx.counters()
    info = ctx.scene_info()
    # algorithmic bytes of closest-hit traversal per step: node records + triangle records actually
    # fetched, plus the ray read (32 B) and hit write (8 B) of every query ("roofline")
    bytes_step = (cc["nodes_closest"] * info["wide_node_bytes"] + cc["tris_closest"] * info["tri_record_bytes"]
                  + cc["closest_rays"] * 40)
    extend_s = kt["extend_ms"] * 1e-3 / args.steps
    peak, peak_src = measured_peaks()
    achieved = bytes_step / extend_s / 1e9 if extend_s > 0 else None
    launches_per_step = kt["iterations"] / args.steps if args.steps else 0
    traffic = committed_traffic()
    # the other two terms of SURVEY 8(d)'s bound, from the same counters: FP32 work of the slab and
    # triangle tests (48 flop per 4-wide node, 51 per triangle test incl. the root-level list) against
    # the FP32 peak of 148 SMs x 128 lanes x 2 x clock, and the kernel's real DRAM traffic (ncu)
    top_tris = cc["closest_rays"] * info["n_top_prims"]
    flop_step = cc["nodes_closest"] * 48.0 + (cc["tris_closest"] + top_tris) * 51.0
    clock_hz = (clocks or {}).get("sm_mhz") or 1965.0
    fp32_peak = 148 * 128 * 2 * clock_hz * 1e6
    roofline = {"kernel": "k_extend_fast (closest-hit traversal)", "bound": "hbm", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak if achieved else None, "traffic": traffic,
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_step / launches_per_step if launches_per_step else None,
                "avg_launch_us": extend_s * 1e6 / launches_per_step if launches_per_step else None,
                "per_ray": {"nodes": cc["nodes_closest"] / max(cc["closest_rays"], 1),
                            "tris": cc["tris_closest"] / max(cc["closest_rays"], 1),
                            "root_level_tris": info["n_top_prims"],
                            "node_bytes": info["wide_node_bytes"], "tri_bytes": info["tri_record_bytes"]},
                "fp32": {"flop_per_step": flop_step, "achieved_tflops": flop_step / extend_s / 1e12 if extend_s > 0 else None,
                         "peak_tflops": fp32_peak / 1e12,
                         "frac": flop_step / extend_s / fp32_peak if extend_s > 0 else None},
                "dram_frac": (traffic / (extend_s / launches_per_step) / 1e9 / peak) if (traffic and launches_per_step and extend_s > 0) else None,
                "note": "algorithmic bytes are node + triangle records + ray/hit stream; the scene (<1 MB) is served from "
                        "shared memory / L1 / L2, so `achieved` is a logical bandwidth: real DRAM traffic (`traffic`, ncu) is "
                        "the ray/hit stream only (`dram_frac` of HBM peak) and the kernel is bound by instruction issue "
                        "(ncu: ~70 % issue-slot utilisation, 22 of 32 lanes per instruction; profiles/)",
                "kernel_share_of_step": {k: split[k] / max(split_ms, 1e-9) for k in ("regen_ms", "extend_ms", "shade_ms", "shadow_ms")},
                "kernel_share_note": "one extra step with events at every kernel boundary; regenerate runs beside the shadow kernel, so the shares add up to more than 1"}

    # rank 0 at N=1 only: under torchrun the host cores are shared (and OMP_NUM_THREADS is forced to 1)
    cpu = None if (args.no_cpu_baseline or world > 1) else cpu_baseline(trt, scene, cam, w, h)

    # first-hit id parity against the unmodified reference kernel, when the oracle library travelled
    id_match = None
    try:
        import reflib
        if reflib.available():
            reflib.init_scene(scene)
            want = reflib.first_hit_ids(w, h, 1, cam)
            ids = torch.zeros(pixels, dtype=torch.int32, device="cuda")
            ctx.trace_primary(w, h, 1, cam, trt.TRAVERSE_FAST, d_id=ids)
            id_match = float((ids.cpu().numpy() == want).mean())
    except Exception as e:  # parity is reported, never allowed to break the bench line
        id_match = f"not checked: {e}"

    line = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "samples_per_s": samples / (ms * 1e-3), "rays_per_sample": rays / max(samples, 1),
            "first_hit_id_match": id_match,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 80 + 32,
                    "d2h_bytes_per_step": pixels * 16, "ms_per_step": e2e_s * 1e3 / args.steps,
                    "api": "trt_render_to_host (scene resident on the device; camera/options from host, "
                           "accumulation buffer cleared on device, result copied to pinned host memory)"},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "pool_paths": args.pool if args.pool else "auto"}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
