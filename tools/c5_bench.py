"""C5 (BASELINE config 5): the instanced teapot field, built on the device from the object array.
Usage: c5_bench.py [grid=40] [spp=4] [ref_build=0|1] [width height]
Reports the device BVH build time (CUDA events) next to the reference's host BVH::build
(single thread, reference src/bvh.cpp:32) when ref_build=1, and the render throughput."""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import tryraytrace_b200 as trt

grid = int(sys.argv[1]) if len(sys.argv) > 1 else 40
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ref_build = int(sys.argv[3]) if len(sys.argv) > 3 else 0
t0 = time.time()
objs, tex = trt.create_scene(5, grid=grid)
t_create = time.time() - t0
cam, w, h = trt.config_camera(5, *(int(x) for x in sys.argv[4:6])) if len(sys.argv) > 5 else trt.config_camera(5)
out = {"config": f"C5 {grid}x{grid} teapots", "triangles": int(len(objs)), "width": w, "height": h, "spp": spp,
       "host_scene_create_s": round(t_create, 2)}
ctx = trt.Context(0)
lights = trt.collect_lights(objs)
# f3: the same scene through the instanced upload -- the mesh is parsed ONCE on the host, the placements are a
# (n, 4) array, the 10 M object records are written by a kernel and the tree is built on the device
t0 = time.time()
unit = trt.load_obj(str(trt.ASSET_DIR / "teapot.obj"))
extra = objs[:2]
inst = np.array([(-175.0 + 9.0 * ix, 0.0, 60.0 - 9.0 * iz, 1.2) for iz in range(grid) for ix in range(grid)], dtype=np.float32)
t_parse = time.time() - t0
t0 = time.time()
ctx.upload_instanced(extra, unit, inst, lights)
ctx.synchronize()
out["instanced_host_prepare_s"] = round(t_parse, 4)
out["instanced_upload_plus_build_wall_s_first"] = round(time.time() - t0, 3)  # first build of the process: allocator start-up
t0 = time.time()
ctx.upload_instanced(extra, unit, inst, lights)
ctx.synchronize()
out["instanced_upload_plus_build_wall_s"] = round(time.time() - t0, 3)
out["instanced_objects_equal_host_array"] = bool(ctx.get_objects().tobytes() == np.ascontiguousarray(objs).tobytes()) if grid <= 12 else "not compared (1.1 GB)"
t0 = time.time()
ctx.init_scene_data(objs, [], None, lights, builder=trt.BUILD_DEVICE_LBVH)
out["upload_plus_build_wall_s"] = round(time.time() - t0, 3)
info = ctx.scene_info()
out["device_build_ms"] = round(info["build_ms"], 2)
t0 = time.time()
ctx.init_scene_data(objs, [], None, lights, builder=trt.BUILD_DEVICE_LBVH)  # second build: allocator warm
out["device_build_ms_warm"] = round(ctx.scene_info()["build_ms"], 2)
out["wide_nodes"], out["wide_depth"] = info["n_wide_nodes"], info["wide_depth"]
acc = torch.zeros(w * h * 4, device="cuda")
torch.cuda.synchronize()  # the library works on its own stream
o = trt.default_opts()  # pool sized to the job by the library
ctx.render(acc, w, h, 1, 1, cam, o); ctx.synchronize()
ctx.reset_counters()
ctx.render(acc, w, h, 2, spp, cam, o); ctx.synchronize()
ms = ctx.last_render_ms()
c = ctx.counters()
rays = c["closest_rays"] + c["shadow_rays"]
out.update(ms_per_spp=round(ms / spp, 3), mrays_per_s=round(rays / ms / 1e3, 1), rays_per_sample=round(rays / max(c["samples"], 1), 3),
           samples_per_s=round(c["samples"] / ms * 1e3))
if ref_build:
    t0 = time.time()
    sc = trt.HostScene(objs, tex)  # BVH::build reorders the objects, then the light list
    out["reference_host_bvh_build_s"] = round(time.time() - t0, 2)
    out["build_speedup_vs_reference_host"] = round(out["reference_host_bvh_build_s"] * 1e3 / max(out["device_build_ms_warm"], 1e-3), 1)
    # the unmodified reference kernel on its own tree, same camera and seeds (test oracle, oracle/_ref)
    sys.path.insert(0, str(ROOT / "tests"))
    import reflib
    if reflib.available():
        ctx.close()
        reflib.init_scene(sc)
        stage = torch.zeros_like(acc)
        torch.cuda.synchronize()
        reflib.render_frames(acc, stage, w, h, 1, 1, cam, 1)
        torch.cuda.synchronize()
        t0 = time.time()
        reflib.render_frames(acc, stage, w, h, 2, spp, cam, 1)
        torch.cuda.synchronize()
        ref_ms = (time.time() - t0) * 1e3
        out["reference_kernel_ms_per_spp"] = round(ref_ms / spp, 2)
        out["reference_kernel_mrays_per_s"] = round(rays / ref_ms / 1e3, 1)  # same seeds => same ray counts
        out["render_speedup_vs_reference_kernel"] = round(out["reference_kernel_ms_per_spp"] / out["ms_per_spp"], 2)
print(json.dumps(out))
