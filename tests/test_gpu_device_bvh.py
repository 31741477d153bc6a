"""GPU tier: the device-side BVH builder (csrc/kernels/bvh_build.cu, LBVH -> 4-wide collapse).

The builder replaces the host re-layout for large scenes (SURVEY 8f.1 / BASELINE config C5).  Its
output must satisfy the same contract as the host builder's: with the reference node array present
the fast traversal over it returns the reference's first-hit ids, d_min and shadow bits exactly;
without it (standalone build from the object array) results agree with the reference-order
traversal except where the reference itself is order dependent (a few rays per million).
"""
import numpy as np
import pytest
import torch

from gpu_common import SceneCache, dev_zeros

pytestmark = pytest.mark.gpu

CASES = {1: (640, 480), 2: (960, 540), 4: (960, 540)}


@pytest.fixture(scope="module")
def scenes(trt, assets):
    return SceneCache(trt, assets)


@pytest.fixture(scope="module")
def ctx(trt):
    c = trt.Context(0)
    yield c
    c.close()


def random_rays(n, seed, lo=(0, 0, 0), hi=(100, 100, 300)):
    g = torch.Generator(device="cuda").manual_seed(seed)
    rays = torch.zeros(n, 8, device="cuda")
    for k in range(3):
        rays[:, k] = torch.rand(n, generator=g, device="cuda") * (hi[k] - lo[k]) + lo[k]
    d = torch.randn(n, 3, generator=g, device="cuda")
    rays[:, 3:6] = d / d.norm(dim=1, keepdim=True)
    rays[:, 6] = torch.rand(n, generator=g, device="cuda") * 150 + 1
    rays[: n // 50, 3] = 0.0
    torch.cuda.synchronize()  # the library reads the rays on its own stream
    return rays


@pytest.mark.parametrize("config", [1, 2, 4])
def test_device_built_tree_matches_unmodified_kernel(trt, ref, ctx, scenes, config):
    """Primary-ray ids and d_min over the device-built tree == the unmodified reference kernel."""
    sc = scenes.get(config)
    ctx.upload(sc, builder=trt.BUILD_DEVICE_LBVH)
    info = ctx.scene_info()
    assert info["builder"] == trt.BUILD_DEVICE_LBVH and info["n_wide_leaf_tris"] + info["n_top_prims"] == len(sc.objects)
    ref.init_scene(sc)
    w, h = CASES[config]
    cam, w, h = trt.config_camera(config, w, h)
    n = w * h
    for frame in (1, 2):
        want = ref.first_hit_ids(w, h, frame, cam)
        rt = dev_zeros(n, torch.float32)
        ref.primary_counts(w, h, frame, cam, None, rt)
        mid, mt = dev_zeros(n, torch.int32), dev_zeros(n, torch.float32)
        ctx.trace_primary(w, h, frame, cam, trt.TRAVERSE_FAST, d_id=mid, d_t=mt)
        assert int((mid.cpu().numpy() != want).sum()) == 0, "first-hit ids differ"
        assert int((mt.view(torch.int32) != rt.view(torch.int32)).sum()) == 0, "d_min differs"
    print(f"config {config}: device build {info['build_ms']:.2f} ms, {info['n_wide_nodes']} wide nodes, depth {info['wide_depth']}")


@pytest.mark.parametrize("config", [1, 2])
def test_device_built_tree_secondary_rays(trt, ctx, scenes, config):
    """Incoherent closest-hit and shadow rays: device-built tree (FAST) == reference order (REF)."""
    sc = scenes.get(config)
    ctx.upload(sc, builder=trt.BUILD_DEVICE_LBVH)
    n = 300_000
    rays = random_rays(n, 23 + config)
    out = {}
    for mode in (trt.TRAVERSE_REF, trt.TRAVERSE_FAST):
        i, t, o = dev_zeros(n, torch.int32), dev_zeros(n, torch.float32), dev_zeros(n, torch.int32)
        ctx.trace_closest(rays, n, mode, i, t)
        ctx.trace_shadow(rays, n, mode, o)
        out[mode] = (i, t, o)
    a, b = out[trt.TRAVERSE_REF], out[trt.TRAVERSE_FAST]
    assert torch.equal(a[0], b[0]) and torch.equal(a[1].view(torch.int32), b[1].view(torch.int32))
    assert torch.equal(a[2], b[2])


def test_standalone_device_build_without_reference_nodes(trt, scenes):
    """nodes=None: the scene is built from the object array alone.  Results equal the
    reference-order traversal of a context that has the node array, except for the rays whose
    outcome in the reference depends on its visit order (replay cases)."""
    sc = scenes.get(2)
    full, alone = trt.Context(0), trt.Context(0)
    try:
        full.upload(sc)
        alone.init_scene_data(sc.objects, sc.texture_files, None, sc.lights, builder=trt.BUILD_DEVICE_LBVH)
        assert alone.scene_info()["n_ref_nodes"] == 0
        n = 300_000
        rays = random_rays(n, 5)
        i0, t0, o0 = dev_zeros(n, torch.int32), dev_zeros(n, torch.float32), dev_zeros(n, torch.int32)
        i1, t1, o1 = dev_zeros(n, torch.int32), dev_zeros(n, torch.float32), dev_zeros(n, torch.int32)
        full.trace_closest(rays, n, trt.TRAVERSE_REF, i0, t0)
        full.trace_shadow(rays, n, trt.TRAVERSE_REF, o0)
        alone.trace_closest(rays, n, trt.TRAVERSE_FAST, i1, t1)
        alone.trace_shadow(rays, n, trt.TRAVERSE_FAST, o1)
        diff = int((i0 != i1).sum())
        assert diff <= n * 1e-4, f"{diff} of {n} closest hits differ"
        assert torch.equal(o0, o1)
        with pytest.raises(trt.TrtError):
            alone.trace_closest(rays, n, trt.TRAVERSE_REF, i1, t1)
        # and it renders
        cam, w, h = trt.config_camera(2, 320, 180)
        acc = dev_zeros(w * h * 4, torch.float32)
        alone.render(acc, w, h, 1, 2, cam)
        alone.synchronize()
        assert float(acc.sum()) > 0
    finally:
        full.close()
        alone.close()


def test_c5_small_field_device_vs_host_builder(trt, ref, scenes):
    """C5 at a 13 x 13 teapot grid (1.07 M triangles): the device builder is used automatically,
    and primary ids over its tree equal the unmodified reference kernel's."""
    sc = scenes.get(5, grid=13)
    c = trt.Context(0)
    try:
        c.upload(sc)  # AUTO -> device LBVH above 256 Ki objects
        info = c.scene_info()
        assert info["builder"] == trt.BUILD_DEVICE_LBVH
        ref.init_scene(sc)
        cam, w, h = trt.config_camera(5, 960, 540)
        n = w * h
        want = ref.first_hit_ids(w, h, 1, cam)
        mid, amb = dev_zeros(n, torch.int32), dev_zeros(n, torch.int32)
        c.trace_primary(w, h, 1, cam, trt.TRAVERSE_FAST, d_id=mid, d_entered=amb)
        bad = int((mid.cpu().numpy() != want).sum())
        print(f"C5 13x13: device build {info['build_ms']:.1f} ms, {info['n_wide_nodes']} wide nodes, depth "
              f"{info['wide_depth']}, hit fraction {(want >= 0).mean():.3f}, replays {int(amb.sum())}")
        assert bad == 0
    finally:
        c.close()


def test_c5_field_incoherent_rays_device_tree_equals_reference_order(trt, scenes):
    """C5 at a 6 x 6 grid (228 k triangles), rays scattered through the field: closest hits and shadow
    bits over the device-built tree (hybrid LBVH / SAH top levels) equal the reference-order traversal."""
    sc = scenes.get(5, grid=6)
    c = trt.Context(0)
    try:
        c.upload(sc, builder=trt.BUILD_DEVICE_LBVH)
        n = 500_000
        rays = random_rays(n, 77, lo=(-200, 0.05, -20), hi=(-100, 12, 80))
        out = {}
        for mode in (trt.TRAVERSE_REF, trt.TRAVERSE_FAST):
            i, t, o = dev_zeros(n, torch.int32), dev_zeros(n, torch.float32), dev_zeros(n, torch.int32)
            c.trace_closest(rays, n, mode, i, t)
            c.trace_shadow(rays, n, mode, o)
            out[mode] = (i, t, o)
        a, b = out[trt.TRAVERSE_REF], out[trt.TRAVERSE_FAST]
        hit_mesh = float((a[0] >= 2).float().mean())  # ids 0/1 are the floor and the light in the sorted array or not; any id counts
        assert torch.equal(a[0], b[0]) and torch.equal(a[1].view(torch.int32), b[1].view(torch.int32))
        assert torch.equal(a[2], b[2])
        assert float((a[0] >= 0).float().mean()) > 0.3 and hit_mesh > 0.05
    finally:
        c.close()


def c5_parts(trt, assets, grid):
    """The C5 field as (extra objects, unit mesh, placements): what create_config_scene(5) instances on the host."""
    objs, _ = trt.create_scene(5, assets, grid=grid)
    unit = trt.load_obj(str(assets / "teapot.obj"))  # offset 0, scale 1: the file's own vertices
    extra = objs[:2]                                  # floor + light
    inst = np.array([(-175.0 + 9.0 * ix, 0.0, 60.0 - 9.0 * iz, 1.2) for iz in range(grid) for ix in range(grid)],
                    dtype=np.float32)
    return objs, extra, unit, inst


def test_instanced_upload_builds_the_loaders_object_array_on_the_device(trt, assets, ctx):
    """f3: the mesh is parsed once and placed by a kernel; the object array the device ends up with is,
    byte for byte, the one the loader builds by re-reading the file per instance (reference
    src/loader.cpp:22-103, one load_obj call per placement), and renders to the same first-hit ids."""
    grid = 3
    objs, extra, unit, inst = c5_parts(trt, assets, grid)
    lights = trt.collect_lights(objs)
    ctx.upload_instanced(extra, unit, inst, lights)
    got = ctx.get_objects()
    assert got.tobytes() == np.ascontiguousarray(objs).tobytes()
    # the reference loader itself, per placement (offset, scale as load_obj arguments)
    one = trt.load_obj(str(assets / "teapot.obj"), offset=tuple(float(v) for v in inst[4][:3]), scale=float(inst[4][3]))
    n = len(unit)
    assert got[2 + 4 * n: 2 + 5 * n].tobytes() == np.ascontiguousarray(one).tobytes()
    cam, w, h = trt.config_camera(5, 480, 270)
    ids_a = dev_zeros(w * h, torch.int32)
    ctx.trace_primary(w, h, 1, cam, trt.TRAVERSE_FAST, d_id=ids_a)
    ctx.init_scene_data(objs, [], None, lights, builder=trt.BUILD_DEVICE_LBVH)
    ids_b = dev_zeros(w * h, torch.int32)
    ctx.trace_primary(w, h, 1, cam, trt.TRAVERSE_FAST, d_id=ids_b)
    assert torch.equal(ids_a, ids_b) and int((ids_a >= 0).sum()) > 1000  # the same scene, and it is in view
