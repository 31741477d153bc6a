"""GPU tier: edge cases of the render entry points -- degenerate image sizes, empty jobs, scenes
without lights or without a tree, tiny pools, light lists longer than the root-level list takes.
Each case is checked against the unmodified reference kernel (ids exact, radiance PSNR >= 60 dB)."""
import numpy as np
import pytest
import torch

from gpu_common import dev_zeros, psnr_8bit

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(trt):
    c = trt.Context(0)
    yield c
    c.close()


def radiance_gate(trt, ref, ctx, sc, cam, w, h, spp, what, min_psnr=60.0, **opt):
    ctx.upload(sc)
    ref.init_scene(sc)
    n = w * h
    want = ref.first_hit_ids(w, h, 1, cam)
    ids = dev_zeros(n, torch.int32)
    ctx.trace_primary(w, h, 1, cam, trt.TRAVERSE_FAST, d_id=ids)
    assert (ids.cpu().numpy() == want).all(), f"{what}: first-hit ids differ"
    a_ref, stage = dev_zeros(n * 4, torch.float32), dev_zeros(n * 4, torch.float32)
    ref.render_frames(a_ref, stage, w, h, 1, spp, cam, cadence=0)
    acc = dev_zeros(n * 4, torch.float32)
    ctx.render(acc, w, h, 1, spp, cam, trt.default_opts(**opt))
    ctx.synchronize()
    a = acc.cpu().numpy()
    assert np.isfinite(a).all()
    p = psnr_8bit(ref.tonemap(a, spp), ref.tonemap(a_ref.cpu().numpy(), spp))
    print(f"{what}: PSNR {p:.1f} dB")
    assert p >= min_psnr, what
    return a


@pytest.mark.parametrize("w,h", [(1, 1), (3, 5), (33, 17), (257, 1)])
def test_degenerate_image_sizes(trt, ref, ctx, assets, w, h):
    sc = trt.HostScene.from_config(1, assets)
    cam, w, h = trt.config_camera(1, w, h)
    # a handful of pixels: one 8-bit level on one of them is already ~50 dB
    radiance_gate(trt, ref, ctx, sc, cam, w, h, 8, f"C1 {w}x{h}", min_psnr=40.0)


def test_zero_frames_is_a_no_op(trt, ctx, assets):
    sc = trt.HostScene.from_config(1, assets)
    ctx.upload(sc)
    cam, w, h = trt.config_camera(1, 64, 48)
    acc = dev_zeros(w * h * 4, torch.float32)
    ctx.render(acc, w, h, 1, 0, cam)
    ctx.synchronize()
    assert float(acc.abs().sum()) == 0.0


def test_tiny_pool_and_short_job(trt, ref, ctx, assets):
    """A pool far smaller than the image (many regeneration rounds) and a one-frame job."""
    sc = trt.HostScene.from_config(2, assets)
    cam, w, h = trt.config_camera(2, 320, 180)
    radiance_gate(trt, ref, ctx, sc, cam, w, h, 4, "C2 pool 256", pool_paths=256)
    radiance_gate(trt, ref, ctx, sc, cam, w, h, 1, "C2 one frame")


@pytest.mark.parametrize("pool", [33280, 40448, 65024])
def test_pool_between_the_compaction_floor_and_twice_the_floor(trt, ref, ctx, assets, pool):
    """Pools of 32 Ki .. 64 Ki slots on a tiny job: the drain compaction clamps its new bound to 32 Ki, so the list
    of dead slots below the bound can hold almost the whole pool (it used to be sized for half of it and
    overflowed into neighbouring device memory)."""
    sc = trt.HostScene.from_config(1, assets)
    cam, w, h = trt.config_camera(1, 64, 64)
    radiance_gate(trt, ref, ctx, sc, cam, w, h, 1, f"C1 64x64, pool {pool}", min_psnr=50.0, pool_paths=pool)
    cam, w, h = trt.config_camera(1, 320, 240)
    radiance_gate(trt, ref, ctx, sc, cam, w, h, 3, f"C1 320x240, pool {pool}", pool_paths=pool)


def test_scene_without_lights(trt, ref, ctx, assets):
    """No emitter: no next-event estimation, no shadow rays (reference :653 light_count == 0)."""
    objs, tex = trt.create_scene(2, assets)
    objs["emission"][:] = 0
    sc = trt.HostScene(objs, tex)
    assert len(sc.lights) == 0
    cam, w, h = trt.config_camera(2, 320, 180)
    ctx.upload(sc)
    acc = dev_zeros(w * h * 4, torch.float32)
    ctx.reset_counters()
    ctx.render(acc, w, h, 1, 2, cam)
    c = ctx.counters()
    assert c["shadow_rays"] == 0 and c["closest_rays"] > 0 and float(acc.abs().sum()) == 0.0


def test_many_lights_stay_in_the_tree(trt, ref, ctx, assets):
    """More emitters than the root-level list takes (kMaxTopLights = 4): they stay in the tree and the
    light list is sampled uniformly (reference :659)."""
    objs, tex = trt.create_scene(2, assets)
    mesh = np.where(np.arange(len(objs)) >= 7)[0]
    objs["emission"][mesh[::200]] = 8.0  # ~32 emissive teapot triangles
    sc = trt.HostScene(objs, tex)
    assert len(sc.lights) > 8
    cam, w, h = trt.config_camera(2, 320, 180)
    radiance_gate(trt, ref, ctx, sc, cam, w, h, 8, "C2 with 30+ lights")
    assert ctx.scene_info()["n_top_prims"] == 6


def test_scene_of_a_few_triangles_has_no_root_level_list(trt, ref, ctx, assets):
    """At most 16 primitives: nothing is lifted out of the tree (the whole scene is a handful of leaves)."""
    objs, tex = trt.create_scene(1, assets)
    sc = trt.HostScene(objs[:9].copy(), tex)  # room shell + two cube faces
    cam, w, h = trt.config_camera(1, 160, 120)
    radiance_gate(trt, ref, ctx, sc, cam, w, h, 4, "9-triangle scene")
    assert ctx.scene_info()["n_top_prims"] == 0


def test_open_scene_with_short_paths_drains_correctly(trt, ref, ctx, assets):
    """C5 at a 2 x 2 grid: most paths leave the scene after a bounce or two, so the job hands out its
    last samples while the pool is already thin -- the drain-phase compaction must not start in the
    iteration that still regenerates (regression: paths were moved into slots about to be refilled)."""
    sc = trt.HostScene.from_config(5, assets, grid=2)
    cam, w, h = trt.config_camera(5, 640, 360)
    for pool in (1 << 16, 1 << 18):
        a = radiance_gate(trt, ref, ctx, sc, cam, w, h, 6, f"C5 2x2 pool {pool}", pool_paths=pool)
        assert float(a.sum()) > 0


def test_render_longer_than_one_job(trt, ref, ctx, assets):
    """600 frames = three wavefront jobs (256 + 256 + 88 frames) on one stream pair: every job drains,
    compacts and hands the side stream back before the next one starts."""
    sc = trt.HostScene.from_config(1, assets)
    cam, w, h = trt.config_camera(1, 96, 64)
    radiance_gate(trt, ref, ctx, sc, cam, w, h, 600, "C1 96x64, 600 frames")


@pytest.mark.parametrize("config,w,h,spp,pool", [(1, 640, 480, 16, 0), (2, 960, 540, 8, 0), (2, 640, 360, 12, 40448)])
def test_slot_state_invariants_hold_under_debug_checks(trt, ref, ctx, assets, monkeypatch, config, w, h, spp, pool):
    """Race / ownership evidence without compute-sanitizer: with TRT_DEBUG_CHECKS=1 the kernels that overwrite slots
    they do not own by construction verify the slot state first -- refill: the slot ended in the last shade pass (dead,
    no shadow ray waiting); drain compaction: source live and beyond the new bound, destination dead and inside it.
    The counter must stay 0, the checked render must pass the same gates as the plain one, and the job must have
    gone through compactions and the drain-tail kernel (iterations well below the plain wavefront's)."""
    monkeypatch.setenv("TRT_DEBUG_CHECKS", "1")
    sc = trt.HostScene.from_config(config, assets)
    cam, _, _ = trt.config_camera(config, w, h)
    ctx.reset_counters()
    radiance_gate(trt, ref, ctx, sc, cam, w, h, spp, f"debug checks C{config} pool {pool}", pool_paths=pool)
    c = ctx.counters()
    assert c["samples"] == w * h * spp
    assert c["check_violations"] == 0
