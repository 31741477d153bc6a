"""GPU tier: material coverage beyond the diffuse benchmark scenes and the
depth-limited variant of BASELINE config C2 ("4-bounce").

  * stock Cornell scene of the reference (create_cornell_box, reference src/scene.cpp:24-123):
    mirror triangle, metallic rough teapot (always the SPEC lobe), textured back wall;
  * a glass teapot (transmission 0.9, ior 1.5, roughness 0.05): REFR lobe with total internal
    reflection and rough transmission (reference src/renderer.cu:592-648), thin-lens camera;
  * max_depth = 4: oracle = the instrumented restatement (oracle/ref_gpu.cu), itself checked here
    against the unmodified kernel at the reference's own depth.
Bar: PSNR >= 40 dB on the tone-mapped image with identical RNG streams, image means within 0.5 %.
"""
import numpy as np
import pytest
import torch

from gpu_common import SceneCache, dev_zeros, psnr_8bit

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scenes(trt, assets):
    return SceneCache(trt, assets)


@pytest.fixture(scope="module")
def ctx(trt):
    c = trt.Context(0)
    yield c
    c.close()


def gate(trt, ref, ctx, sc, cam, w, h, spp, what, **opt):
    ctx.upload(sc)
    ref.init_scene(sc)
    acc_ref, stage = dev_zeros(w * h * 4, torch.float32), dev_zeros(w * h * 4, torch.float32)
    ref.render_frames(acc_ref, stage, w, h, 1, spp, cam, cadence=0)
    acc = dev_zeros(w * h * 4, torch.float32)
    ctx.render(acc, w, h, 1, spp, cam, trt.default_opts(pool_paths=1 << 18, **opt))
    ctx.synchronize()
    a, b = acc.cpu().numpy().reshape(-1, 4)[:, :3], acc_ref.cpu().numpy().reshape(-1, 4)[:, :3]
    assert np.isfinite(a).all()
    p = psnr_8bit(ref.tonemap(acc.cpu().numpy(), spp), ref.tonemap(acc_ref.cpu().numpy(), spp))
    rel = np.abs(a.mean(0) - b.mean(0)) / np.maximum(b.mean(0), 1e-6)
    print(f"{what}: PSNR {p:.2f} dB, mean rel err {rel}")
    assert p >= 40.0 and (rel < 5e-3).all()


def test_stock_cornell_scene_specular_and_texture(trt, ref, ctx, scenes):
    sc = scenes.get(0)
    cam, w, h = trt.config_camera(0, 600, 400)
    gate(trt, ref, ctx, sc, cam, w, h, 16, "stock cornell (mirror, metallic teapot, texture)")


def test_glass_teapot_refraction_and_thin_lens(trt, ref, ctx, assets):
    objs, tex = trt.create_scene(2, assets)
    mesh = np.arange(len(objs)) >= 7  # the room shell comes first (reference src/scene.cpp:59-91)
    objs["transmission"][mesh] = 0.9
    objs["ior"][mesh] = 1.5
    objs["roughness"][mesh] = 0.05
    sc = trt.HostScene(objs, tex)
    cam = trt.CameraController((50, 45, 230), yaw=-90.0, pitch=-6.0, aperture=2.0, focus_dist=150.0).get_params(640, 360)
    gate(trt, ref, ctx, sc, cam, 640, 360, 16, "glass teapot, aperture 2")


def test_depth_limited_render_matches_restatement(trt, ref, ctx, scenes):
    """MAX_DEPTH = 4 ("4-bounce" C2).  The restatement equals the unmodified kernel at depth 30, and
    this library equals the restatement at depth 4."""
    sc = scenes.get(2)
    ctx.upload(sc)
    ref.init_scene(sc)
    cam, w, h = trt.config_camera(2, 640, 360)
    spp = 8
    unmod, stage = dev_zeros(w * h * 4, torch.float32), dev_zeros(w * h * 4, torch.float32)
    ref.render_frames(unmod, stage, w, h, 1, spp, cam, cadence=0)
    rest30 = dev_zeros(w * h * 4, torch.float32)
    ref.full_counts(rest30, w, h, 1, spp, cam, max_depth=30)
    p30 = psnr_8bit(ref.tonemap(rest30.cpu().numpy(), spp), ref.tonemap(unmod.cpu().numpy(), spp))
    assert p30 >= 40.0, f"restatement vs unmodified kernel: {p30:.1f} dB"
    rest4 = dev_zeros(w * h * 4, torch.float32)
    c4 = ref.full_counts(rest4, w, h, 1, spp, cam, max_depth=4)
    acc = dev_zeros(w * h * 4, torch.float32)
    ctx.reset_counters()
    ctx.render(acc, w, h, 1, spp, cam, trt.default_opts(pool_paths=1 << 18, max_depth=4, count_rays=1))
    got = ctx.counters()
    p4 = psnr_8bit(ref.tonemap(acc.cpu().numpy(), spp), ref.tonemap(rest4.cpu().numpy(), spp))
    print(f"depth 30 restatement vs unmodified {p30:.1f} dB; depth 4 ours vs restatement {p4:.1f} dB; "
          f"rays ours {got['closest_rays']}+{got['shadow_rays']} oracle {c4['closest_rays']}+{c4['shadow_rays']}")
    assert p4 >= 40.0
    assert abs(got["closest_rays"] - c4["closest_rays"]) <= 1e-4 * c4["closest_rays"]
    assert abs(got["shadow_rays"] - c4["shadow_rays"]) <= 1e-4 * c4["shadow_rays"]
    assert float(acc.sum()) < float(unmod.sum())  # truncating paths removes energy
