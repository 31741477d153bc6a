"""GPU tier: parity of the CUDA path with the reference, through the C ABI.

Oracles (oracle/_ref/libtrt_ref.so, built from the unmodified reference sources):
  * ref.first_hit_ids   -- the UNMODIFIED kernel, first-hit ids via the ID-as-emission scene;
  * ref.primary_counts  -- instrumented restatement of reference renderer.cu:319-425 (rays, d_min,
                           visit counters), itself checked here against the unmodified kernel's ids;
  * ref.launch          -- the unmodified kernel for radiance.
Bars: bit-exact for rays, ids, d_min and counters; PSNR >= 40 dB on the tone-mapped image for
radiance with identical RNG streams (SURVEY 8d gate (i)).
"""
import numpy as np
import pytest
import torch

from gpu_common import SceneCache, dev_zeros, psnr_8bit

pytestmark = pytest.mark.gpu

# resolution-reduced variants keep the oracle runs short; full sizes are covered by the bench
CASES = {1: (640, 480), 2: (960, 540), 3: (960, 540), 4: (960, 540)}


@pytest.fixture(scope="module")
def scenes(trt, assets):
    return SceneCache(trt, assets)


@pytest.fixture(scope="module")
def ctx(trt):
    c = trt.Context(0)
    yield c
    c.close()


def setup(trt, ref, ctx, scenes, config):
    sc = scenes.get(config)
    ctx.upload(sc)
    ref.init_scene(sc)
    w, h = CASES[config]
    cam, w, h = trt.config_camera(config, w, h)
    return sc, cam, w, h


def test_rng_states_match_curand(trt, ref, ctx):
    """Device skip-ahead == curand_init(1984+frame, pixel, 0) (reference renderer.cu:326)."""
    w, h = 1920, 1080
    rs = np.random.RandomState(7)
    for frame in (1, 2, 64, 1000):
        st = dev_zeros(w * h * 6, torch.int32)
        ctx.rng_states(w, h, frame, 0, w * h, st)
        mine = st.cpu().numpy().view(np.uint32).reshape(-1, 6)
        pix = np.concatenate([[0, 1, w - 1, w, w * h - 1], rs.randint(0, w * h, 200)])
        want = ref.xorwow_states(1984 + frame, pix)
        assert (mine[pix] == want).all()


@pytest.mark.parametrize("config", [1, 2, 3, 4])
def test_reference_order_primary_parity(trt, ref, ctx, scenes, config):
    """TRAVERSE_REF: primary rays, ids, d_min and the three visit counters are bit-equal."""
    sc, cam, w, h = setup(trt, ref, ctx, scenes, config)
    n = w * h
    for frame in (1, 2, 3, 4):
        ids_unmodified = ref.first_hit_ids(w, h, frame, cam)
        r = {k: dev_zeros(n, torch.int32) for k in ("id", "f", "e", "t3")}
        rt, rray = dev_zeros(n, torch.float32), dev_zeros(n * 6, torch.float32)
        ref.primary_counts(w, h, frame, cam, r["id"], rt, rray, r["f"], r["e"], r["t3"])
        m = {k: dev_zeros(n, torch.int32) for k in ("id", "f", "e", "t3")}
        mt, mray = dev_zeros(n, torch.float32), dev_zeros(n * 6, torch.float32)
        ctx.trace_primary(w, h, frame, cam, trt.TRAVERSE_REF, d_id=m["id"], d_t=mt, d_ray=mray,
                          d_fetched=m["f"], d_entered=m["e"], d_tris=m["t3"])
        diag = dict(
            oracle_vs_unmodified=int((r["id"].cpu().numpy() != ids_unmodified).sum()),
            rays=int((mray.view(torch.int32) != rray.view(torch.int32)).sum()),
            ids=int((m["id"].cpu().numpy() != ids_unmodified).sum()),
            t=int((mt.view(torch.int32) != rt.view(torch.int32)).sum()),
            fetched=int((m["f"] != r["f"]).sum()), entered=int((m["e"] != r["e"]).sum()),
            tris=int((m["t3"] != r["t3"]).sum()))
        print(f"config {config} frame {frame} mismatches: {diag}")
        # the instrumented oracle reproduces the unmodified kernel
        assert diag["oracle_vs_unmodified"] == 0
        # ours against both
        assert diag["rays"] == 0, "primary rays differ"
        assert diag["ids"] == 0, "first-hit ids differ"
        assert diag["t"] == 0, "d_min differs"
        assert diag["fetched"] == 0 and diag["entered"] == 0 and diag["tris"] == 0, "visit counters differ"


@pytest.mark.parametrize("config", [1, 2, 3, 4])
def test_fast_primary_ids_match_unmodified_kernel(trt, ref, ctx, scenes, config):
    """TRAVERSE_FAST (wide BVH + exact accept + replay): ids and d_min equal the reference's on
    100% of pixels for frame seeds 1..4."""
    sc, cam, w, h = setup(trt, ref, ctx, scenes, config)
    n = w * h
    replays = 0
    for frame in (1, 2, 3, 4):
        ids_unmodified = ref.first_hit_ids(w, h, frame, cam)
        rt = dev_zeros(n, torch.float32)
        ref.primary_counts(w, h, frame, cam, None, rt)
        mid, mt, amb = dev_zeros(n, torch.int32), dev_zeros(n, torch.float32), dev_zeros(n, torch.int32)
        ctx.trace_primary(w, h, frame, cam, trt.TRAVERSE_FAST, d_id=mid, d_t=mt, d_entered=amb)
        bad_id = int((mid.cpu().numpy() != ids_unmodified).sum())
        bad_t = int((mt.view(torch.int32) != rt.view(torch.int32)).sum())
        print(f"config {config} frame {frame}: id mismatches {bad_id}, d_min mismatches {bad_t}, replays {int(amb.sum())}")
        assert bad_id == 0, "first-hit ids differ"
        assert bad_t == 0, "d_min differs"
        replays += int(amb.sum())
    assert replays < 4 * n * 1e-3, f"replay rate too high: {replays}"


@pytest.mark.parametrize("config", [1, 2])
def test_secondary_rays_fast_equals_reference_order(trt, ref, ctx, scenes, config):
    """Incoherent rays: FAST and REF agree with each other on closest hit (id, t) and on shadow-ray occlusion
    for every ray (both are checked against the oracle in test_arbitrary_rays_match_the_oracle)."""
    sc, cam, w, h = setup(trt, ref, ctx, scenes, config)
    g = torch.Generator(device="cuda").manual_seed(11 + config)
    n = 400_000
    rays = torch.zeros(n, 8, device="cuda")
    rays[:, 0] = torch.rand(n, generator=g, device="cuda") * 100
    rays[:, 1] = torch.rand(n, generator=g, device="cuda") * 100
    rays[:, 2] = torch.rand(n, generator=g, device="cuda") * 300
    d = torch.randn(n, 3, generator=g, device="cuda")
    rays[:, 3:6] = d / d.norm(dim=1, keepdim=True)
    rays[:, 6] = torch.rand(n, generator=g, device="cuda") * 150 + 1
    rays[: n // 50, 3] = 0.0  # axis-parallel directions exercise the safe_inv / infinite-reciprocal paths
    torch.cuda.synchronize()  # the library reads the rays on its own stream
    out = {}
    for mode in (trt.TRAVERSE_REF, trt.TRAVERSE_FAST):
        i, t, o = dev_zeros(n, torch.int32), dev_zeros(n, torch.float32), dev_zeros(n, torch.int32)
        ctx.trace_closest(rays, n, mode, i, t)
        ctx.trace_shadow(rays, n, mode, o)
        out[mode] = (i, t, o)
    a, b = out[trt.TRAVERSE_REF], out[trt.TRAVERSE_FAST]
    assert torch.equal(a[0], b[0]) and torch.equal(a[1].view(torch.int32), b[1].view(torch.int32))
    assert torch.equal(a[2], b[2])
    assert 0.05 < (a[0] >= 0).float().mean() <= 1.0 and 0.0 < a[2].float().mean() < 1.0


def _incoherent_rays(n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    rays = torch.zeros(n, 8, device="cuda")
    rays[:, 0] = torch.rand(n, generator=g, device="cuda") * 100
    rays[:, 1] = torch.rand(n, generator=g, device="cuda") * 100
    rays[:, 2] = torch.rand(n, generator=g, device="cuda") * 300
    d = torch.randn(n, 3, generator=g, device="cuda")
    rays[:, 3:6] = d / d.norm(dim=1, keepdim=True)
    rays[:, 6] = torch.rand(n, generator=g, device="cuda") * 150 + 1
    rays[: n // 50, 3] = 0.0  # axis-parallel directions exercise the safe_inv / infinite-reciprocal paths
    rays[n // 50: n // 25, 4] = 0.0
    torch.cuda.synchronize()
    return rays


@pytest.mark.parametrize("config", [1, 2, 3, 4])
def test_arbitrary_rays_match_the_oracle(trt, ref, ctx, scenes, config):
    """Incoherent closest-hit AND shadow rays, both traversal modes, against the ORACLE (not against each other):
    closest hit = restatement of reference renderer.cu:371-425 (ids checked against the unmodified kernel above),
    shadow = the reference's own unmodified __device__ trace_shadow (renderer.cu:273-314).  Bit-exact."""
    sc, cam, w, h = setup(trt, ref, ctx, scenes, config)
    n = 400_000
    rays = _incoherent_rays(n, 23 + config)
    want_id, want_t, want_occ = dev_zeros(n, torch.int32), dev_zeros(n, torch.float32), dev_zeros(n, torch.int32)
    ref.trace_rays(rays, n, 0, d_id=want_id, d_t=want_t)
    ref.trace_rays(rays, n, 1, d_occ=want_occ)
    assert 0.05 < (want_id >= 0).float().mean() <= 1.0 and 0.0 < want_occ.float().mean() < 1.0
    for mode in (trt.TRAVERSE_REF, trt.TRAVERSE_FAST):
        i, t, o = dev_zeros(n, torch.int32), dev_zeros(n, torch.float32), dev_zeros(n, torch.int32)
        ctx.trace_closest(rays, n, mode, i, t)
        ctx.trace_shadow(rays, n, mode, o)
        bad_id = int((i != want_id).sum())
        hit = want_id >= 0
        bad_t = int((t.view(torch.int32)[hit] != want_t.view(torch.int32)[hit]).sum())
        bad_occ = int((o != want_occ).sum())
        print(f"config {config} mode {mode}: id mismatches {bad_id}, d_min mismatches {bad_t}, occlusion mismatches {bad_occ}")
        assert bad_id == 0, "closest-hit ids differ from the oracle"
        assert bad_t == 0, "closest-hit distances differ from the oracle"
        assert bad_occ == 0, "shadow-ray occlusion differs from the reference's trace_shadow"


@pytest.mark.parametrize("config", [3, 4])
def test_fast_primary_ids_at_the_named_resolution(trt, ref, ctx, scenes, config):
    """Frame 1 of C3 (1920x1080) and C4 (3840x2160) at the resolution BASELINE.json names: first-hit ids of the
    fast path equal the unmodified kernel's on every pixel."""
    sc = scenes.get(config)
    ctx.upload(sc)
    ref.init_scene(sc)
    cam, w, h = trt.config_camera(config)
    n = w * h
    want = ref.first_hit_ids(w, h, 1, cam)
    ids = dev_zeros(n, torch.int32)
    ctx.trace_primary(w, h, 1, cam, trt.TRAVERSE_FAST, d_id=ids)
    bad = int((ids.cpu().numpy() != want).sum())
    print(f"config {config} {w}x{h}: id mismatches {bad} of {n}")
    assert bad == 0


@pytest.mark.parametrize("config,mode", [(1, "ref"), (1, "fast"), (2, "fast"), (3, "fast")])
def test_radiance_same_stream_gate(trt, ref, ctx, scenes, config, mode):
    """Same seeds, same spp, against the unmodified reference kernel (SURVEY 8d gate (i), tightened to what is
    measured): PSNR >= 80 dB on the tone-mapped 8-bit image (SURVEY asks for 40), at least 99.9 % of the linear
    radiance values within 1e-4 (relative, floor 1), image means within 0.1 %.  Differences can only come from
    FP re-association in shading, the summation order of the accumulate, and the rare decision flips they cause."""
    sc, cam, w, h = setup(trt, ref, ctx, scenes, config)
    spp = 16
    acc_ref, stage = dev_zeros(w * h * 4, torch.float32), dev_zeros(w * h * 4, torch.float32)
    ref.render_frames(acc_ref, stage, w, h, 1, spp, cam, cadence=0)
    acc = dev_zeros(w * h * 4, torch.float32)
    opts = trt.default_opts(traversal=trt.TRAVERSE_REF if mode == "ref" else trt.TRAVERSE_FAST, pool_paths=1 << 18)
    ctx.render(acc, w, h, 1, spp, cam, opts)
    ctx.synchronize()
    a, b = acc.cpu().numpy().reshape(-1, 4)[:, :3], acc_ref.cpu().numpy().reshape(-1, 4)[:, :3]
    assert np.isfinite(a).all()
    img_a = ref.tonemap(acc.cpu().numpy(), spp)
    img_b = ref.tonemap(acc_ref.cpu().numpy(), spp)
    p = psnr_8bit(img_a, img_b)
    rel = np.abs(a.mean(0) - b.mean(0)) / np.maximum(b.mean(0), 1e-6)
    frac_equal = np.mean(np.abs(a - b) <= 1e-4 * np.maximum(np.abs(b), 1.0))
    print(f"config {config} {mode}: PSNR {p:.2f} dB, mean rel err {rel}, values within 1e-4: {frac_equal:.6f}")
    assert p >= 80.0
    assert frac_equal >= 0.999
    assert (rel < 1e-3).all()


@pytest.mark.parametrize("config,w,h,n_spp,ref_spp", [(1, 320, 240, 64, 4096), (2, 480, 270, 64, 2048)])
def test_radiance_estimator_gate(trt, ref, ctx, scenes, config, w, h, n_spp, ref_spp):
    """SURVEY 8d gate (ii), valid whatever the RNG streams: with R_inf = the unmodified kernel at ref_spp (disjoint
    seeds), MSE(ours@N, R_inf) <= 1.1 x MSE(reference@N with other seeds, R_inf) on linear radiance, and the
    per-channel mean of the whole image within 0.5 % of R_inf's (a mis-reproduced estimator quirk shows up as bias).
    The MSE of a 64-spp path-traced image is itself a noisy, firefly-dominated statistic (two reference runs with
    different seeds differ by tens of percent), so radiance is clamped at 4 for the MSE and the reference side is the
    largest of three independent runs."""
    sc = scenes.get(config)
    ctx.upload(sc)
    ref.init_scene(sc)
    cam, w, h = trt.config_camera(config, w, h)
    n = w * h
    stage = dev_zeros(n * 4, torch.float32)

    def reference(first, spp):
        a = dev_zeros(n * 4, torch.float32)
        ref.render_frames(a, stage, w, h, first, spp, cam, cadence=0)
        return a.cpu().numpy().reshape(-1, 4)[:, :3].astype(np.float64) / spp

    r_inf = reference(100_001, ref_spp)
    ours = dev_zeros(n * 4, torch.float32)
    ctx.render(ours, w, h, 1, n_spp, cam, trt.default_opts(pool_paths=1 << 18))
    ctx.synchronize()
    ours = ours.cpu().numpy().reshape(-1, 4)[:, :3].astype(np.float64) / n_spp

    def mse(a):
        return float(np.mean((np.minimum(a, 4.0) - np.minimum(r_inf, 4.0)) ** 2))

    mse_ours = mse(ours)
    mse_refs = [mse(reference(1 + k * n_spp, n_spp)) for k in (1, 2, 3)]
    bias = np.abs(ours.mean(0) - r_inf.mean(0)) / np.maximum(r_inf.mean(0), 1e-9)
    print(f"config {config}: MSE ours {mse_ours:.6g} reference runs {[f'{m:.6g}' for m in mse_refs]} "
          f"ratio to the largest {mse_ours / max(mse_refs):.4f}, mean rel err {bias}")
    assert mse_ours <= 1.1 * max(mse_refs)
    assert (bias <= 5e-3).all()


def test_render_matches_instrumented_ray_counts(trt, ref, ctx, scenes):
    """The wavefront renderer traces the same rays as the reference loop: ray counts equal those of
    the instrumented restatement to within the rare decision flips."""
    sc, cam, w, h = setup(trt, ref, ctx, scenes, 1)
    want = ref.full_counts(None, w, h, 1, 4, cam)
    acc = dev_zeros(w * h * 4, torch.float32)
    ctx.reset_counters()
    ctx.render(acc, w, h, 1, 4, cam, trt.default_opts(count_rays=1, pool_paths=1 << 18))
    got = ctx.counters()
    assert got["samples"] == 4 * w * h
    assert abs(got["closest_rays"] - want["closest_rays"]) <= 1e-4 * want["closest_rays"]
    assert abs(got["shadow_rays"] - want["shadow_rays"]) <= 1e-4 * want["shadow_rays"]


def test_frame_sharding_equals_single_pass(trt, ctx, scenes):
    """Sample sharding (frame_stride): two interleaved halves sum to the full render."""
    sc = scenes.get(1)
    ctx.upload(sc)
    cam, w, h = trt.config_camera(1, 320, 240)
    o = trt.default_opts(pool_paths=1 << 16)
    full = dev_zeros(w * h * 4, torch.float32)
    ctx.render(full, w, h, 1, 8, cam, o)
    parts = dev_zeros(w * h * 4, torch.float32)
    ctx.render(parts, w, h, 1, 4, cam, o, frame_stride=2)
    ctx.render(parts, w, h, 2, 4, cam, o, frame_stride=2)
    ctx.synchronize()
    a, b = full.cpu().numpy(), parts.cpu().numpy()
    assert np.allclose(a, b, rtol=1e-5, atol=1e-5)


def test_render_to_host_and_tonemap(trt, ref, ctx, scenes):
    sc = scenes.get(1)
    ctx.upload(sc)
    cam, w, h = trt.config_camera(1, 320, 240)
    host = torch.zeros(w * h * 4, dtype=torch.float32).pin_memory()
    ctx.render_to_host(host, w, h, 1, 4, cam, trt.default_opts(pool_paths=1 << 16))
    dev = dev_zeros(w * h * 4, torch.float32)
    ctx.render(dev, w, h, 1, 4, cam, trt.default_opts(pool_paths=1 << 16))
    ctx.synchronize()
    assert np.allclose(host.numpy(), dev.cpu().numpy(), rtol=1e-5, atol=1e-5)
    assert host.numpy().reshape(-1, 4)[:, :3].sum() > 0
    argb = dev_zeros(w * h, torch.int32)
    ctx.tonemap(dev, w, h, 4, argb)
    ctx.synchronize()
    want = ref.tonemap(dev.cpu().numpy(), 4)
    got = argb.cpu().numpy().view(np.uint32)
    assert (got != want).mean() < 1e-5  # double pow on device vs glibc: identical up to rare last-bit ties


def test_call_order_and_argument_errors(trt):
    c = trt.Context(0)
    cam, w, h = trt.config_camera(1, 64, 48)
    acc = dev_zeros(w * h * 4, torch.float32)
    with pytest.raises(trt.TrtError, match="before trt_upload_scene"):
        c.render(acc, w, h, 1, 1, cam)
    with pytest.raises(trt.TrtError):
        c.init_scene_data(np.zeros(0, dtype=trt.OBJECT), [], np.zeros(0, dtype=trt.NODE), [])
    objs = np.zeros(1, dtype=trt.OBJECT)
    objs["v1"]["x"] = 1
    objs["v2"]["y"] = 1
    objs["tex_id"] = -1
    nodes = trt.BVH()
    objs = nodes.build(objs)
    with pytest.raises(trt.TrtError, match="light index"):
        c.init_scene_data(objs, [], nodes.get_nodes(), [5])
    c.init_scene_data(objs, [], nodes.get_nodes(), [])
    with pytest.raises(trt.TrtError, match="max_depth"):
        c.render(acc, w, h, 1, 1, cam, trt.default_opts(max_depth=0))
    c.render(acc, w, h, 1, 1, cam)  # a one-triangle scene with no lights renders (black)
    c.synchronize()
    assert float(acc.abs().sum()) == 0.0
    c.close()
