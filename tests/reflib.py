"""ctypes bindings of the test oracles (test infrastructure only).

oracle/_ref/libtrt_ref.so  -- the unmodified reference sources + driver (oracle/ref_gpu.cu)
oracle/_build/liboracle_cpu.so -- CPU restatement (oracle/cpu_oracle.cpp)
"""
import ctypes as C
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF_PATH = ROOT / "oracle" / "_ref" / "libtrt_ref.so"
CPU_PATH = ROOT / "oracle" / "_build" / "liboracle_cpu.so"

_ref = None
_cpu = None


def available():
    return REF_PATH.exists()


def cpu_available():
    return CPU_PATH.exists()


def ref():
    global _ref
    if _ref is None:
        _ref = C.CDLL(str(REF_PATH))
    return _ref


def cpu():
    global _cpu
    if _cpu is None:
        _cpu = C.CDLL(str(CPU_PATH))
    return _cpu


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _dp(x):
    if x is None:
        return None
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(int(x))


def _f3(v):
    return (C.c_float * 3)(*[float(t) for t in v])


def load_obj(OBJECT, filename, offset, scale, albedo, metallic, roughness):
    L = ref()
    fn = str(filename).encode()
    n = L.ref_load_obj(fn, None, 0, _f3(offset), C.c_float(scale), _f3(albedo), C.c_float(metallic), C.c_float(roughness))
    out = np.zeros(n, dtype=OBJECT)
    if n:
        L.ref_load_obj(fn, _p(out), n, _f3(offset), C.c_float(scale), _f3(albedo), C.c_float(metallic), C.c_float(roughness))
    return out


def bvh_build(OBJECT, NODE, objects):
    objs = np.ascontiguousarray(objects, dtype=OBJECT).copy()
    nodes = np.zeros(2 * len(objs), dtype=NODE)
    n = ref().ref_bvh_build(_p(objs), len(objs), _p(nodes), len(nodes))
    assert n > 0
    return objs, nodes[:n].copy()


def camera_params(CAMERA, pos, mouse_dx, mouse_dy, w, h):
    cam = np.zeros(1, dtype=CAMERA)
    ref().ref_camera_params(_f3(pos), C.c_float(mouse_dx), C.c_float(mouse_dy), w, h, _p(cam))
    return cam


def xorwow_states(seed, pixels):
    px = np.ascontiguousarray(pixels, dtype=np.int32)
    out = np.zeros((len(px), 6), dtype=np.uint32)
    ref().ref_xorwow_states(C.c_ulonglong(seed), _p(px), len(px), _p(out))
    return out


def xorwow_draws(seed, pixel, n):
    u = np.zeros(n, dtype=np.uint32)
    f = np.zeros(n, dtype=np.float32)
    ref().ref_xorwow_draws(C.c_ulonglong(seed), int(pixel), n, _p(u), _p(f))
    return u, f


def tonemap(accum, frames):
    a = np.ascontiguousarray(accum, dtype=np.float32).reshape(-1, 4)
    out = np.zeros(len(a), dtype=np.uint32)
    ref().ref_tonemap(_p(a), len(a), frames, _p(out))
    return out


# ---- GPU side of the reference -----------------------------------------------------------
def init_scene(scene):
    tex = ";".join(scene.texture_files).encode()
    li = np.ascontiguousarray(scene.lights, dtype=np.int32)
    rc = ref().ref_init_scene(_p(scene.objects), len(scene.objects), _p(scene.nodes), len(scene.nodes), _p(li), len(li), tex)
    assert rc == 0, "reference init_scene_data failed"


def launch(d_accum, w, h, frame_seed, cam, tx=16, ty=16):
    ref().ref_launch(_dp(d_accum), w, h, frame_seed, tx, ty, _p(cam))


def render_frames(d_accum, d_staging, w, h, first, n, cam, cadence=1):
    ms = C.c_float()
    rc = ref().ref_render_frames(_dp(d_accum), _dp(d_staging), w, h, first, n, _p(cam), cadence, C.byref(ms))
    assert rc == 0
    return ms.value


def first_hit_ids(w, h, frame_seed, cam):
    ids = np.zeros(w * h, dtype=np.int32)
    rc = ref().ref_first_hit_ids(w, h, frame_seed, _p(cam), _p(ids))
    assert rc == 0, f"ID-as-emission decode failed on {rc} pixels"
    return ids


def primary_counts(w, h, frame_seed, cam, d_id=None, d_t=None, d_ray=None, d_fetched=None, d_entered=None, d_tris=None, seed_base=1984):
    rc = ref().ref_primary_counts(w, h, seed_base, frame_seed, _p(cam), _dp(d_id), _dp(d_t), _dp(d_ray), _dp(d_fetched), _dp(d_entered), _dp(d_tris))
    assert rc == 0


def trace_rays(d_rays, n, shadow, d_id=None, d_t=None, d_occ=None):
    """Arbitrary rays through the oracle: closest hit = restatement of reference renderer.cu:371-425,
    shadow = the reference's own unmodified trace_shadow (:273-314)."""
    rc = ref().ref_trace_rays(_dp(d_rays), int(n), int(shadow), _dp(d_id), _dp(d_t), _dp(d_occ))
    assert rc == 0, f"ref_trace_rays failed ({rc})"


def full_counts(d_accum, w, h, first, n, cam, max_depth=30, rr_threshold=3, seed_base=1984):
    tot = np.zeros(5, dtype=np.uint64)
    rc = ref().ref_full_counts(_dp(d_accum), w, h, seed_base, first, n, _p(cam), max_depth, rr_threshold, _p(tot))
    assert rc == 0
    return dict(closest_rays=int(tot[0]), shadow_rays=int(tot[1]), nodes_fetched=int(tot[2]), nodes_entered=int(tot[3]), tris_tested=int(tot[4]))


# ---- scenes built with the reference's own host code only (no product library involved) ------------
class ReferenceScene:
    """BASELINE config C2 / C4 assembled the way reference src/main.cpp:79-101 does, through the reference's
    own create_cornell_box (room shell), load_obj and BVH::build compiled into oracle/_ref: what the
    `--impl reference` arm of bench.py renders.  Byte-identical to tryraytrace_b200.HostScene.from_config
    (tests/test_host_surface.py checks that)."""

    MESH = {2: ("teapot.obj", (48.0, 5.0, 80.0), 14.0), 4: ("pumpkin.obj", (51.6, 29.5, 146.0), 0.6)}
    CAMERA = {2: ((50.0, 45.0, 230.0), 60.0, 1920, 1080), 4: ((50.0, 45.0, 230.0), 60.0, 3840, 2160)}

    def __init__(self, config, OBJECT, NODE, asset_dir=None):
        import os
        if config not in self.MESH:
            raise ValueError("reference scenes are built for C2 and C4")
        asset_dir = Path(asset_dir) if asset_dir else ROOT / "assets"
        L = ref()
        buf = np.zeros(8192, dtype=OBJECT)
        cwd = os.getcwd()
        os.chdir(asset_dir.parent)  # create_cornell_box reads "assets/teapot.obj" relative to the working directory
        try:
            n = L.ref_create_cornell(_p(buf), len(buf), None, 0)
        finally:
            os.chdir(cwd)
        assert n >= 7
        shell = buf[:7].copy()
        shell["tex_id"] = -1  # the back wall is untextured outside C3
        name, offset, scale = self.MESH[config]
        mesh = load_obj(OBJECT, asset_dir / name, offset, scale, (0.75, 0.75, 0.75), 0.0, 1.0)
        self.objects, self.nodes = bvh_build(OBJECT, NODE, np.concatenate([shell, mesh]))
        e = self.objects["emission"]
        self.lights = np.nonzero((e["x"] > 0.1) | (e["y"] > 0.1) | (e["z"] > 0.1))[0].astype(np.int32)  # src/main.cpp:88-96
        self.texture_files = []

    @classmethod
    def camera(cls, config, CAMERA, width=None, height=None):
        pos, pitch_units, w, h = cls.CAMERA[config]
        w, h = width or w, height or h
        return camera_params(CAMERA, pos, 0.0, pitch_units, w, h), w, h
