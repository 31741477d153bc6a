"""tryraytrace_b200 -- Python harness over the C ABI of the B200 path-tracing core.

The product is the native library ``tryraytrace_b200/lib/libtrt_b200.so`` (C++ host
surface + hand-written sm_100a CUDA kernels, see ``include/trt_capi.h``).  This package
only binds it with ctypes for the tests and the benchmark; it holds no rendering code
and has no CPU fallback: importing works anywhere, but every compute call raises
``TrtError`` without a CUDA device, and loading fails loudly if the library is missing.

Names mirror the reference's host interface (``load_obj``, ``BVH.build``,
``create_cornell_box``, ``CameraController.get_params``, ``init_scene_data``,
``launch_render_kernel``; reference include/*.h).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
REPO_ROOT = _HERE.parent
LIB_PATH = _HERE / "lib" / "libtrt_b200.so"
ASSET_DIR = REPO_ROOT / "assets"

TRAVERSE_FAST = 0
TRAVERSE_REF = 1

# ---- record layouts: tryraytrace_b200/records.py (importable without the library) --------
from .records import VEC, OBJECT, NODE, CAMERA  # noqa: E402,F401


BUILD_AUTO, BUILD_HOST_SAH, BUILD_DEVICE_LBVH = 0, 1, 2


class TrtError(RuntimeError):
    pass


class Opts(C.Structure):
    _fields_ = [("max_depth", C.c_int), ("rr_threshold", C.c_int), ("seed_base", C.c_int),
                ("traversal", C.c_int), ("pool_paths", C.c_int), ("count_rays", C.c_int),
                ("time_kernels", C.c_int), ("reserved", C.c_int * 1)]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("samples", "closest_rays", "shadow_rays", "nodes_fetched",
                                          "tris_tested", "replays", "iterations", "kernel_launches",
                                          "nodes_closest", "tris_closest", "tree_closest", "tree_shadow",
                                          "check_violations")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class KernelTimes(C.Structure):
    _fields_ = [("regen_ms", C.c_float), ("extend_ms", C.c_float), ("shade_ms", C.c_float),
                ("shadow_ms", C.c_float), ("iterations", C.c_int), ("reserved", C.c_int * 3)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "reserved"}


class SceneInfo(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("n_objects", "n_ref_nodes", "n_lights", "n_textures", "n_wide_nodes",
                                       "n_wide_leaf_tris", "n_top_prims", "wide_node_bytes", "tri_record_bytes",
                                       "wide_depth", "builder")] + [("build_ms", C.c_float), ("n_underivable", C.c_int),
                                                                    ("reserved", C.c_int * 3)]

    def as_dict(self):
        return {n: (float(getattr(self, n)) if n == "build_ms" else int(getattr(self, n)))
                for n, _ in self._fields_ if n != "reserved"}


class Image(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("rgb", C.c_void_p)]


_lib = None


def lib():
    """The native library; raises if it has not been built (no fallback exists)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise TrtError(f"{LIB_PATH} is missing: build it with `make lib` or __graft_entry__.build()")
        L = C.CDLL(str(LIB_PATH))
        L.trt_last_error.restype = C.c_char_p
        L.trt_version.restype = C.c_char_p
        L.trt_stream.restype = C.c_void_p
        L.trt_stream.argtypes = [C.c_void_p]
        L.trt_free.argtypes = [C.c_void_p]
        L.trt_free.restype = None
        _lib = L
    return _lib


def _check(rc):
    if rc < 0:
        raise TrtError(f"trt error {rc}: {lib().trt_last_error().decode()}")
    return rc


def _ptr(x):
    """Device pointer from an int, None, or anything with data_ptr() (a torch tensor)."""
    if x is None:
        return None
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(int(x))


def _np_ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _f3(v):
    return (C.c_float * 3)(*[float(t) for t in v])


def default_opts(**kw) -> Opts:
    o = Opts()
    lib().trt_default_opts(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


# ---- host surface -------------------------------------------------------------------------
def load_obj(filename, offset=(0, 0, 0), scale=1.0, albedo=(0.75, 0.75, 0.75), metallic=0.0, roughness=1.0):
    """load_obj (reference include/loader.h:12): returns the appended objects as an OBJECT array."""
    L = lib()
    fn = str(filename).encode()
    n = _check(L.trt_load_obj(fn, None, 0, _f3(offset), C.c_float(scale), _f3(albedo), C.c_float(metallic),
                              C.c_float(roughness)))
    out = np.zeros(n, dtype=OBJECT)
    if n:
        _check(L.trt_load_obj(fn, _np_ptr(out), n, _f3(offset), C.c_float(scale), _f3(albedo),
                              C.c_float(metallic), C.c_float(roughness)))
    return out


class BVH:
    """BVH (reference include/bvh.h:33-51).  build() returns the reordered objects (the
    reference reorders its argument in place) and keeps the nodes."""

    def __init__(self):
        self.nodes = np.zeros(0, dtype=NODE)

    def build(self, objects):
        objs = np.ascontiguousarray(objects, dtype=OBJECT).copy()
        n = len(objs)
        nodes = np.zeros(max(2 * n, 1), dtype=NODE)
        if n:
            cnt = _check(lib().trt_bvh_build(_np_ptr(objs), n, _np_ptr(nodes), len(nodes)))
            self.nodes = nodes[:cnt].copy()
        return objs

    def get_nodes(self):
        return self.nodes


def collect_lights(objects):
    """Light list of reference src/main.cpp:88-96."""
    objs = np.ascontiguousarray(objects, dtype=OBJECT)
    n = _check(lib().trt_collect_lights(_np_ptr(objs), len(objs), None, 0))
    out = np.zeros(n, dtype=np.int32)
    if n:
        _check(lib().trt_collect_lights(_np_ptr(objs), len(objs), _np_ptr(out), n))
    return out


class CameraController:
    """CameraController (reference include/camera.h) driven by explicit yaw/pitch."""

    def __init__(self, position, yaw=-90.0, pitch=0.0, aperture=0.0, focus_dist=240.0):
        self.pos, self.yaw, self.pitch = tuple(position), yaw, pitch
        self.aperture, self.focus_dist = aperture, focus_dist

    def get_params(self, width, height):
        cam = np.zeros(1, dtype=CAMERA)
        _check(lib().trt_camera_params(_f3(self.pos), C.c_float(self.yaw), C.c_float(self.pitch),
                                       C.c_float(self.aperture), C.c_float(self.focus_dist), width, height,
                                       _np_ptr(cam)))
        return cam


def create_scene(config, asset_dir=None, grid=0):
    """Scene factory: 0 = create_cornell_box (reference src/scene.cpp:24), 1..5 = SURVEY 8(d) C1..C5.
    Returns (objects, texture_files)."""
    L = lib()
    ad = str(asset_dir or ASSET_DIR).encode()
    tex = C.create_string_buffer(4096)
    n = _check(L.trt_scene_create(config, ad, grid, None, 0, tex, len(tex)))
    out = np.zeros(n, dtype=OBJECT)
    _check(L.trt_scene_create(config, ad, grid, _np_ptr(out), n, tex, len(tex)))
    files = [t for t in tex.value.decode().split(";") if t]
    return out, files


def create_cornell_box(asset_dir=None):
    return create_scene(0, asset_dir)


def load_ppm(filename):
    w, h, p = C.c_int(), C.c_int(), C.c_void_p()
    _check(lib().trt_load_ppm(str(filename).encode(), C.byref(w), C.byref(h), C.byref(p)))
    n = w.value * h.value * 3
    arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_ubyte)), shape=(n,)).copy().reshape(h.value, w.value, 3)
    lib().trt_free(p)
    return arr


def write_earth_ppm(filename, w=2048, h=1024):
    _check(lib().trt_write_ppm_earth(str(filename).encode(), w, h))


def ensure_earth_ppm(asset_dir=None):
    p = Path(asset_dir or ASSET_DIR) / "earth.ppm"
    if not p.exists():
        p.parent.mkdir(parents=True, exist_ok=True)
        write_earth_ppm(p)
    return p


# Camera placement of the benchmark configs: pos, yaw, pitch, width, height, spp
CONFIGS = {
    1: dict(name="C1 cube", pos=(50, 50, 295.6), yaw=-90.0, pitch=0.0, width=640, height=480, spp=16),
    2: dict(name="C2 teapot", pos=(50, 45, 230), yaw=-90.0, pitch=-6.0, width=1920, height=1080, spp=64),
    3: dict(name="C3 cow+teddy", pos=(50, 45, 230), yaw=-90.0, pitch=-6.0, width=1920, height=1080, spp=256),
    4: dict(name="C4 pumpkin", pos=(50, 45, 230), yaw=-90.0, pitch=-6.0, width=3840, height=2160, spp=1024),
    5: dict(name="C5 teapot field", pos=(5, 90, 200), yaw=-90.0, pitch=-25.0, width=3840, height=2160, spp=16),
    0: dict(name="stock cornell", pos=(50, 50, 295.6), yaw=-90.0, pitch=0.0, width=1200, height=800, spp=16),
}


class HostScene:
    """A scene prepared the way reference src/main.cpp:79-101 prepares it: factory -> BVH::build
    (which reorders the objects) -> light list on the sorted array."""

    def __init__(self, objects, texture_files=()):
        bvh = BVH()
        self.objects = bvh.build(objects)
        self.nodes = bvh.get_nodes()
        self.lights = collect_lights(self.objects)
        self.texture_files = list(texture_files)

    @classmethod
    def from_config(cls, config, asset_dir=None, grid=0):
        objs, tex = create_scene(config, asset_dir, grid)
        if tex:
            ensure_earth_ppm(asset_dir)
        return cls(objs, tex)


def config_camera(config, width=None, height=None):
    c = CONFIGS[config]
    w, h = width or c["width"], height or c["height"]
    return CameraController(c["pos"], c["yaw"], c["pitch"]).get_params(w, h), w, h


# ---- renderer boundary ----------------------------------------------------------------------
class Context:
    """One GPU context of the core (trt_create .. trt_destroy)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        _check(lib().trt_create(device, C.byref(self._h)))
        self.device = device

    def close(self):
        if self._h:
            lib().trt_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # init_scene_data (reference include/renderer.h:35-38)
    def init_scene_data(self, objects, texture_files, nodes, light_indices, builder=0):
        """builder: BUILD_AUTO / BUILD_HOST_SAH / BUILD_DEVICE_LBVH (include/trt_capi.h).  nodes=None with the
        device builder builds the scene from the object array alone (no BVH::build on the host)."""
        objs = np.ascontiguousarray(objects, dtype=OBJECT)
        nd = np.ascontiguousarray(nodes, dtype=NODE) if nodes is not None else np.zeros(0, dtype=NODE)
        li = np.ascontiguousarray(light_indices, dtype=np.int32)
        imgs, keep = [], []
        for f in texture_files:
            a = np.ascontiguousarray(load_ppm(f))
            keep.append(a)
            imgs.append(Image(a.shape[1], a.shape[0], a.ctypes.data))
        arr = (Image * max(len(imgs), 1))(*imgs)
        _check(lib().trt_upload_scene_ex(self._h, _np_ptr(objs), len(objs), _np_ptr(nd) if len(nd) else None, len(nd),
                                         _np_ptr(li) if len(li) else None, len(li), arr, len(imgs), int(builder)))

    def upload_instanced(self, extra, unit, instances, light_indices, texture_files=()):
        """extra: OBJECT records placed first; unit: the mesh, parsed once; instances: (n, 4) float32 rows of
        (offset.xyz, scale).  The object array is generated and the BVH built on the device (trt_upload_instanced)."""
        ex = np.ascontiguousarray(extra, dtype=OBJECT)
        un = np.ascontiguousarray(unit, dtype=OBJECT)
        inst = np.ascontiguousarray(instances, dtype=np.float32).reshape(-1, 4)
        li = np.ascontiguousarray(light_indices, dtype=np.int32)
        imgs, keep = [], []
        for f in texture_files:
            a = np.ascontiguousarray(load_ppm(f))
            keep.append(a)
            imgs.append(Image(a.shape[1], a.shape[0], a.ctypes.data))
        arr = (Image * max(len(imgs), 1))(*imgs)
        _check(lib().trt_upload_instanced(self._h, _np_ptr(ex) if len(ex) else None, len(ex), _np_ptr(un), len(un),
                                          _np_ptr(inst), len(inst), _np_ptr(li) if len(li) else None, len(li), arr, len(imgs)))

    def get_objects(self):
        n = self.scene_info()["n_objects"]
        out = np.zeros(n, dtype=OBJECT)
        _check(lib().trt_get_objects(self._h, _np_ptr(out), n))
        return out

    def upload(self, scene: HostScene, builder=0):
        self.init_scene_data(scene.objects, scene.texture_files, scene.nodes, scene.lights, builder)

    def scene_info(self):
        s = SceneInfo()
        _check(lib().trt_scene_info_get(self._h, C.byref(s)))
        return s.as_dict()

    # launch_render_kernel (reference include/renderer.h:57), batched over frames
    def render(self, d_accum, width, height, first_frame_seed, n_frames, cam, opts=None, frame_stride=1):
        cam = np.ascontiguousarray(cam, dtype=CAMERA)
        _check(lib().trt_render(self._h, _ptr(d_accum), width, height, first_frame_seed, n_frames, frame_stride,
                                _np_ptr(cam), C.byref(opts) if opts is not None else None))

    def launch_render_kernel(self, accum_buffer, width, height, frame_seed, tx, ty, cam):
        self.render(accum_buffer, width, height, frame_seed, 1, cam)

    def render_to_host(self, h_accum, width, height, first_frame_seed, n_frames, cam, opts=None, frame_stride=1):
        """h_accum: numpy float32 array of w*h*4 or a pinned torch CPU tensor."""
        cam = np.ascontiguousarray(cam, dtype=CAMERA)
        p = _np_ptr(h_accum) if isinstance(h_accum, np.ndarray) else _ptr(h_accum)
        _check(lib().trt_render_to_host(self._h, p, width, height, first_frame_seed, n_frames, frame_stride,
                                        _np_ptr(cam), C.byref(opts) if opts is not None else None))

    def trace_primary(self, width, height, frame_seed, cam, traversal=TRAVERSE_FAST, seed_base=1984, d_id=None,
                      d_t=None, d_ray=None, d_fetched=None, d_entered=None, d_tris=None):
        cam = np.ascontiguousarray(cam, dtype=CAMERA)
        _check(lib().trt_trace_primary(self._h, width, height, frame_seed, _np_ptr(cam), traversal, seed_base,
                                       _ptr(d_id), _ptr(d_t), _ptr(d_ray), _ptr(d_fetched), _ptr(d_entered),
                                       _ptr(d_tris)))

    def trace_closest(self, d_rays, n, traversal, d_id, d_t=None):
        _check(lib().trt_trace_closest(self._h, _ptr(d_rays), n, traversal, _ptr(d_id), _ptr(d_t)))

    def trace_shadow(self, d_rays, n, traversal, d_occ):
        _check(lib().trt_trace_shadow(self._h, _ptr(d_rays), n, traversal, _ptr(d_occ)))

    def rng_states(self, width, height, frame_seed, first_pixel, n, d_states, seed_base=1984):
        _check(lib().trt_rng_states(self._h, width, height, frame_seed, seed_base, first_pixel, n, _ptr(d_states)))

    def tonemap(self, d_accum, width, height, frames, d_argb):
        _check(lib().trt_tonemap(self._h, _ptr(d_accum), width, height, frames, _ptr(d_argb)))

    def synchronize(self):
        _check(lib().trt_synchronize(self._h))

    def counters(self):
        c = Counters()
        _check(lib().trt_get_counters(self._h, C.byref(c)))
        return c.as_dict()

    def reset_counters(self):
        _check(lib().trt_reset_counters(self._h))

    def last_render_ms(self):
        ms = C.c_float()
        _check(lib().trt_last_render_ms(self._h, C.byref(ms)))
        return ms.value

    def kernel_times(self):
        k = KernelTimes()
        _check(lib().trt_kernel_times_get(self._h, C.byref(k)))
        return k.as_dict()

    def stream(self):
        return lib().trt_stream(self._h)

    def set_stream(self, cuda_stream):
        """Launch on a caller-owned stream (an int cudaStream_t, e.g. torch's current stream)."""
        _check(lib().trt_set_stream(self._h, C.c_void_p(int(cuda_stream)) if cuda_stream else None))
