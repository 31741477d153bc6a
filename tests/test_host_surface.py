"""CPU tier: the host surface (loader, BVH builder, camera, light list, scene factories)
against the reference's own compiled code (oracle/_ref) and the survey's golden values."""
import re
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
MESHES = ["cube.obj", "temp.obj", "teddy.obj", "cow.obj", "teapot.obj", "pumpkin.obj"]
# SURVEY Appendix C
TRI_COUNTS = {"cube.obj": 12, "temp.obj": 25, "teddy.obj": 3192, "cow.obj": 5804, "teapot.obj": 6320, "pumpkin.obj": 10000}


def canon(a):
    """Bytes of a record array with the pad lane of every Vec zeroed (padding carries no meaning
    and the reference leaves whatever was on its stack there)."""
    a = a.copy()
    for name in a.dtype.names:
        if a.dtype[name].names and "_" in a.dtype[name].names:
            a[name]["_"] = 0
    for pad in ("pad1", "pad2", "pad3", "_p"):
        if pad in a.dtype.names:
            a[pad] = 0
    return a.tobytes()


def test_record_layouts(trt):
    # SURVEY Appendix B.1
    assert trt.OBJECT.itemsize == 112 and trt.NODE.itemsize == 48 and trt.CAMERA.itemsize == 80
    f = trt.OBJECT.fields
    assert [f[k][1] for k in ("v0", "v1", "v2", "albedo", "emission", "metallic", "roughness", "ior", "transmission", "tex_id", "pad1")] == \
        [0, 16, 32, 48, 64, 80, 84, 88, 92, 96, 100]
    n = trt.NODE.fields
    assert [n[k][1] for k in ("min", "max", "a", "b", "axis", "is_leaf")] == [0, 16, 32, 36, 40, 44]
    c = trt.CAMERA.fields
    assert [c[k][1] for k in ("pos", "cx", "cy", "dir", "lens_radius", "focus_dist")] == [0, 16, 32, 48, 64, 68]


def test_capi_exports_every_declared_symbol(trt):
    """The shared library loads and exports every function include/trt_capi.h declares."""
    hdr = (ROOT / "include" / "trt_capi.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(trt_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 25
    L = trt.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert b"sm_100a" in L.trt_version()


def test_error_reporting_without_device(trt):
    """No CPU fallback: without a CUDA device trt_create fails loudly (on a GPU box it succeeds)."""
    import torch
    h = C.c_void_p()
    rc = trt.lib().trt_create(0, C.byref(h))
    if torch.cuda.is_available():
        assert rc == 0
        trt.lib().trt_destroy(h)
    else:
        assert rc == -2
        assert b"no CUDA device" in trt.lib().trt_last_error()
    # argument errors are reported, not crashed on
    assert trt.lib().trt_create(0, None) == -1
    assert trt.lib().trt_render(None, None, 1, 1, 1, 1, 1, None, None) == -1


@pytest.mark.parametrize("mesh", MESHES)
def test_load_obj_matches_reference(trt, ref, assets, mesh):
    args = dict(offset=(48.0, 5.0, 80.0), scale=14.0, albedo=(0.75, 0.25, 0.5), metallic=0.25, roughness=0.75)
    mine = trt.load_obj(assets / mesh, **args)
    theirs = ref.load_obj(trt.OBJECT, assets / mesh, args["offset"], args["scale"], args["albedo"], args["metallic"], args["roughness"])
    assert len(mine) == TRI_COUNTS[mesh]
    assert canon(mine) == canon(theirs)


def test_load_obj_parse_rules(trt, tmp_path):
    """Reference src/loader.cpp:45-94: only bare-index triangles, bounds-checked, extra indices ignored."""
    p = tmp_path / "m.obj"
    p.write_text("# comment\nv 0 0 0\nv 1 0 0\nv 0 1 0\nv 0 0 1\nvn 0 0 1\nvt 0 0\n"
                 "f 1 2 3\nf 1/1 2/2 3/3\nf 1 2 3 4\nf 1 2 9\nf 0 1 2\nf 2 3 4\ng grp\n")
    o = trt.load_obj(p, offset=(1, 2, 3), scale=2.0)
    assert len(o) == 3  # plain, first three of the quad, last; slashes / out-of-range / zero index dropped
    assert tuple(o[0]["v1"])[:3] == (3.0, 2.0, 3.0)
    assert (o["tex_id"] == -1).all() and (o["ior"] == 0).all() and (o["transmission"] == 0).all()
    assert (o["emission"]["x"] == 0).all()
    assert len(trt.load_obj(tmp_path / "missing.obj")) == 0  # prints and returns, like the reference


def test_empty_and_degenerate_inputs(trt, tmp_path):
    p = tmp_path / "e.obj"
    p.write_text("")
    assert len(trt.load_obj(p)) == 0
    bvh = trt.BVH()
    assert len(bvh.build(np.zeros(0, dtype=trt.OBJECT))) == 0 and len(bvh.get_nodes()) == 0
    one = np.zeros(1, dtype=trt.OBJECT)
    one["v1"]["x"] = 1
    one["v2"]["y"] = 1
    bvh.build(one)
    nd = bvh.get_nodes()
    assert len(nd) == 1 and nd[0]["is_leaf"] == 1 and nd[0]["b"] == 1
    # flat axis padded by 1e-3 (reference src/bvh.cpp:19-27)
    assert nd[0]["min"]["z"] == np.float32(-1e-3) and nd[0]["max"]["z"] == np.float32(1e-3)


@pytest.mark.parametrize("mesh", MESHES)
def test_bvh_matches_reference(trt, ref, assets, mesh):
    objs = trt.load_obj(assets / mesh, offset=(0, 0, 0), scale=1.0)
    bvh = trt.BVH()
    mine_objs = bvh.build(objs)
    ref_objs, ref_nodes = ref.bvh_build(trt.OBJECT, trt.NODE, objs)
    assert canon(mine_objs) == canon(ref_objs)
    assert canon(bvh.get_nodes()) == canon(ref_nodes)


@pytest.mark.parametrize("config", [1, 2, 3, 4])
def test_config_scene_bvh_matches_reference(trt, ref, assets, config):
    objs, tex = trt.create_scene(config, assets)
    bvh = trt.BVH()
    mine_objs = bvh.build(objs)
    ref_objs, ref_nodes = ref.bvh_build(trt.OBJECT, trt.NODE, objs)
    assert canon(mine_objs) == canon(ref_objs)
    assert canon(bvh.get_nodes()) == canon(ref_nodes)
    assert len(bvh.get_nodes()) == 2 * len(objs) - 1


def test_cube_bvh_golden(trt, assets):
    """SURVEY Appendix B.2 (GCC 13.3 libstdc++ tie order)."""
    objs = trt.load_obj(assets / "cube.obj")
    faces = {canon(objs[i:i + 1]): i for i in range(len(objs))}
    bvh = trt.BVH()
    srt = bvh.build(objs)
    nodes = bvh.get_nodes()
    assert len(nodes) == 23
    assert [faces[canon(srt[i:i + 1])] for i in range(len(srt))] == [10, 6, 0, 11, 4, 2, 7, 5, 8, 1, 3, 9]
    assert "".join("L" if n["is_leaf"] else "I" for n in nodes) == "IIILILLILILLIILILLILILL"
    inner = {i: (int(n["axis"]), int(n["b"])) for i, n in enumerate(nodes) if not n["is_leaf"]}
    assert inner == {0: (0, 12), 1: (1, 7), 2: (0, 4), 4: (1, 6), 7: (0, 9), 9: (1, 11), 12: (1, 18), 13: (0, 15),
                     15: (0, 17), 18: (0, 20), 20: (0, 22)}
    r = nodes[0]
    assert tuple(r["min"])[:3] == (np.float32(-1.001),) * 3 and tuple(r["max"])[:3] == (np.float32(1.001),) * 3
    l3 = nodes[3]
    assert tuple(l3["min"])[:3] == (np.float32(-1.001), -1.0, -1.0) and tuple(l3["max"])[:3] == (np.float32(-0.999), 1.0, 1.0)
    # invariants of Appendix A.4
    leaves = [n for n in nodes if n["is_leaf"]]
    assert [int(n["a"]) for n in leaves] == list(range(12)) and all(n["b"] == 1 for n in leaves)
    assert all(int(n["a"]) == i + 1 for i, n in enumerate(nodes) if not n["is_leaf"])


def test_cornell_matches_reference(trt, ref, assets, monkeypatch):
    monkeypatch.chdir(ROOT)  # create_cornell_box reads assets/teapot.obj relative to the cwd
    mine, tex = trt.create_cornell_box(assets)
    buf = np.zeros(len(mine) + 16, dtype=trt.OBJECT)
    t = C.create_string_buffer(1024)
    n = ref.ref().ref_create_cornell(buf.ctypes.data_as(C.c_void_p), len(buf), t, 1024)
    assert n == len(mine) == 6327
    assert canon(buf[:n]) == canon(mine)
    assert t.value.decode() == "assets/earth.ppm" and Path(tex[0]).name == "earth.ppm"
    lights = trt.collect_lights(mine)
    assert list(lights) == [6]


@pytest.mark.parametrize("pitch_units,wh", [(0.0, (1200, 800)), (60.0, (1920, 1080)), (250.0, (3840, 2160)), (-35.0, (640, 480))])
def test_camera_params_match_reference(trt, ref, pitch_units, wh):
    """get_params (reference src/camera.cpp:139-163); the reference reaches a pitch only through
    process_mouse, so the same rotation is applied on both sides."""
    pos = (50.0, 45.0, 230.0)
    theirs = ref.camera_params(trt.CAMERA, pos, 0.0, pitch_units, *wh)
    pitch = np.float32(0.0) - np.float32(pitch_units) * np.float32(0.1)
    mine = trt.CameraController(pos, yaw=-90.0, pitch=float(pitch)).get_params(*wh)
    assert canon(mine) == canon(theirs)


def test_ppm_roundtrip(trt, tmp_path):
    p = tmp_path / "e.ppm"
    trt.write_earth_ppm(p, 64, 32)
    img = trt.load_ppm(p)
    assert img.shape == (32, 64, 3)
    x, y = 63, 31
    assert tuple(img[y, x]) == (255, 255, (x ^ y) & 255)
    assert tuple(img[0, 0]) == (0, 0, 0)
    with pytest.raises(trt.TrtError):
        trt.load_ppm(tmp_path / "nope.ppm")
    bad = tmp_path / "bad.ppm"
    bad.write_bytes(b"P5\n2 2\n255\n0000")
    with pytest.raises(trt.TrtError):
        trt.load_ppm(bad)


@pytest.mark.parametrize("config", [2, 4])
def test_reference_only_scene_equals_the_library_scene(trt, ref, assets, config):
    """bench.py --impl reference builds its scene with the reference's compiled host code alone
    (reflib.ReferenceScene: create_cornell_box shell, load_obj, BVH::build); it is the same scene, byte for byte,
    as the one the library arm renders."""
    theirs = ref.ReferenceScene(config, trt.OBJECT, trt.NODE, assets)
    mine = trt.HostScene.from_config(config, assets)
    assert canon(theirs.objects) == canon(mine.objects)
    assert theirs.nodes.tobytes() == mine.nodes.tobytes()
    assert list(theirs.lights) == list(mine.lights)
    cam_t, w, h = ref.ReferenceScene.camera(config, trt.CAMERA)
    cam_m, w2, h2 = trt.config_camera(config)
    assert (w, h) == (w2, h2) and canon(np.asarray(cam_t)) == canon(np.asarray(cam_m))
