"""Record layouts of the reference's host interface as numpy dtypes: struct Vec
(include/common.h:24-97), Object (include/scene.h:30-55), LinearBVHNode (include/bvh.h:12-28), CameraParams
(include/scene.h:64-72).  Pure numpy: importing this module does not load libtrt_b200.so (bench.py's reference arm
uses it to hold the reference's own records)."""
import numpy as np

VEC = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("_", "<f4")])
OBJECT = np.dtype([
    ("v0", VEC), ("v1", VEC), ("v2", VEC), ("albedo", VEC), ("emission", VEC),
    ("metallic", "<f4"), ("roughness", "<f4"), ("ior", "<f4"), ("transmission", "<f4"),
    ("tex_id", "<i4"), ("pad1", "<f4"), ("pad2", "<f4"), ("pad3", "<f4"),
])
NODE = np.dtype([("min", VEC), ("max", VEC), ("a", "<i4"), ("b", "<i4"), ("axis", "<i4"), ("is_leaf", "<i4")])
CAMERA = np.dtype([("pos", VEC), ("cx", VEC), ("cy", VEC), ("dir", VEC),
                   ("lens_radius", "<f4"), ("focus_dist", "<f4"), ("_p", "<f4", (2,))])
assert OBJECT.itemsize == 112 and NODE.itemsize == 48 and CAMERA.itemsize == 80 and VEC.itemsize == 16
