"""Generates tests/golden/*.npz ON A B200 from the UNMODIFIED reference kernel (oracle/_ref, built from
/root/reference by oracle/Makefile): first-hit ids (ID-as-emission scene), d_min of the instrumented
restatement (bit pattern), and the 8-spp accumulation buffer of the unmodified kernel, for C1 and C2 at
reduced resolution, frame seeds 1 and 2.  The fixtures let the parity tests run where the reference
sources (and so oracle/_ref) are not available.  MUFU approximations are hardware: the fixtures are
valid for sm_100 (B200) only.
Usage (GPU box): python tools/make_goldens.py <out_dir>"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import torch
import tryraytrace_b200 as trt
import reflib

out = Path(sys.argv[1] if len(sys.argv) > 1 else ROOT / "tests" / "golden")
out.mkdir(parents=True, exist_ok=True)
CASES = {1: (320, 240), 2: (320, 180)}
for config, (w, h) in CASES.items():
    sc = trt.HostScene.from_config(config)
    reflib.init_scene(sc)
    cam, w, h = trt.config_camera(config, w, h)
    n = w * h
    data = {"width": w, "height": h, "config": config, "camera": cam.view(np.uint8).copy(), "spp": 8}
    for frame in (1, 2):
        data[f"ids_f{frame}"] = reflib.first_hit_ids(w, h, frame, cam)
        t = torch.zeros(n, device="cuda")
        torch.cuda.synchronize()
        reflib.primary_counts(w, h, frame, cam, None, t)
        data[f"dmin_bits_f{frame}"] = t.cpu().numpy().view(np.uint32)
    acc, stage = torch.zeros(n * 4, device="cuda"), torch.zeros(n * 4, device="cuda")
    torch.cuda.synchronize()
    reflib.render_frames(acc, stage, w, h, 1, 8, cam, 0)
    data["accum_8spp"] = acc.cpu().numpy()
    np.savez_compressed(out / f"reference_c{config}_{w}x{h}.npz", **data)
    print("wrote", out / f"reference_c{config}_{w}x{h}.npz")
