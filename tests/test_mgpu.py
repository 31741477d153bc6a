"""The single-process multi-GPU layer (include/trt_mgpu.h, libtrt_b200_mgpu.so): sample-index split
over the GPUs of one box + one ncclAllReduce of the accumulation buffer per pass, for C / C++ hosts.
It is exercised through a plain C program in a subprocess (tests/dropin/mgpu_main.c), never loaded
into this process: the library links NCCL, and the torch-based harness must keep its own copy.
CPU tier: the library exports what the header declares and the C host links.  GPU tier: the pass on
all visible GPUs (1 on the single-GPU box) equals the single-GPU pass up to FP32 summation order."""
import json
import os
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
LIBDIR = ROOT / "tryraytrace_b200" / "lib"
LIB = LIBDIR / "libtrt_b200_mgpu.so"
SRC = ROOT / "tests" / "dropin" / "mgpu_main.c"
BIN = ROOT / "build" / "mgpu_main"


def build_host():
    BIN.parent.mkdir(exist_ok=True)
    cmd = ["gcc", "-O2", f"-I{ROOT / 'include'}", "-I/usr/local/cuda/include", str(SRC), "-o", str(BIN), f"-L{LIBDIR}",
           "-ltrt_b200_mgpu", "-ltrt_b200", "-lm", f"-Wl,-rpath,{LIBDIR}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return BIN


def test_mgpu_library_exports_the_header_and_the_c_host_links():
    assert LIB.exists(), "libtrt_b200_mgpu.so was not built (make lib)"
    declared = set(re.findall(r"\b(trt_mgpu_\w+)\s*\(", (ROOT / "include" / "trt_mgpu.h").read_text()))
    exported = subprocess.run(["nm", "-D", "--defined-only", str(LIB)], capture_output=True, text=True).stdout
    missing = [s for s in declared if f" T {s}" not in exported]
    assert declared and not missing, f"not exported: {missing}"
    assert build_host().exists()


@pytest.mark.gpu
def test_multi_gpu_pass_equals_single_gpu_pass(assets):
    import torch
    n = max(1, min(torch.cuda.device_count(), 8))
    if "TRT_EXPECT_GPUS" in os.environ:  # multi-GPU boxes: the driver of the run states how many it handed out
        assert n == int(os.environ["TRT_EXPECT_GPUS"]), f"expected {os.environ['TRT_EXPECT_GPUS']} GPUs, see {n}"
    b = build_host() if not BIN.exists() else BIN
    # single node: keep NCCL's bootstrap off the network interfaces it would otherwise probe
    env = dict(os.environ, LD_LIBRARY_PATH=f"{LIBDIR}:{os.environ.get('LD_LIBRARY_PATH', '')}", NCCL_SOCKET_IFNAME="lo",
               NCCL_IB_DISABLE="1")
    for config, w, h, frames in ((2, 480, 270, 7),):  # 7 frames: an uneven split over an even GPU count
        r = subprocess.run([str(b), str(assets), str(config), str(w), str(h), str(frames), str(n)], capture_output=True,
                           text=True, env=env, timeout=600)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        assert r.returncode == 0 and lines, r.stdout[-400:] + r.stderr[-400:]
        d = json.loads(lines[-1])
        print(d)
        assert d["n_gpus"] == n and d["max_rel_diff"] < 1e-5 and d["sum_multi"] > 0
        if torch.cuda.device_count() > 1:
            assert d["n_gpus"] > 1, "more than one GPU is visible but the pass ran on one"
