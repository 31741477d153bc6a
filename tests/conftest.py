"""pytest configuration: the `gpu` marker, and fixtures that locate the native libraries.

CPU tier (`-m "not gpu"`): host logic, the oracles against golden vectors, C-ABI exports.
GPU tier (`-m gpu`): parity of the CUDA path against the reference oracle, through the C ABI.
"""
import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(Path(__file__).resolve().parent))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def trt():
    import tryraytrace_b200 as t
    t.lib()
    return t


@pytest.fixture(scope="session")
def ref():
    import reflib
    if not reflib.available():
        pytest.skip("oracle/_ref/libtrt_ref.so not built (run `make oracle` where /root/reference exists)")
    return reflib


@pytest.fixture(scope="session")
def assets():
    d = ROOT / "assets"
    if not (d / "teapot.obj").exists():
        pytest.skip("mesh assets not staged (run `make assets` where /root/reference exists)")
    return d
