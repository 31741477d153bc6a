"""CPU tier: the XORWOW skip-ahead algebra against golden vectors and
against the CUDA toolkit's own curand_init evaluated on the host (oracle/curand_host.cu)."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

GOLDEN = [  # frame, pixel, d, v0..v4
    (1, 0, 0xcca81940, [0xfa0d0a3d, 0xe72b68cd, 0xf8a42704, 0xdcd8f87c, 0xf3097c41]),
    (1, 1, 0xcca81940, [0xa81ccd2b, 0x9f111e06, 0xed423aec, 0x7d2ef9f3, 0x7fa09f27]),
    (1, 1200, 0xcca81940, [0xcfb7aeb4, 0x6c739ac8, 0xc5b55c44, 0x5150f248, 0x4888e7fd]),
    (1, 959999, 0xcca81940, [0x9bdba8c6, 0xde3056c3, 0x0429dd48, 0x31689919, 0x9cdbf8b1]),
    (1, 2073599, 0xcca81940, [0xc4eabace, 0xab55b2bd, 0x79022f6e, 0xdc2ffd4c, 0xec3fade9]),
    (1, 8294399, 0xcca81940, [0x447a9ffd, 0xdaea8390, 0x097dd335, 0x064b487b, 0x4dc2922d]),
    (2, 0, 0x913055bf, [0xbe9546bc, 0xa2a32c42, 0xf8a42704, 0xdcd8f87c, 0xb791b8c0]),
    (64, 3, 0x14c5a5d5, [0xce09a044, 0xf431ecee, 0x7b3c1c05, 0x5b54dfe2, 0x2468dfba]),
]


def host_state(trt, seed, sub):
    out = (C.c_uint32 * 6)()
    assert trt.lib().trt_xorwow_init_host(C.c_uint64(seed), C.c_uint64(sub), out) == 0
    return list(out)


@pytest.mark.parametrize("frame,pix,d,v", GOLDEN)
def test_golden_states(trt, frame, pix, d, v):
    assert host_state(trt, 1984 + frame, pix) == v + [d]


@pytest.mark.parametrize("frame,pix,d,v", GOLDEN)
def test_curand_header_agrees_with_goldens(ref, frame, pix, d, v):
    """Pins the known-answer source itself."""
    s = ref.xorwow_states(1984 + frame, [pix])[0]
    assert [int(x) for x in s] == v + [d]


def test_golden_outputs(ref):
    """First draws of frame 1 / pixel 0 (Appendix B.3): curand() words and curand_uniform floats."""
    u, f = ref.xorwow_draws(1985, 0, 4)
    assert [int(x) for x in u] == [0x5aba028c, 0xda9bd7bf, 0x65ade379, 0xf8538f43]
    assert f[0] == np.float32(0.354400784) and f[1] == np.float32(0.853940487)


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 2**31 - 1), pix=st.integers(0, 3840 * 2160 - 1))
def test_random_states_match_curand(trt, ref, seed, pix):
    assert host_state(trt, seed, pix) == [int(x) for x in ref.xorwow_states(seed, [pix])[0]]


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(1985, 4000), w=st.sampled_from([640, 1200, 1920, 3840, 37]), row=st.integers(0, 2159), col=st.integers(0, 36))
def test_row_column_decomposition(trt, ref, seed, w, row, col):
    """state(pixel = row*w + col) = (M^w)^row * M^col * v0 -- the split the kernels use."""
    out = (C.c_uint32 * 6)()
    assert trt.lib().trt_xorwow_rowcol_host(C.c_uint64(seed), w, row, col, out) == 0
    assert list(out) == [int(x) for x in ref.xorwow_states(seed, [row * w + col])[0]]
