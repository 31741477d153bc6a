"""Shared helpers of the GPU parity tests (test infrastructure)."""
import numpy as np
import torch


def dev_zeros(n, dtype):
    """Device buffer handed to the library, which works on its OWN (non-blocking) stream: the fill
    that torch enqueues on its stream has to be complete before the library may touch the buffer."""
    t = torch.zeros(n, dtype=dtype, device="cuda")
    torch.cuda.synchronize()
    return t


def psnr_8bit(a_argb, b_argb):
    """PSNR between two ARGB8888 images over the three colour channels."""
    a = np.stack([(a_argb >> s) & 255 for s in (16, 8, 0)], -1).astype(np.float64)
    b = np.stack([(b_argb >> s) & 255 for s in (16, 8, 0)], -1).astype(np.float64)
    mse = np.mean((a - b) ** 2)
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


class SceneCache:
    """Builds each config scene once per test session (host side) and remembers it."""

    def __init__(self, trt, assets):
        self.trt, self.assets, self._s = trt, assets, {}

    def get(self, config, grid=0):
        k = (config, grid)
        if k not in self._s:
            self._s[k] = self.trt.HostScene.from_config(config, self.assets, grid)
        return self._s[k]
