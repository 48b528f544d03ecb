"""``Pyramid``: the transform-domain value type (reference ``dtcwt/numpy/common.py:5-32``).

Device tensors are the storage; NumPy views are made lazily and memoised, the
pattern the reference's OpenCL backend uses (``dtcwt/opencl/transform2d.py:63-84``):

* ``lowpass_t`` / ``highpasses_t`` / ``scales_t`` -- ``torch.Tensor`` on the GPU.
  Sub-bands are stored planar (``[6][h][w]`` / ``[28][a][b][c]`` complex) and
  exposed with the reference's index order (``[h][w][6]`` / ``[a][b][c][28]``)
  through a permuted view, so ``highpasses_t[l][..., d]`` is band ``d`` and is
  contiguous.
* ``lowpass`` / ``highpasses`` / ``scales`` -- NumPy arrays with exactly the
  reference's shapes and dtypes, copied from the device on first access.  Code
  written against the reference (e.g. ``dtcwt.registration``) reads these.
"""
from __future__ import annotations

import numpy as np
import torch

__all__ = ["Pyramid"]


def _to_numpy(t):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        return t.detach().cpu().numpy()
    return np.asarray(t)


class Pyramid(object):
    def __init__(self, lowpass, highpasses, scales=None):
        self.lowpass_t = lowpass
        self.highpasses_t = tuple(highpasses)
        self.scales_t = tuple(scales) if scales is not None else None
        self._np = {}

    @property
    def lowpass(self):
        if "lowpass" not in self._np:
            self._np["lowpass"] = _to_numpy(self.lowpass_t)
        return self._np["lowpass"]

    @property
    def highpasses(self):
        if "highpasses" not in self._np:
            self._np["highpasses"] = tuple(_to_numpy(h) for h in self.highpasses_t)
        return self._np["highpasses"]

    @property
    def scales(self):
        if self.scales_t is None:
            return None
        if "scales" not in self._np:
            self._np["scales"] = tuple(_to_numpy(s) for s in self.scales_t)
        return self._np["scales"]
