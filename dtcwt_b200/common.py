"""``Pyramid``: the transform-domain value type (reference ``dtcwt/numpy/common.py:5-32``).

Device tensors are the storage; NumPy arrays are made lazily and memoised, the
pattern the reference's OpenCL backend uses (``dtcwt/opencl/transform2d.py:63-84``):

* ``lowpass_t`` / ``highpasses_t`` / ``scales_t`` -- ``torch.Tensor`` on the GPU.
  Sub-bands are stored planar (``[6][h][w]`` / ``[28][a][b][c]`` complex) and
  exposed with the reference's index order (``[h][w][6]`` / ``[a][b][c][28]``)
  through a permuted view, so ``highpasses_t[l][..., d]`` is band ``d`` and is
  contiguous.
* ``lowpass`` / ``highpasses`` / ``scales`` -- NumPy arrays with exactly the
  reference's shapes and dtypes, copied from the device on first access.  Code
  written against the reference (e.g. ``dtcwt.registration``) reads these.

One source of truth.  The reference's Pyramid has plain mutable attributes and
user code edits them in place (``p.highpasses[0][:] = 0`` to drop a level,
thresholding for denoising) before calling ``inverse``.  To keep that working,
a NumPy attribute that has been read (or assigned) becomes the authoritative
copy of that component: ``tensors()`` -- what every ``inverse`` calls -- uploads
it again and refreshes the ``*_t`` attribute.  Components whose NumPy side was never
touched cost nothing.  Code that works on the ``*_t`` tensors should not read the
NumPy attributes in between, or should call ``drop_numpy()`` after editing tensors.
"""
from __future__ import annotations

import numpy as np
import torch

__all__ = ["Pyramid", "pyramid_parts"]


def pyramid_parts(pyramid):
    """(lowpass, highpasses) of a Pyramid-like object: device tensors of our own :class:`Pyramid` (with edited NumPy
    attributes folded back, see :meth:`Pyramid.tensors`), else the object's ``lowpass`` / ``highpasses`` attributes
    (the reference's numpy Pyramid, reference ``dtcwt/numpy/common.py:8-11``: "any class which corresponds to this
    interface")."""
    if isinstance(pyramid, Pyramid):
        return pyramid.tensors()
    lo = getattr(pyramid, "lowpass_t", None)
    hs = getattr(pyramid, "highpasses_t", None)
    if lo is None or hs is None:
        lo, hs = pyramid.lowpass, pyramid.highpasses
    return lo, hs


def _to_numpy(t):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        return t.detach().cpu().numpy()
    return np.asarray(t)


def _like(arr, old):
    """NumPy array -> tensor where `old` lives (or on the CPU; the transforms move it)."""
    t = torch.from_numpy(np.ascontiguousarray(arr))
    if isinstance(old, torch.Tensor):
        t = t.to(old.device)
    return t


class Pyramid(object):
    def __init__(self, lowpass, highpasses, scales=None):
        self.lowpass_t = lowpass
        self.highpasses_t = tuple(highpasses)
        self.scales_t = tuple(scales) if scales is not None else None
        self._np = {}

    # ------------------------------------------------------------------ NumPy side (reference attribute names)
    @property
    def lowpass(self):
        if "lowpass" not in self._np:
            self._np["lowpass"] = _to_numpy(self.lowpass_t)
        return self._np["lowpass"]

    @lowpass.setter
    def lowpass(self, value):
        if isinstance(value, torch.Tensor):
            self.lowpass_t = value
            self._np.pop("lowpass", None)
        else:
            self._np["lowpass"] = np.asarray(value)

    @property
    def highpasses(self):
        if "highpasses" not in self._np:
            self._np["highpasses"] = tuple(_to_numpy(h) for h in self.highpasses_t)
        return self._np["highpasses"]

    @highpasses.setter
    def highpasses(self, value):
        value = tuple(value)
        if all(isinstance(v, torch.Tensor) or v is None for v in value):
            self.highpasses_t = value
            self._np.pop("highpasses", None)
        else:
            self._np["highpasses"] = tuple(None if v is None else _to_numpy(v) for v in value)

    @property
    def scales(self):
        if self.scales_t is None and "scales" not in self._np:
            return None
        if "scales" not in self._np:
            self._np["scales"] = tuple(_to_numpy(s) for s in self.scales_t)
        return self._np["scales"]

    @scales.setter
    def scales(self, value):
        if value is None:
            self.scales_t = None
            self._np.pop("scales", None)
            return
        value = tuple(value)
        if all(isinstance(v, torch.Tensor) for v in value):
            self.scales_t = value
            self._np.pop("scales", None)
        else:
            self._np["scales"] = tuple(_to_numpy(v) for v in value)

    # ------------------------------------------------------------------ tensor side
    def drop_numpy(self):
        """Forget the memoised NumPy copies: the ``*_t`` tensors are authoritative again."""
        self._np = {}

    def tensors(self):
        """-> (lowpass, highpasses) as tensors, after folding back any NumPy attribute that was read or
        assigned (it may have been edited in place, as code written for the reference does)."""
        if "lowpass" in self._np:
            self.lowpass_t = _like(self._np["lowpass"], self.lowpass_t)
        if "highpasses" in self._np:
            old = self.highpasses_t
            new = []
            for i, h in enumerate(self._np["highpasses"]):
                ref = old[i] if i < len(old) else None
                new.append(None if h is None else _like(h, ref if ref is not None else self.lowpass_t))
            self.highpasses_t = tuple(new)
        return self.lowpass_t, self.highpasses_t
