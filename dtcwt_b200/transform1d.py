"""1-D DT-CWT on the GPU -- drop-in for ``dtcwt.numpy.Transform1d``.

Mirrors ``dtcwt/numpy/transform1d.py:14-184``: the transform runs down the
columns of an ``(n,)`` or ``(n, c)`` array (the columns are the batch); odd ``n``
raises ``ValueError`` (:70-71); a level whose input length is not a multiple of 4
is edge-extended by one sample per side (:95-96) and cropped by the inverse
(:164-165); a 1-D input comes back 1-D (:177-180).
"""
from __future__ import annotations

import numpy as np

from . import _ops
from .coeffs import biort as _biort, qshift as _qshift
from .common import Pyramid, pyramid_parts
from .defaults import DEFAULT_BIORT, DEFAULT_QSHIFT

__all__ = ["Transform1d"]


def _vec(h):
    return np.asarray(h, dtype=np.float64).reshape(-1)


class Transform1d(object):
    def __init__(self, biort=DEFAULT_BIORT, qshift=DEFAULT_QSHIFT):
        # like the reference (:22-24) names are resolved at call time
        self.biort = biort
        self.qshift = qshift

    def _taps(self):
        try:
            b = _biort(self.biort)
        except TypeError:
            b = self.biort
        try:
            q = _qshift(self.qshift)
        except TypeError:
            q = self.qshift
        h0o, g0o, h1o, g1o = (_vec(h) for h in b[:4])
        h0a, h0b, g0a, g0b, h1a, h1b, g1a, g1b = (_vec(h) for h in q[:8])
        return dict(h0o=h0o, g0o=g0o, h1o=h1o, g1o=g1o, h0a=h0a, h0b=h0b, g0a=g0a, g0b=g0b,
                    h1a=h1a, h1b=h1b, g1a=g1a, g1b=g1b)

    def forward(self, X, nlevels=3, include_scale=False):
        t = self._taps()
        X = _ops.as_real_tensor(X)
        if X.dim() == 1:
            X = X.unsqueeze(1)
        if X.dim() != 2:
            raise ValueError("X must be a vector or a matrix whose columns are transformed")
        if X.shape[0] % 2 != 0:
            raise ValueError("Size of input X must be a multiple of 2")
        if nlevels == 0:
            return Pyramid(X, (), ()) if include_scale else Pyramid(X, ())
        if t["h0o"].shape[0] % 2 == 0 or t["h1o"].shape[0] % 2 == 0:
            raise ValueError("even-length biorthogonal filters are not supported by the 1-D transform")
        Yh, Ysc = [], []
        Hi = _ops.colfilter(X, t["h1o"], 0)
        Lo = _ops.colfilter(X, t["h0o"], 0)
        Yh.append(_ops.pack1d(Hi))
        Ysc.append(Lo)
        for _ in range(1, nlevels):
            pad = (1, 1) if Lo.shape[0] % 4 else (0, 0)
            Hi = _ops.coldfilt(Lo, t["h1b"], t["h1a"], 0, pad)
            Lo = _ops.coldfilt(Lo, t["h0b"], t["h0a"], 0, pad)
            Yh.append(_ops.pack1d(Hi))
            Ysc.append(Lo)
        return Pyramid(Lo, tuple(Yh), tuple(Ysc)) if include_scale else Pyramid(Lo, tuple(Yh))

    def inverse(self, pyramid, gain_mask=None):
        t = self._taps()
        Lo, Yh = pyramid_parts(pyramid)
        Lo = _ops.as_real_tensor(Lo, "lowpass")
        a = len(Yh)
        if a == 0:
            return Lo
        flat = Lo.dim() == 1
        if flat:
            Lo = Lo.unsqueeze(1)
        Yh = [_ops.as_complex_tensor(h, Lo.dtype) for h in Yh]
        Yh = [(h.unsqueeze(1) if h.dim() == 1 else h).contiguous() for h in Yh]
        gm = np.ones(a) if gain_mask is None else np.asarray(gain_mask, dtype=np.float64).reshape(-1)
        if gm.shape[0] != a:
            raise ValueError("gain_mask must have one entry per level")
        for lev in range(a - 1, 0, -1):
            if Lo.shape[0] != 2 * Yh[lev].shape[0] or Lo.shape[1] != Yh[lev].shape[1]:
                raise ValueError("Yh sizes are not valid for DTWAVEIFM")
            have, need = 2 * Lo.shape[0], 2 * Yh[lev - 1].shape[0]
            if have == need:
                crop = 0
            elif have - 2 == need:
                crop = 1
            else:
                raise ValueError("Yh sizes are not valid for DTWAVEIFM")
            Hi = _ops.unpack1d(Yh[lev], gm[lev])
            out = _ops.colifilt(Lo, t["g0b"], t["g0a"], 0, crop)
            _ops.colifilt(Hi, t["g1b"], t["g1a"], 0, crop, out=out, accumulate=True)
            Lo = out
        if Lo.shape[0] != 2 * Yh[0].shape[0] or Lo.shape[1] != Yh[0].shape[1]:
            raise ValueError("Yh sizes are not valid for DTWAVEIFM")
        Hi = _ops.unpack1d(Yh[0], gm[0])
        Z = _ops.colfilter(Lo, t["g0o"], 0)
        _ops.colfilter(Hi, t["g1o"], 0, out=Z, accumulate=True)
        return Z.reshape(-1) if Z.shape[1] == 1 else Z
