"""3-D DT-CWT on the GPU -- drop-in for ``dtcwt.numpy.Transform3d``.

Mirrors ``dtcwt/numpy/transform3d.py:15-526``: ``Transform3d(biort, qshift,
ext_mode=4)``, ``forward(X, nlevels=3, include_scale=False,
discard_level_1=False)``, ``inverse(pyramid)``; 28 complex sub-bands per level in
the reference's order (7 octants HLL LHL HHL LLH HLH LHH HHH x (p, q, r, s),
:280-288); ``ext_mode`` 4 / 8 padding (:322-335) and cropping (:505-524).

The reference filters 2-D slices of an octant "work cube" in Python loops; here
every filter call processes the whole (batched) volume along one axis, in the
reference's axis order: 2, 1, 0 forward and 1, 0, 2 inverse.

Reference quirk NOT reproduced: its ``discard_level_1`` inverse writes the last
pass back transposed (:452-454), returning the reconstruction with axes 0 and 2
swapped (and raising for non-cubic volumes).  This class returns the correctly
oriented volume; ``tests/test_oracle.py`` pins the relationship.
"""
from __future__ import annotations

import numpy as np

from . import _ops
from .coeffs import biort as _biort, qshift as _qshift
from .common import Pyramid, pyramid_parts
from .defaults import DEFAULT_BIORT, DEFAULT_QSHIFT

__all__ = ["Transform3d"]

# (axis0, axis1, axis2) filter type, 0 = lowpass 1 = highpass, in output order
_OCTANTS = ((0, 1, 0), (1, 0, 0), (1, 1, 0), (0, 0, 1), (0, 1, 1), (1, 0, 1), (1, 1, 1))


def _vec(h):
    return np.asarray(h, dtype=np.float64).reshape(-1)


class Transform3d(object):
    def __init__(self, biort=DEFAULT_BIORT, qshift=DEFAULT_QSHIFT, ext_mode=4):
        try:
            self.biort = _biort(biort)
        except TypeError:
            self.biort = biort
        try:
            self.qshift = _qshift(qshift)
        except TypeError:
            self.qshift = qshift
        self.ext_mode = ext_mode

    def _taps(self):
        if len(self.biort) not in (4, 6):
            raise ValueError("Biort wavelet must have 6 or 4 components.")
        if len(self.qshift) not in (8, 12):
            raise ValueError("Qshift wavelet must have 12 or 8 components.")
        if self.ext_mode != 4 and self.ext_mode != 8:
            raise ValueError("ext_mode must be one of 4 or 8")
        h0o, g0o, h1o, g1o = (_vec(h) for h in self.biort[:4])
        h0a, h0b, g0a, g0b, h1a, h1b, g1a, g1b = (_vec(h) for h in self.qshift[:8])
        return dict(h0o=h0o, g0o=g0o, h1o=h1o, g1o=g1o, h0a=h0a, h0b=h0b, g0a=g0a, g0b=g0b,
                    h1a=h1a, h1b=h1b, g1a=g1a, g1b=g1b)

    # ------------------------------------------------------------------ public API
    def forward(self, X, nlevels=3, include_scale=False, discard_level_1=False):
        t = self._taps()
        X = _ops.as_real_tensor(X)
        if X.dim() > 3:
            raise ValueError("forward() takes one volume; use forward_channels for a batch [N][D0][D1][D2]")
        while X.dim() < 3:
            X = X.unsqueeze(-1)          # numpy.atleast_3d appends axes (reference :87)
        p = self._forward_n(X.unsqueeze(0), t, nlevels, include_scale, discard_level_1)
        hp = tuple(None if h is None else h[0] for h in p.highpasses_t)
        sc = None if p.scales_t is None else tuple(s[0] for s in p.scales_t)
        return Pyramid(p.lowpass_t[0], hp, sc)

    def forward_channels(self, X, nlevels=3, include_scale=False, discard_level_1=False):
        """Batched forward: X is ``[N][D0][D1][D2]``."""
        t = self._taps()
        X = _ops.as_real_tensor(X)
        if X.dim() != 4:
            raise ValueError("forward_channels needs a [N][D0][D1][D2] batch")
        return self._forward_n(X, t, nlevels, include_scale, discard_level_1)

    def inverse(self, pyramid):
        t = self._taps()
        Yl, Yh = pyramid_parts(pyramid)
        Yl = _ops.as_real_tensor(Yl, "lowpass")
        batched = Yl.dim() == 4
        if not batched:
            Yl = Yl.unsqueeze(0)
        planar = []
        for h in Yh:
            if h is None:
                planar.append(None)
                continue
            h = _ops.as_complex_tensor(h, Yl.dtype)
            if h.shape[-1] != 28:
                raise ValueError("3-D highpass arrays must have 28 sub-bands on their last axis")
            if h.dim() == 4:
                h = h.unsqueeze(0)
            planar.append(h.permute(0, 4, 1, 2, 3).contiguous())
        Z = self._inverse_n(Yl, planar, t)
        return Z if batched else Z[0]

    inverse_channels = inverse

    # ------------------------------------------------------------------ forward
    def _forward_n(self, X, t, nlevels, include_scale, discard_level_1):
        Yl = X
        Yh, Ysc = [None] * nlevels, [None] * nlevels
        for lev in range(nlevels):
            if lev == 0:
                mult = 2 if self.ext_mode == 4 else 4
                if any(int(s) % mult for s in Yl.shape[1:]):
                    raise ValueError("Input shape should be a multiple of %d in each direction when "
                                     "self.ext_mode == %d" % (mult, self.ext_mode))
                even = t["h0o"].shape[0] % 2 == 0
                if even != (t["h1o"].shape[0] % 2 == 0):
                    raise ValueError("level-1 lowpass and highpass filters must both have odd or both have even length")
                if even and discard_level_1:
                    raise ValueError("discard_level_1 needs odd-length level-1 filters (the reference's "
                                     "_level1_xfm_no_highpass writes n+1 samples into n, transform3d.py:304-313)")
                if even:
                    Yl, Yh[0] = self._split_even(Yl, t)
                elif discard_level_1:
                    fused = _ops.lowpass3d(Yl, t["h0o"])
                    if fused is not None:
                        Yl = fused
                    else:
                        for ax in (3, 2, 1):     # reference _level1_xfm_no_highpass :291-315
                            Yl = _ops.colfilter(Yl, t["h0o"], ax)
                else:
                    fused = _ops.fwd3d_level1(Yl, t["h0o"], t["h1o"])
                    if fused is not None:
                        Yl, Yh[0] = fused
                    else:
                        Yl, Yh[0] = self._split(Yl, lambda A, ax, hi: _ops.colfilter(A, t["h1o"] if hi else t["h0o"], ax))
            else:
                n = 1 if self.ext_mode == 4 else 2
                pads = {ax: ((n, n) if int(Yl.shape[ax]) % (4 * n) else (0, 0)) for ax in (1, 2, 3)}

                def dfilt(A, ax, hi, pads=pads):
                    if hi:
                        return _ops.coldfilt(A, t["h1b"], t["h1a"], ax, pads[ax])
                    return _ops.coldfilt(A, t["h0b"], t["h0a"], ax, pads[ax])

                fused = _ops.fwd3d_levelq(Yl, t["h0b"], t["h0a"], t["h1b"], t["h1a"], [pads[ax][0] for ax in (1, 2, 3)])
                if fused is not None:
                    Yl, Yh[lev] = fused
                else:
                    Yl, Yh[lev] = self._split(Yl, dfilt)
            Ysc[lev] = Yl
        views = tuple(None if h is None else h.permute(0, 2, 3, 4, 1) for h in Yh)
        return Pyramid(Yl, views, tuple(Ysc)) if include_scale else Pyramid(Yl, views)

    @staticmethod
    def _split_even(X, t):
        """Level 1 with even-length taps (e.g. Haar; reference transform3d.py:223-251, tests/test_xfm3.py:42-58): every
        axis pass returns n+1 samples, LLL keeps them all and the seven highpass octants are read back at the ORIGINAL
        size (the reference's x*a / x*b slices, :232-237, 280-288)."""
        n, d0, d1, d2 = X.shape
        parts = {(): X}
        for ax in (3, 2, 1):
            nxt = {}
            for key, A in parts.items():
                nxt[(0,) + key] = _ops.colfilter(A, t["h0o"], ax)
                nxt[(1,) + key] = _ops.colfilter(A, t["h1o"], ax)
            parts = nxt
        lll = parts[(0, 0, 0)]
        yh = _ops.new_highpass(n, 28, (d0 // 2, d1 // 2, d2 // 2), lll.dtype, lll.device)
        for i, o in enumerate(_OCTANTS):
            _ops.cube2c(parts.pop(o)[:, :d0, :d1, :d2].contiguous(), yh, 4 * i)
        return lll, yh

    @staticmethod
    def _merge_even(Yl, yh, t):
        """Inverse of :meth:`_split_even` (reference transform3d.py:385-440 with an even-length filter): every merge reads
        the first n samples of its axis (:408-413) and returns n+1; the first row / column / slice is dropped (:437-438)."""
        n, a0, a1, a2 = (int(s) for s in (yh.shape[0], 2 * yh.shape[2], 2 * yh.shape[3], 2 * yh.shape[4]))
        if tuple(Yl.shape) != (n, a0 + 1, a1 + 1, a2 + 1):
            raise ValueError("lowpass and highpass sizes are not valid for the inverse 3-D transform")
        parts = {(0, 0, 0): Yl[:, :a0, :a1, :a2].contiguous()}
        for i, o in enumerate(_OCTANTS):
            parts[o] = _ops.c2cube(yh, 4 * i)
        for ax in (1, 0, 2):
            nxt = {}
            for key in [k for k in parts if k[ax] == 0]:
                other = tuple(1 if i == ax else key[i] for i in range(3))
                out = _ops.colfilter(parts[key], t["g0o"], ax + 1)
                _ops.colfilter(parts[other], t["g1o"], ax + 1, out=out, accumulate=True)
                nxt[key] = out
            parts = nxt
        return parts[(0, 0, 0)][:, 1:, 1:, 1:].contiguous()

    @staticmethod
    def _split(X, filt):
        """One analysis level: lo/hi along axis 2, then 1, then 0 (tensor axes 3, 2, 1) -> LLL + 28 bands."""
        parts = {(): X}
        for ax in (3, 2, 1):
            nxt = {}
            for key, A in parts.items():
                nxt[(0,) + key] = filt(A, ax, False)
                nxt[(1,) + key] = filt(A, ax, True)
            parts = nxt
        lll = parts[(0, 0, 0)]
        n = lll.shape[0]
        half = tuple(int(s) // 2 for s in lll.shape[1:])
        yh = _ops.new_highpass(n, 28, half, lll.dtype, lll.device)
        for i, o in enumerate(_OCTANTS):
            _ops.cube2c(parts.pop(o), yh, 4 * i)
        return lll, yh

    # ------------------------------------------------------------------ inverse
    def _inverse_n(self, Yl, Yh, t):
        L = len(Yh)
        for lev in range(L - 1, -1, -1):
            if lev == 0:
                fused = None
                if Yh[0] is None:
                    fused = _ops.lowpass3d(Yl, t["g0o"], inverse=True)
                    if fused is None:
                        for ax in (2, 1, 3):     # reference _level1_ifm_no_highpass :442-456, axes 1, 0, 2
                            Yl = _ops.colfilter(Yl, t["g0o"], ax)
                elif t["g0o"].shape[0] % 2 == 0:
                    fused = self._merge_even(Yl, Yh[0], t)
                else:
                    self._check_sizes(Yl, Yh[0])
                    fused = _ops.inv3d_level1(Yl, Yh[0], t["g0o"], t["g1o"])
                if fused is not None:
                    Yl = fused
                elif Yh[0] is not None:
                    Yl = self._merge(Yl, Yh[0], lambda A, ax, hi, out: _ops.colfilter(
                        A, t["g1o"] if hi else t["g0o"], ax, out=out, accumulate=out is not None))
            else:
                cur = [int(s) for s in Yh[lev].shape[2:]]
                prev = [int(s) for s in Yh[lev - 1].shape[2:]] if Yh[lev - 1] is not None else [2 * s for s in cur]
                n = 1 if self.ext_mode == 4 else 2
                crops = {ax + 1: (n if cur[ax] * 2 != prev[ax] else 0) for ax in range(3)}

                def ifilt(A, ax, hi, out, crops=crops):
                    ha, hb = (t["g1b"], t["g1a"]) if hi else (t["g0b"], t["g0a"])
                    return _ops.colifilt(A, ha, hb, ax, crops[ax], out=out, accumulate=out is not None)

                self._check_sizes(Yl, Yh[lev])
                fused = _ops.inv3d_levelq(Yl, Yh[lev], t["g0b"], t["g0a"], t["g1b"], t["g1a"], [crops[ax] for ax in (1, 2, 3)])
                Yl = fused if fused is not None else self._merge(Yl, Yh[lev], ifilt)
        return Yl

    @staticmethod
    def _check_sizes(Yl, yh):
        if tuple(Yl.shape[1:]) != tuple(2 * int(s) for s in yh.shape[2:]) or Yl.shape[0] != yh.shape[0]:
            raise ValueError("lowpass and highpass sizes are not valid for the inverse 3-D transform")

    @staticmethod
    def _merge(Yl, yh, filt):
        """One synthesis level: merge lo/hi pairs along axis 1, then 0, then 2 (tensor axes 2, 1, 3)."""
        Transform3d._check_sizes(Yl, yh)
        parts = {(0, 0, 0): Yl}
        for i, o in enumerate(_OCTANTS):
            parts[o] = _ops.c2cube(yh, 4 * i)
        for ax in (1, 0, 2):
            nxt = {}
            for key in [k for k in parts if k[ax] == 0]:
                other = tuple(1 if i == ax else key[i] for i in range(3))
                out = filt(parts[key], ax + 1, False, None)
                filt(parts[other], ax + 1, True, out)
                nxt[key] = out
            parts = nxt
        return parts[(0, 0, 0)]
