"""2-D DT-CWT on the GPU -- drop-in for ``dtcwt.numpy.Transform2d``.

Mirrors the reference class (``dtcwt/numpy/transform2d.py:15-295``): same
constructor, ``forward(X, nlevels=3, include_scale=False)`` and
``inverse(pyramid, gain_mask=None)``, same padding rules (odd sizes repeat the
last row/column, :86-94; a level whose input is not a multiple of 4 is
edge-extended by one sample per side, :134-140, and cropped again by the
inverse, :263-271), same exceptions.

Batches: ``forward_channels`` / ``inverse_channels`` (names from the reference's
TensorFlow backend, ``dtcwt/tf/transform2d.py:179,422``) take ``[N][H][W]``.
Every image of a batch is independent, so a batch is also the unit that is
sharded across GPUs (see ``dtcwt_b200.parallel``).
"""
from __future__ import annotations

import logging

import numpy as np
import torch

from . import _ops
from .coeffs import biort as _biort, qshift as _qshift
from .common import Pyramid, pyramid_parts
from .defaults import DEFAULT_BIORT, DEFAULT_QSHIFT

__all__ = ["Transform2d"]

# (band0, band1) written by q2c of the three highpass images, reference transform2d.py:122-127:
#   vertical-highpass x horizontal-lowpass -> bands 0,5 ; lowpass x highpass -> 2,3 ; high x high -> 1,4
_BANDS_HL, _BANDS_LH, _BANDS_HH = (0, 5), (2, 3), (1, 4)


def _vec(h):
    return np.array(h, dtype=np.float64).reshape(-1)      # a private copy: it is frozen and cached


class Transform2d(object):
    """*biort* / *qshift* are family names (``dtcwt_b200.coeffs``) or tuples of tap vectors:
    ``(h0o, g0o, h1o, g1o[, h2o, g2o])`` and ``(h0a, h0b, g0a, g0b, h1a, h1b, g1a, g1b[, h2a, h2b, g2a, g2b])``.
    """

    def __init__(self, biort=DEFAULT_BIORT, qshift=DEFAULT_QSHIFT):
        try:
            self.biort = _biort(biort)
        except TypeError:
            self.biort = biort
        try:
            self.qshift = _qshift(qshift)
        except TypeError:
            self.qshift = qshift

    # ------------------------------------------------------------------ tap bookkeeping
    def _taps(self):
        """Tap vectors by name; memoised on the CONTENT of `biort` / `qshift` (they stay plain public attributes, as in
        the reference, so a user may replace or edit them in place) and frozen, so that the launch wrappers can cache
        their ctypes views."""
        key = tuple(np.asarray(h, dtype=np.float64).tobytes() for h in tuple(self.biort) + tuple(self.qshift))
        cached = getattr(self, "_taps_cache", None)
        if cached is not None and cached[0] == key:
            return cached[3]
        t = self._taps_uncached()
        for v in t.values():
            if v is not None:
                v.setflags(write=False)
        self._taps_cache = (key, self.biort, self.qshift, t)
        return t

    def _taps_uncached(self):
        if len(self.biort) == 4:
            h0o, g0o, h1o, g1o = self.biort
            h2o = g2o = None
        elif len(self.biort) == 6:
            h0o, g0o, h1o, g1o, h2o, g2o = self.biort
        else:
            raise ValueError("Biort wavelet must have 6 or 4 components.")
        if len(self.qshift) == 8:
            h0a, h0b, g0a, g0b, h1a, h1b, g1a, g1b = self.qshift
            h2a = h2b = g2a = g2b = None
        elif len(self.qshift) == 12:
            h0a, h0b, g0a, g0b, h1a, h1b, g1a, g1b, h2a, h2b, g2a, g2b = self.qshift
        else:
            raise ValueError("Qshift wavelet must have 12 or 8 components.")
        v = lambda h: None if h is None else _vec(h)  # noqa: E731
        return dict(h0o=v(h0o), g0o=v(g0o), h1o=v(h1o), g1o=v(g1o), h2o=v(h2o), g2o=v(g2o),
                    h0a=v(h0a), h0b=v(h0b), g0a=v(g0a), g0b=v(g0b), h1a=v(h1a), h1b=v(h1b),
                    g1a=v(g1a), g1b=v(g1b), h2a=v(h2a), h2b=v(h2b), g2a=v(g2a), g2b=v(g2b))

    # ------------------------------------------------------------------ public API
    def forward(self, X, nlevels=3, include_scale=False):
        """*nlevels*-level transform of one 2-D image; returns a :class:`Pyramid`."""
        t = self._taps()
        X = _ops.as_real_tensor(X)
        if X.dim() >= 3:
            raise ValueError("The entered image is {0}, which is invalid for the 2D transform. "
                             "Use forward_channels for a batch.".format("x".join(str(s) for s in X.shape)))
        while X.dim() < 2:
            X = X.unsqueeze(0)
        return self._unbatch(self._forward_n(X.unsqueeze(0), t, nlevels, include_scale))

    def inverse(self, pyramid, gain_mask=None):
        """Reconstruct from a Pyramid-like object (ours, the reference's, or any object with
        ``lowpass`` / ``highpasses``); *gain_mask* is ``(6, nlevels)`` (reference :214-217)."""
        t = self._taps()
        Yl, Yh = self._pyramid_tensors(pyramid)
        batched = Yl.dim() == 3
        if not batched:
            Yl = Yl.unsqueeze(0)
        Z = self._inverse_n(Yl, Yh, t, gain_mask)
        return Z if batched else Z[0]

    # data_format strings of the reference's TensorFlow backend (dtcwt/tf/transform2d.py:179-330):
    #   "nhw" / "chw"  X [N][H][W]      -> lowpass [N][h][w],      highpasses [N][h][w][6]
    #   "hwn" / "hwc"  X [H][W][N]      -> lowpass [h][w][N],      highpasses [h][w][N][6]
    #   "nchw"         X [N][C][H][W]   -> lowpass [N][C][h][w],   highpasses [N][C][h][w][6]
    #   "nhwc"         X [N][H][W][C]   -> lowpass [N][h][w][C],   highpasses [N][h][w][C][6]
    # The kernels always see [M][H][W] (M = N or N*C) and planar [M][6][h][w] sub-bands; everything returned is a view.
    _FORMATS_3D = ("nhw", "chw", "hwn", "hwc")
    _FORMATS_4D = ("nchw", "nhwc")

    @classmethod
    def _check_format(cls, data_format, ndim):
        data_format = data_format.lower()
        if data_format not in cls._FORMATS_3D + cls._FORMATS_4D:
            raise ValueError("The data format must be one of: {}".format(cls._FORMATS_3D + cls._FORMATS_4D))
        if ndim != (3 if data_format in cls._FORMATS_3D else 4):
            raise ValueError("The entered variable has incorrect shape for the specified data_format %s." % data_format)
        return data_format

    @staticmethod
    def _to_batch(X, data_format):
        """user layout -> ([M][H][W] contiguous, (N, C) or None)"""
        if data_format in ("nhw", "chw"):
            return X.contiguous(), None
        if data_format in ("hwn", "hwc"):
            return X.permute(2, 0, 1).contiguous(), None
        if data_format == "nchw":
            n, c = X.shape[:2]
            return X.reshape((n * c,) + tuple(X.shape[2:])).contiguous(), (n, c)
        n, c = X.shape[0], X.shape[3]                             # nhwc
        return X.permute(0, 3, 1, 2).reshape((n * c,) + tuple(X.shape[1:3])).contiguous(), (n, c)

    @staticmethod
    def _from_batch(A, data_format, nc, trailing=0):
        """[M][h][w](+[6]) -> user layout (a view); trailing = number of axes after (h, w)"""
        tail = tuple(range(3, 3 + trailing))
        if data_format in ("nhw", "chw"):
            return A
        if data_format in ("hwn", "hwc"):
            return A.permute((1, 2, 0) + tail)
        A = A.unflatten(0, nc)                                    # [N][C][h][w](+[6])
        if data_format == "nchw":
            return A
        return A.permute((0, 2, 3, 1) + tuple(t + 1 for t in tail))   # nhwc

    @staticmethod
    def _to_batch_like(A, data_format, trailing=0):
        """inverse of _from_batch for pyramids handed back by the user: -> [M][h][w](+[6])"""
        tail = tuple(range(3, 3 + trailing))
        if data_format in ("nhw", "chw"):
            return A
        if data_format in ("hwn", "hwc"):
            return A.permute((2, 0, 1) + tail)
        if data_format == "nhwc":
            A = A.permute((0, 3, 1, 2) + tuple(t + 1 for t in tail))
        return A.reshape((-1,) + tuple(A.shape[2:]))

    def forward_channels(self, X, data_format="nhw", nlevels=3, include_scale=False):
        """Batched forward transform; *data_format* as in the reference's TensorFlow backend (see above)."""
        t = self._taps()
        X = _ops.as_real_tensor(X)
        data_format = self._check_format(data_format, X.dim())
        Xb, nc = self._to_batch(X, data_format)
        p = self._forward_n(Xb, t, nlevels, include_scale)
        if data_format in ("nhw", "chw"):
            return p
        lo = self._from_batch(p.lowpass_t, data_format, nc)
        hp = tuple(self._from_batch(h, data_format, nc, 1) for h in p.highpasses_t)
        sc = None if p.scales_t is None else tuple(self._from_batch(s_, data_format, nc) for s_ in p.scales_t)
        return Pyramid(lo, hp, sc)

    def inverse_channels(self, pyramid, data_format="nhw", gain_mask=None):
        """Batched inverse of :meth:`forward_channels` (same *data_format*); returns the images in that layout."""
        t = self._taps()
        lo, hs = pyramid_parts(pyramid)
        lo = _ops.as_real_tensor(lo, "lowpass")
        data_format = self._check_format(data_format, lo.dim())
        nc = (lo.shape[0], lo.shape[1] if data_format == "nchw" else lo.shape[3]) if data_format in self._FORMATS_4D else None
        Yl = self._to_batch_like(lo, data_format).contiguous()
        planar = []
        for h in hs:
            h = _ops.as_complex_tensor(h, Yl.dtype)
            if h.shape[-1] != 6:
                raise ValueError("highpass arrays must have 6 sub-bands on their last axis")
            h = self._to_batch_like(h, data_format, 1)             # [M][h][w][6]
            planar.append(h.permute(0, 3, 1, 2).contiguous())      # no copy when the storage is already planar
        Z = self._inverse_n(Yl, planar, t, gain_mask)
        return self._from_batch(Z, data_format, nc)

    # ------------------------------------------------------------------ pyramid plumbing
    @staticmethod
    def _unbatch(p):
        hp = tuple(h[0] for h in p.highpasses_t)
        sc = None if p.scales_t is None else tuple(s[0] for s in p.scales_t)
        return Pyramid(p.lowpass_t[0], hp, sc)

    @staticmethod
    def _pyramid_tensors(pyramid, batch_dims=None):
        """-> (lowpass [N..][h][w] real, [planar highpass [M][6][h][w] complex per level])."""
        lo, hs = pyramid_parts(pyramid)
        lo = _ops.as_real_tensor(lo, "lowpass")
        planar = []
        for h in hs:
            h = _ops.as_complex_tensor(h, lo.dtype)
            if h.shape[-1] != 6:
                raise ValueError("highpass arrays must have 6 sub-bands on their last axis")
            if h.dim() == 3:
                h = h.unsqueeze(0)
            h = h.reshape((-1,) + tuple(h.shape[-3:]))          # [M][h][w][6]
            planar.append(h.permute(0, 3, 1, 2).contiguous())   # no copy when already planar
        return lo, planar

    # ------------------------------------------------------------------ forward
    def _forward_n(self, X, t, nlevels, include_scale):
        N, H, W = X.shape
        if H < 1 or W < 1:
            raise ValueError("empty image")
        ph, pw = H % 2, W % 2          # odd size: repeat last row / column (reference :86-94)
        if nlevels == 0:
            if ph or pw:
                X = torch.nn.functional.pad(X.unsqueeze(1), (0, pw, 0, ph), mode="replicate").squeeze(1)
            return Pyramid(X, (), ()) if include_scale else Pyramid(X, ())
        if t["h0o"].shape[0] % 2 == 0 or t["h1o"].shape[0] % 2 == 0:
            raise ValueError("even-length biorthogonal filters are not supported by the 2-D transform")
        Yh, Ysc = [], []
        chained = None
        if nlevels >= 2 and not include_scale and t["h2o"] is None and t["h2a"] is None:
            # levels 1 and 2 chunk by chunk with LoLo1 kept in L2 (same kernels, same results; _ops.chain_mode)
            chained = _ops.fwd2d_level12(X, t["h0o"], t["h1o"], t["h0b"], t["h0a"], t["h1b"], t["h1a"], (ph, pw),
                                         internal2="lolo2" if nlevels > 2 else None)
        if chained is not None:
            LoLo, yh1, yh2 = chained
            Yh += [yh1, yh2]
            Ysc += [None, LoLo]
        else:
            # a LoLo that is only the next level's input (not the pyramid's lowpass, not a requested scale) is scratch
            LoLo, yh = self._fwd_level1(X, t, ph, pw, internal=(nlevels > 1 and not include_scale))
            Yh.append(yh)
            Ysc.append(LoLo)
        for lev in range(len(Yh), nlevels):
            LoLo, yh = self._fwd_levelq(LoLo, t, internal=("lolo%d" % (lev + 1)) if (lev < nlevels - 1 and not include_scale) else None)
            Yh.append(yh)
            Ysc.append(LoLo)
        if ph or pw:
            logging.warning("The image entered is now a %dx%d NOT a %dx%d.", H + ph, W + pw, H, W)
            logging.warning("The %s been duplicated, prior to decomposition.",
                            {(1, 1): "bottom row and rightmost column have", (1, 0): "bottom row has",
                             (0, 1): "rightmost column has"}[(ph, pw)])
        views = tuple(h.permute(0, 2, 3, 1) for h in Yh)         # [N][h][w][6] views of planar storage
        return Pyramid(LoLo, views, tuple(Ysc)) if include_scale else Pyramid(LoLo, views)

    def _fwd_level1(self, X, t, ph, pw, internal=False):
        """Level 1 (reference :112-130): undecimated biort filters, vertical axis first."""
        N = X.shape[0]
        fused = _ops.fwd2d_level1(X, t["h0o"], t["h1o"], (ph, pw), internal)
        if fused is not None:
            # `_bp` families (reference :116-121): a second launch overwrites the diagonal sub-bands with the band-pass ones
            if t["h2o"] is None or _ops.fwd2d_level1_hh(X, fused[1], t["h2o"], (ph, pw)):
                return fused
        Lo = _ops.colfilter(X, t["h0o"], 1, (0, ph))
        Hi = _ops.colfilter(X, t["h1o"], 1, (0, ph))
        LoLo = _ops.colfilter(Lo, t["h0o"], 2, (0, pw))
        yh = _ops.new_highpass(N, 6, (LoLo.shape[1] // 2, LoLo.shape[2] // 2), X.dtype, X.device)
        _ops.q2c(_ops.colfilter(Hi, t["h0o"], 2, (0, pw)), yh, *_BANDS_HL)
        _ops.q2c(_ops.colfilter(Lo, t["h1o"], 2, (0, pw)), yh, *_BANDS_LH)
        if t["h2o"] is not None:
            Ba = _ops.colfilter(X, t["h2o"], 1, (0, ph))
            _ops.q2c(_ops.colfilter(Ba, t["h2o"], 2, (0, pw)), yh, *_BANDS_HH)
        else:
            _ops.q2c(_ops.colfilter(Hi, t["h1o"], 2, (0, pw)), yh, *_BANDS_HH)
        return LoLo, yh

    def _fwd_levelq(self, LoLo, t, internal=None):
        """Level >= 2 (reference :132-160): decimating q-shift pairs."""
        N, r, c = LoLo.shape
        pr = (1, 1) if r % 4 else (0, 0)
        pc = (1, 1) if c % 4 else (0, 0)
        fused = _ops.fwd2d_levelq(LoLo, t["h0b"], t["h0a"], t["h1b"], t["h1a"], (pr[0], pc[0]), internal)
        if fused is not None:
            if t["h2a"] is None or _ops.fwd2d_levelq_hh(LoLo, fused[1], t["h2b"], t["h2a"], (pr[0], pc[0])):       # reference :145-157
                return fused
        Lo = _ops.coldfilt(LoLo, t["h0b"], t["h0a"], 1, pr)
        Hi = _ops.coldfilt(LoLo, t["h1b"], t["h1a"], 1, pr)
        out = _ops.coldfilt(Lo, t["h0b"], t["h0a"], 2, pc)
        yh = _ops.new_highpass(N, 6, (out.shape[1] // 2, out.shape[2] // 2), LoLo.dtype, LoLo.device)
        _ops.q2c(_ops.coldfilt(Hi, t["h0b"], t["h0a"], 2, pc), yh, *_BANDS_HL)
        _ops.q2c(_ops.coldfilt(Lo, t["h1b"], t["h1a"], 2, pc), yh, *_BANDS_LH)
        if t["h2a"] is not None:
            Ba = _ops.coldfilt(LoLo, t["h2b"], t["h2a"], 1, pr)
            _ops.q2c(_ops.coldfilt(Ba, t["h2b"], t["h2a"], 2, pc), yh, *_BANDS_HH)
        else:
            _ops.q2c(_ops.coldfilt(Hi, t["h1b"], t["h1a"], 2, pc), yh, *_BANDS_HH)
        return out, yh

    # ------------------------------------------------------------------ inverse
    def _inverse_n(self, Z, Yh, t, gain_mask):
        a = len(Yh)
        gm = np.ones((6, a)) if gain_mask is None else np.array(gain_mask, dtype=np.float64)
        if gm.shape != (6, a):
            raise ValueError("gain_mask must have shape (6, %d)" % a)
        N = Z.shape[0]
        for h in Yh:
            if h.shape[0] != N:
                raise ValueError("lowpass and highpass batch sizes differ")
        for lev in range(a, 1, -1):                      # reference :240-273
            want = (2 * Yh[lev - 2].shape[2], 2 * Yh[lev - 2].shape[3])
            if lev == 2 and t["g2a"] is None and t["g2o"] is None:
                # levels 2 and 1 chunk by chunk with the level-1 lowpass kept in L2 (_ops.chain_mode)
                self._check_lowpass(Z, Yh[1])
                out = _ops.inv2d_level21(Z, Yh[1], Yh[0], t["g0b"], t["g0a"], t["g1b"], t["g1a"], gm[:, 1], t["g0o"], t["g1o"],
                                         gm[:, 0], self._crops(Z, want))
                if out is not None:
                    return out
            Z = self._inv_levelq(Z, Yh[lev - 1], t, gm[:, lev - 1], want)
        if a >= 1:                                        # reference :275-293
            Z = self._inv_level1(Z, Yh[0], t, gm[:, 0])
        return Z

    @staticmethod
    def _crops(Z, want):
        crops = []
        for have, need in zip((2 * Z.shape[1], 2 * Z.shape[2]), want):
            if have == need:
                crops.append(0)
            elif have - 2 == need:
                crops.append(1)      # this level's input had been edge-extended (reference :263-268)
            else:
                raise ValueError("Sizes of highpasses are not valid for DTWAVEIFM2")
        return tuple(crops)

    @staticmethod
    def _check_lowpass(Z, yh):
        if (Z.shape[1], Z.shape[2]) != (2 * yh.shape[2], 2 * yh.shape[3]):
            raise ValueError("Sizes of highpasses are not valid for DTWAVEIFM2")

    def _inv_levelq(self, Z, yh, t, g, want):
        self._check_lowpass(Z, yh)
        cr, cc = self._crops(Z, want)
        if t["g2a"] is None:
            fused = _ops.inv2d_levelq(Z, yh, t["g0b"], t["g0a"], t["g1b"], t["g1a"], g, (cr, cc))
            if fused is not None:
                return fused
        else:
            # `_bp` (reference :254-262): the ordinary launch without the diagonal sub-bands, then out += their g2 contribution
            g_rest = np.array(g, dtype=np.float64)
            g_rest[[1, 4]] = 0.0
            fused = _ops.inv2d_levelq(Z, yh, t["g0b"], t["g0a"], t["g1b"], t["g1a"], g_rest, (cr, cc))
            if fused is not None and _ops.inv2d_levelq_hh(yh, fused, Z.shape[1], Z.shape[2], t["g2b"], t["g2a"], g, (cr, cc)):
                return fused
        lh = _ops.c2q(yh, _BANDS_HL[0], _BANDS_HL[1], g[0], g[5])
        hl = _ops.c2q(yh, _BANDS_LH[0], _BANDS_LH[1], g[2], g[3])
        hh = _ops.c2q(yh, _BANDS_HH[0], _BANDS_HH[1], g[1], g[4])
        y1 = _ops.colifilt(Z, t["g0b"], t["g0a"], 1, cr)
        _ops.colifilt(lh, t["g1b"], t["g1a"], 1, cr, out=y1, accumulate=True)
        y2 = _ops.colifilt(hl, t["g0b"], t["g0a"], 1, cr)
        if t["g2a"] is not None:
            y3 = _ops.colifilt(hh, t["g2b"], t["g2a"], 1, cr)
        else:
            _ops.colifilt(hh, t["g1b"], t["g1a"], 1, cr, out=y2, accumulate=True)
        out = _ops.colifilt(y1, t["g0b"], t["g0a"], 2, cc)
        _ops.colifilt(y2, t["g1b"], t["g1a"], 2, cc, out=out, accumulate=True)
        if t["g2a"] is not None:
            _ops.colifilt(y3, t["g2b"], t["g2a"], 2, cc, out=out, accumulate=True)
        return out

    def _inv_level1(self, Z, yh, t, g):
        self._check_lowpass(Z, yh)
        if t["g2o"] is None:
            fused = _ops.inv2d_level1(Z, yh, t["g0o"], t["g1o"], g)
            if fused is not None:
                return fused
        else:
            g_rest = np.array(g, dtype=np.float64)
            g_rest[[1, 4]] = 0.0
            fused = _ops.inv2d_level1(Z, yh, t["g0o"], t["g1o"], g_rest)
            if fused is not None and _ops.inv2d_level1_hh(yh, fused, t["g2o"], g):          # reference :279-292
                return fused
        lh = _ops.c2q(yh, _BANDS_HL[0], _BANDS_HL[1], g[0], g[5])
        hl = _ops.c2q(yh, _BANDS_LH[0], _BANDS_LH[1], g[2], g[3])
        hh = _ops.c2q(yh, _BANDS_HH[0], _BANDS_HH[1], g[1], g[4])
        y1 = _ops.colfilter(Z, t["g0o"], 1)
        _ops.colfilter(lh, t["g1o"], 1, out=y1, accumulate=True)
        y2 = _ops.colfilter(hl, t["g0o"], 1)
        if t["g2o"] is not None:
            y3 = _ops.colfilter(hh, t["g2o"], 1)
        else:
            _ops.colfilter(hh, t["g1o"], 1, out=y2, accumulate=True)
        out = _ops.colfilter(y1, t["g0o"], 2)
        _ops.colfilter(y2, t["g1o"], 2, out=out, accumulate=True)
        if t["g2o"] is not None:
            _ops.colfilter(y3, t["g2o"], 2, out=out, accumulate=True)
        return out
