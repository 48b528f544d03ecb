"""CPU oracle for the DT-CWT hot path -- TEST INFRASTRUCTURE ONLY.

This module is a numpy restatement of the algorithm of the reference's numpy
backend.  It exists to CHECK the CUDA path; nothing under ``dtcwt_b200/`` may
import it (only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do).

Parity status: PINNED.  ``tests/test_oracle.py`` checks every function here
against (a) the MATLAB golden summaries of ``tests/verification.npz``
(committed as ``tests/golden/verification_subset.npz``), (b) full-array
outputs of the unmodified reference run in the build container
(``tests/golden/ref_*.npz``, made by ``tests/golden/make_golden.py``) and, when
``/root/reference`` is present, (c) the live reference on random inputs.

Every filter is written from the closed-form index maps
rather than from the reference's extend/convolve/slice pipeline:

    refl(i; r)      = i mod 2r, mirrored to 2r-1-i when >= r   (utils.py:136-153)
    colfilter       Y[i]  = sum_k h[k]  X[refl(i + m-1-k - m//2)]      (lowlevel.py:47-80)
    coldfilt        Ya[i] = sum_j ha[j] X[refl(4i + m   - 2j)]
                    Yb[i] = sum_j hb[j] X[refl(4i + m+1 - 2j)]         (lowlevel.py:82-154)
    colifilt        four output phases per two inputs                  (lowlevel.py:156-260)

Arithmetic follows the reference's conventions: computation in the dtype of
the data (float32 stays float32, taps are rounded to that dtype first,
``lowlevel.py:33``), integers are promoted to float64 (``utils.py:98-105``).
The per-output summation ORDER also follows the reference (tap 0 first; the
two polyphase halves of coldfilt/colifilt summed separately, then added) so
float32 results agree with it to the last few ulps.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "reflect_index", "colfilter", "coldfilt", "colifilt",
    "q2c", "c2q", "c2q1d", "cube2c", "c2cube",
    "Pyramid", "Transform1d", "Transform2d", "Transform3d",
]


# --------------------------------------------------------------------------- helpers
def _asfloat(X):
    """ints -> float64, float32/float64 kept (reference utils.py:98-105)."""
    X = np.asarray(X)
    if X.dtype in (np.float32, np.float64):
        return X
    if np.issubdtype(X.dtype, np.complexfloating):
        return X
    return X.astype(np.float64)


def _complex_of(dtype):
    """float32 -> complex64, everything else complex128 (utils.py:107-124)."""
    return np.complex64 if np.dtype(dtype) == np.float32 else np.complex128


def _taps(h, dtype):
    return np.asarray(h, dtype=np.float64).reshape(-1).astype(dtype)


def reflect_index(i, r):
    """Half-sample symmetric fold of integer indices onto [0, r).

    Equals ``reflect(i, -0.5, r-0.5)`` of the reference (utils.py:136-153):
    ... 1 0 | 0 1 ... r-1 | r-1 r-2 ...  with period 2r.
    """
    i = np.mod(np.asarray(i, dtype=np.int64), 2 * r)
    return np.where(i >= r, 2 * r - 1 - i, i)


def _gather_mac(X, taps, index_rows):
    """sum_k taps[k] * X[index_rows[k]] accumulated in order k = 0, 1, ..."""
    acc = np.zeros((len(index_rows[0]),) + X.shape[1:], dtype=X.dtype)
    for t, idx in zip(taps, index_rows):
        acc += X[idx] * t
    return acc


# --------------------------------------------------------------------------- the three filters
def colfilter(X, h):
    """Undecimated symmetric-extension FIR along axis 0 (lowlevel.py:47-80).

    m odd -> same number of rows; m even -> one more row.
    """
    X = _asfloat(X)
    h = _taps(h, X.dtype)
    r, m = X.shape[0], h.shape[0]
    n_out = r if (m % 2) else r + 1
    i = np.arange(n_out)
    rows = [reflect_index(i + (m - 1 - k) - m // 2, r) for k in range(m)]
    return _gather_mac(X, h, rows)


def _check_dual(X, ha, hb, mult, what):
    if X.shape[0] % mult != 0:
        raise ValueError("No. of rows in X must be a multiple of %d" % mult)
    if np.asarray(ha).shape != np.asarray(hb).shape:
        raise ValueError("Shapes of ha and hb must be the same")
    if np.asarray(ha).shape[0] % 2 != 0:
        raise ValueError("Lengths of ha and hb must be even")


def coldfilt(X, ha, hb):
    """2:1 decimating dual filter along axis 0 (lowlevel.py:82-154)."""
    X = _asfloat(X)
    _check_dual(X, ha, hb, 4, "coldfilt")
    pos = float(np.sum(np.asarray(ha, float) * np.asarray(hb, float))) > 0
    ha = _taps(ha, X.dtype)
    hb = _taps(hb, X.dtype)
    r, m = X.shape[0], ha.shape[0]
    i = np.arange(r // 4)
    base = 4 * i + m
    # polyphase halves: even-numbered taps first, then odd-numbered (lowlevel.py:151-152)
    ya = (_gather_mac(X, ha[0::2], [reflect_index(base - 2 * j, r) for j in range(0, m, 2)]) +
          _gather_mac(X, ha[1::2], [reflect_index(base - 2 * j, r) for j in range(1, m, 2)]))
    yb = (_gather_mac(X, hb[0::2], [reflect_index(base + 1 - 2 * j, r) for j in range(0, m, 2)]) +
          _gather_mac(X, hb[1::2], [reflect_index(base + 1 - 2 * j, r) for j in range(1, m, 2)]))
    Y = np.empty((r // 2,) + X.shape[1:], dtype=X.dtype)
    if pos:
        Y[0::2], Y[1::2] = ya, yb
    else:
        Y[0::2], Y[1::2] = yb, ya
    return Y


def colifilt(X, ha, hb):
    """1:2 interpolating dual filter along axis 0 (lowlevel.py:156-260).

    The reference's all-zero shortcut (lowlevel.py:202) is a no-op for results
    when X is identically zero and wrong otherwise; this
    restatement simply always computes.
    """
    X = _asfloat(X)
    _check_dual(X, ha, hb, 2, "colifilt")
    pos = float(np.sum(np.asarray(ha, float) * np.asarray(hb, float))) > 0
    ha = _taps(ha, X.dtype)
    hb = _taps(hb, X.dtype)
    r, m = X.shape[0], ha.shape[0]
    m2 = m // 2
    i = np.arange(r // 2)
    k = np.arange(m2)
    Y = np.empty((2 * r,) + X.shape[1:], dtype=X.dtype)
    if m2 % 2:  # lowlevel.py:232-258
        ia = [reflect_index(2 * i + m2 - 2 * kk, r) for kk in k]
        ib = [reflect_index(2 * i + m2 - 1 - 2 * kk, r) for kk in k]
        if not pos:
            ia, ib = ib, ia
        Y[0::4] = _gather_mac(X, ha[0::2], ib)
        Y[1::4] = _gather_mac(X, hb[0::2], ia)
        Y[2::4] = _gather_mac(X, ha[1::2], ib)
        Y[3::4] = _gather_mac(X, hb[1::2], ia)
    else:       # lowlevel.py:205-231
        d = (-2, -1, 0, 1) if pos else (-1, -2, 1, 0)
        idx = [[reflect_index(2 * i + m2 - 2 * kk + dd, r) for kk in k] for dd in d]
        Y[0::4] = _gather_mac(X, ha[1::2], idx[0])
        Y[1::4] = _gather_mac(X, hb[1::2], idx[1])
        Y[2::4] = _gather_mac(X, ha[0::2], idx[2])
        Y[3::4] = _gather_mac(X, hb[0::2], idx[3])
    return Y


def _along(fn, X, axis, *taps):
    """Apply an axis-0 filter along ``axis`` of an n-d array."""
    Xm = np.moveaxis(X, axis, 0)
    return np.moveaxis(fn(Xm, *taps), 0, axis)


# --------------------------------------------------------------------------- sub-band packing
def q2c(y):
    """2x2 quads -> two complex sub-bands (transform2d.py:301-322).

    a b / c d  ->  z0 = ((a-d) + j(b+c))/sqrt2,  z1 = ((a+d) + j(b-c))/sqrt2
    """
    y = _asfloat(y)
    s = y.dtype.type(np.sqrt(0.5))
    a, b = y[0::2, 0::2] * s, y[0::2, 1::2] * s
    c, d = y[1::2, 0::2] * s, y[1::2, 1::2] * s
    z = np.empty(a.shape + (2,), dtype=_complex_of(y.dtype))
    z[..., 0].real, z[..., 0].imag = a - d, b + c
    z[..., 1].real, z[..., 1].imag = a + d, b - c
    return z


def c2q(w, gain):
    """Inverse of q2c with one gain per sub-band (transform2d.py:324-350)."""
    w = np.asarray(w)
    rdt = w.real.dtype
    sc = np.sqrt(0.5) * np.asarray(gain, dtype=np.float64)
    w0, w1 = w[..., 0] * sc[0], w[..., 1] * sc[1]
    P, Q = w0 + w1, w0 - w1
    x = np.empty((2 * w.shape[0], 2 * w.shape[1]), dtype=rdt)
    x[0::2, 0::2] = P.real
    x[0::2, 1::2] = P.imag
    x[1::2, 0::2] = Q.imag
    x[1::2, 1::2] = -Q.real
    return x


def c2q1d(x):
    """complex (n, c) -> real (2n, c): even rows real, odd rows imag (transform1d.py:186-196)."""
    x = np.asarray(x)
    z = np.empty((2 * x.shape[0],) + x.shape[1:], dtype=x.real.dtype)
    z[0::2], z[1::2] = x.real, x.imag
    return z


def cube2c(y):
    """2x2x2 octets -> four complex sub-bands (transform3d.py:532-579)."""
    y = _asfloat(y)
    A, B = y[0::2, 0::2, 0::2], y[0::2, 1::2, 0::2]
    C, D = y[1::2, 0::2, 0::2], y[1::2, 1::2, 0::2]
    E, F = y[0::2, 0::2, 1::2], y[0::2, 1::2, 1::2]
    G, H = y[1::2, 0::2, 1::2], y[1::2, 1::2, 1::2]
    z = np.empty(A.shape + (4,), dtype=_complex_of(y.dtype))
    h = y.dtype.type(0.5)
    z[..., 0].real, z[..., 0].imag = (A - G - D - F) * h, (B - H + C + E) * h
    z[..., 1].real, z[..., 1].imag = (A - G + D + F) * h, (-B + H + C + E) * h
    z[..., 2].real, z[..., 2].imag = (A + G + D - F) * h, (B + H - C + E) * h
    z[..., 3].real, z[..., 3].imag = (A + G - D + F) * h, (-B - H - C + E) * h
    return z


def c2cube(z):
    """Inverse of cube2c (transform3d.py:581-619)."""
    z = np.asarray(z)
    pr, pi = z[..., 0].real, z[..., 0].imag
    qr, qi = z[..., 1].real, z[..., 1].imag
    rr, ri = z[..., 2].real, z[..., 2].imag
    sr, si = z[..., 3].real, z[..., 3].imag
    y = np.empty(tuple(2 * n for n in z.shape[:3]), dtype=z.real.dtype)
    h = z.real.dtype.type(0.5)
    y[0::2, 0::2, 0::2] = (pr + qr + rr + sr) * h   # A
    y[1::2, 0::2, 1::2] = (-pr - qr + rr + sr) * h  # G
    y[1::2, 1::2, 0::2] = (-pr + qr + rr - sr) * h  # D
    y[0::2, 1::2, 1::2] = (-pr + qr - rr + sr) * h  # F
    y[0::2, 1::2, 0::2] = (pi - qi + ri - si) * h   # B
    y[1::2, 1::2, 1::2] = (-pi + qi + ri - si) * h  # H
    y[1::2, 0::2, 0::2] = (pi + qi - ri - si) * h   # C
    y[0::2, 0::2, 1::2] = (pi + qi + ri + si) * h   # E
    return y


# --------------------------------------------------------------------------- pyramid + transforms
class Pyramid(object):
    """lowpass / highpasses / scales value type (numpy/common.py:5-32)."""

    def __init__(self, lowpass, highpasses, scales=None):
        self.lowpass = _asfloat(lowpass)
        self.highpasses = tuple(None if h is None else np.asarray(h) for h in highpasses)
        self.scales = None if scales is None else tuple(_asfloat(s) for s in scales)


def _split_biort(biort):
    if len(biort) == 4:
        h0o, g0o, h1o, g1o = biort
        return h0o, g0o, h1o, g1o, None, None
    if len(biort) == 6:
        return tuple(biort)
    raise ValueError("Biort wavelet must have 6 or 4 components.")


def _split_qshift(qshift):
    if len(qshift) == 8:
        return tuple(qshift) + (None,) * 4
    if len(qshift) == 12:
        return tuple(qshift)
    raise ValueError("Qshift wavelet must have 12 or 8 components.")


def _edge_pad(X, axis, n=1):
    """n replicated samples on EACH side of ``axis`` (transform2d.py:134-140)."""
    first = np.take(X, [0] * n, axis=axis)
    last = np.take(X, [-1] * n, axis=axis)
    return np.concatenate((first, X, last), axis=axis)


class Transform2d(object):
    """2-D DT-CWT (numpy/transform2d.py:15-295); taps are given as tuples."""

    def __init__(self, biort, qshift):
        self.biort = biort
        self.qshift = qshift

    def forward(self, X, nlevels=3, include_scale=False):
        h0o, _, h1o, _, h2o, _ = _split_biort(self.biort)
        q = _split_qshift(self.qshift)
        h0a, h0b, h1a, h1b, h2a, h2b = q[0], q[1], q[4], q[5], q[8], q[9]
        X = np.atleast_2d(_asfloat(X))
        if X.ndim >= 3:
            raise ValueError("2-D transform needs a 2-D array")
        # odd sizes: repeat last row / column (transform2d.py:86-94)
        if X.shape[0] % 2:
            X = np.concatenate((X, X[-1:, :]), axis=0)
        if X.shape[1] % 2:
            X = np.concatenate((X, X[:, -1:]), axis=1)
        if nlevels == 0:
            return Pyramid(X, (), ()) if include_scale else Pyramid(X, ())
        Yh, Ysc = [], []
        cdt = _complex_of(X.dtype)

        def bands(hl, lh, hh):
            out = np.empty((hl.shape[0] // 2, hl.shape[1] // 2, 6), dtype=cdt)
            out[:, :, [0, 5]] = q2c(hl)   # vertical highpass x horizontal lowpass
            out[:, :, [2, 3]] = q2c(lh)
            out[:, :, [1, 4]] = q2c(hh)
            return out

        # level 1 (transform2d.py:112-130): filter axis 0 first, then axis 1
        Lo = colfilter(X, h0o)
        Hi = colfilter(X, h1o)
        LoLo = _along(colfilter, Lo, 1, h0o)
        if h2o is not None:
            Ba = colfilter(X, h2o)
            hh = _along(colfilter, Ba, 1, h2o)
        else:
            hh = _along(colfilter, Hi, 1, h1o)
        Yh.append(bands(_along(colfilter, Hi, 1, h0o), _along(colfilter, Lo, 1, h1o), hh))
        Ysc.append(LoLo)
        for _ in range(1, nlevels):  # transform2d.py:132-160
            if LoLo.shape[0] % 4:
                LoLo = _edge_pad(LoLo, 0)
            if LoLo.shape[1] % 4:
                LoLo = _edge_pad(LoLo, 1)
            Lo = coldfilt(LoLo, h0b, h0a)
            Hi = coldfilt(LoLo, h1b, h1a)
            if h2a is not None:
                Ba = coldfilt(LoLo, h2b, h2a)
                hh = _along(coldfilt, Ba, 1, h2b, h2a)
            else:
                hh = _along(coldfilt, Hi, 1, h1b, h1a)
            LoLo = _along(coldfilt, Lo, 1, h0b, h0a)
            Yh.append(bands(_along(coldfilt, Hi, 1, h0b, h0a), _along(coldfilt, Lo, 1, h1b, h1a), hh))
            Ysc.append(LoLo)
        return Pyramid(LoLo, tuple(Yh), tuple(Ysc)) if include_scale else Pyramid(LoLo, tuple(Yh))

    def inverse(self, pyramid, gain_mask=None):
        _, g0o, _, g1o, _, g2o = _split_biort(self.biort)
        q = _split_qshift(self.qshift)
        g0a, g0b, g1a, g1b, g2a, g2b = q[2], q[3], q[6], q[7], q[10], q[11]
        Z, Yh = pyramid.lowpass, pyramid.highpasses
        L = len(Yh)
        gm = np.ones((6, L)) if gain_mask is None else np.array(gain_mask)
        for lev in range(L, 1, -1):  # transform2d.py:240-273
            w = Yh[lev - 1]
            lh = c2q(w[:, :, [0, 5]], gm[[0, 5], lev - 1])
            hl = c2q(w[:, :, [2, 3]], gm[[2, 3], lev - 1])
            hh = c2q(w[:, :, [1, 4]], gm[[1, 4], lev - 1])
            y1 = colifilt(Z, g0b, g0a) + colifilt(lh, g1b, g1a)
            if g2a is not None:
                y2 = colifilt(hl, g0b, g0a)
                y3 = colifilt(hh, g2b, g2a)
                Z = (_along(colifilt, y1, 1, g0b, g0a) + _along(colifilt, y2, 1, g1b, g1a) +
                     _along(colifilt, y3, 1, g2b, g2a))
            else:
                y2 = colifilt(hl, g0b, g0a) + colifilt(hh, g1b, g1a)
                Z = _along(colifilt, y1, 1, g0b, g0a) + _along(colifilt, y2, 1, g1b, g1a)
            S = 2 * np.array(Yh[lev - 2].shape[:2])
            if Z.shape[0] != S[0]:
                Z = Z[1:-1, :]
            if Z.shape[1] != S[1]:
                Z = Z[:, 1:-1]
            if np.any(np.array(Z.shape) != S):
                raise ValueError("Sizes of highpasses are not valid for DTWAVEIFM2")
        if L >= 1:  # transform2d.py:275-293
            w = Yh[0]
            lh = c2q(w[:, :, [0, 5]], gm[[0, 5], 0])
            hl = c2q(w[:, :, [2, 3]], gm[[2, 3], 0])
            hh = c2q(w[:, :, [1, 4]], gm[[1, 4], 0])
            y1 = colfilter(Z, g0o) + colfilter(lh, g1o)
            if g2o is not None:
                y2 = colfilter(hl, g0o)
                y3 = colfilter(hh, g2o)
                Z = (_along(colfilter, y1, 1, g0o) + _along(colfilter, y2, 1, g1o) +
                     _along(colfilter, y3, 1, g2o))
            else:
                y2 = colfilter(hl, g0o) + colfilter(hh, g1o)
                Z = _along(colfilter, y1, 1, g0o) + _along(colfilter, y2, 1, g1o)
        return Z


class Transform1d(object):
    """1-D DT-CWT on the columns of (n,) / (n, c) (numpy/transform1d.py:14-184)."""

    def __init__(self, biort, qshift):
        self.biort = biort
        self.qshift = qshift

    def forward(self, X, nlevels=3, include_scale=False):
        h0o, _, h1o, _ = self.biort[:4]
        h0a, h0b, _, _, h1a, h1b, _, _ = self.qshift[:8]
        X = _asfloat(X)
        if X.ndim == 1:
            X = X[:, None]
        if X.shape[0] % 2:
            raise ValueError("Size of input X must be a multiple of 2")
        if nlevels == 0:
            return Pyramid(X, (), ()) if include_scale else Pyramid(X, ())
        Yh, Ysc = [], []
        Hi, Lo = colfilter(X, h1o), colfilter(X, h0o)
        Yh.append(Hi[0::2] + 1j * Hi[1::2])
        Ysc.append(Lo)
        for _ in range(1, nlevels):
            if Lo.shape[0] % 4:
                Lo = _edge_pad(Lo, 0)
            Hi = coldfilt(Lo, h1b, h1a)
            Lo = coldfilt(Lo, h0b, h0a)
            Yh.append(Hi[0::2] + 1j * Hi[1::2])
            Ysc.append(Lo)
        cdt = _complex_of(X.dtype)
        Yh = tuple(y.astype(cdt) for y in Yh)
        return Pyramid(Lo, Yh, tuple(Ysc)) if include_scale else Pyramid(Lo, Yh)

    def inverse(self, pyramid, gain_mask=None):
        _, g0o, _, g1o = self.biort[:4]
        _, _, g0a, g0b, _, _, g1a, g1b = self.qshift[:8]
        Lo, Yh = pyramid.lowpass, pyramid.highpasses
        L = len(Yh)
        gm = np.ones(L) if gain_mask is None else np.asarray(gain_mask)
        if L == 0:
            return Lo
        rdt = Lo.dtype
        for lev in range(L - 1, 0, -1):
            Hi = c2q1d(Yh[lev] * gm[lev]).astype(rdt)
            Lo = colifilt(Lo, g0b, g0a) + colifilt(Hi, g1b, g1a)
            if Lo.shape[0] != 2 * Yh[lev - 1].shape[0]:
                Lo = Lo[1:-1]
            if Lo.shape[0] != 2 * Yh[lev - 1].shape[0] or Lo.shape[1:] != Yh[lev - 1].shape[1:]:
                raise ValueError("Yh sizes are not valid for DTWAVEIFM")
        Hi = c2q1d(Yh[0] * gm[0]).astype(rdt)
        Z = colfilter(Lo, g0o) + colfilter(Hi, g1o)
        return Z.reshape(-1) if Z.shape[1] == 1 else Z


_OCTANTS = ((0, 1, 0), (1, 0, 0), (1, 1, 0), (0, 0, 1), (0, 1, 1), (1, 0, 1), (1, 1, 1))
"""(axis0, axis1, axis2) filter type (0 = lowpass, 1 = highpass) of the seven 3-D
sub-band groups in output order HLL LHL HHL LLH HLH LHH HHH (transform3d.py:280-288)."""


class Transform3d(object):
    """3-D DT-CWT (numpy/transform3d.py:15-526).

    The reference works in an octant "work cube" with Python loops over 2-D
    slices; this restatement filters whole arrays along one axis at a time, in
    the reference's axis order (2, then 1, then 0 forward; 1, 0, then 2
    inverse), which gives the same values.
    """

    def __init__(self, biort, qshift, ext_mode=4):
        self.biort = biort
        self.qshift = qshift
        self.ext_mode = ext_mode

    # ---- forward
    def forward(self, X, nlevels=3, include_scale=False, discard_level_1=False):
        h0o, _, h1o, _ = self.biort[:4]
        h0a, h0b, _, _, h1a, h1b, _, _ = self.qshift[:8]
        if self.ext_mode not in (4, 8):
            raise ValueError("ext_mode must be one of 4 or 8")
        Yl = np.atleast_3d(_asfloat(X))
        Yh, Ysc = [None] * nlevels, [None] * nlevels
        for lev in range(nlevels):
            if lev == 0:
                mult = 2 if self.ext_mode == 4 else 4
                if any(n % mult for n in Yl.shape):
                    raise ValueError("Input shape should be a multiple of %d in each direction" % mult)
                if discard_level_1:
                    for ax in (2, 1, 0):
                        Yl = _along(colfilter, Yl, ax, h0o)
                else:
                    Yl, Yh[0] = self._fwd_level(Yl, colfilter, (h0o,), (h1o,), level1=True)
            else:
                pad = 1 if self.ext_mode == 4 else 2
                for ax in range(3):
                    if Yl.shape[ax] % (4 * pad):
                        Yl = _edge_pad(Yl, ax, pad)
                Yl, Yh[lev] = self._fwd_level(Yl, coldfilt, (h0b, h0a), (h1b, h1a), level1=False)
            Ysc[lev] = Yl.copy()
        return Pyramid(Yl, tuple(Yh), tuple(Ysc)) if include_scale else Pyramid(Yl, tuple(Yh))

    @staticmethod
    def _fwd_level(X, filt, lo, hi, level1):
        def both(A, ax):
            return _along(filt, A, ax, *lo), _along(filt, A, ax, *hi)

        even = level1 and (np.asarray(lo[0]).size % 2 == 0)
        parts = {(): X}
        for ax in (2, 1, 0):
            nxt = {}
            for key, A in parts.items():
                l, h = both(A, ax)
                nxt[(0,) + key] = l
                nxt[(1,) + key] = h
            parts = nxt
        # keys are now (axis0, axis1, axis2) filter types
        Yl = parts[(0, 0, 0)]
        if even:
            # even-length level-1 taps (e.g. Haar, tests/test_xfm3.py:42-58) give n+1 samples per axis: the reference's
            # work cube has one extra row per octant (transform3d.py:223-224); LLL keeps all n+1, the seven highpass
            # octants are read back through the x*a / x*b slices of the ORIGINAL size n (:232-237, :280-288)
            n0, n1, n2 = X.shape
            Yh = np.concatenate([cube2c(parts[o][:n0, :n1, :n2]) for o in _OCTANTS], axis=3)
        else:
            Yh = np.concatenate([cube2c(parts[o]) for o in _OCTANTS], axis=3)
        return Yl, Yh

    # ---- inverse
    def inverse(self, pyramid):
        _, g0o, _, g1o = self.biort[:4]
        _, _, g0a, g0b, _, _, g1a, g1b = self.qshift[:8]
        Yl, Yh = pyramid.lowpass, pyramid.highpasses
        L = len(Yh)
        for lev in range(L - 1, -1, -1):
            if lev == 0:
                if Yh[0] is None:
                    # NOTE the reference writes the axis-2 result back transposed
                    # (transform3d.py:452-454): for cubes it returns this array with axes 0
                    # and 2 swapped, for n0 != n2 it raises.  The oracle (and the CUDA path)
                    # return the untransposed reconstruction; see DESIGN.md "reference quirks".
                    for ax in (1, 0, 2):
                        Yl = _along(colfilter, Yl, ax, g0o)
                else:
                    Yl = self._inv_level(Yl, Yh[0], colfilter, (g0o,), (g1o,))
            else:
                Z = self._inv_level(Yl, Yh[lev], colifilt, (g0b, g0a), (g1b, g1a))
                prev = (np.array(Yh[lev - 1].shape[:3]) if Yh[lev - 1] is not None
                        else np.array(Yh[lev].shape[:3]) * 2)
                cur = np.array(Yh[lev].shape[:3])
                c = 1 if self.ext_mode == 4 else 2
                sl = tuple(slice(c, -c) if cur[ax] * 2 != prev[ax] else slice(None) for ax in range(3))
                Yl = Z[sl]
        return Yl

    @staticmethod
    def _inv_level(Yl, Yh, filt, lo, hi):
        parts = {(0, 0, 0): Yl}
        for n, o in enumerate(_OCTANTS):
            parts[o] = c2cube(Yh[..., 4 * n:4 * n + 4])
        even = filt is colfilter and (np.asarray(lo[0]).size % 2 == 0)
        if even:
            # transform3d.py:385-440 with an even-length filter: every merge reads only the first n samples of an axis
            # (the x*a / x*b slices, :408-413, 429, 431, 436), each merge returns n+1 of them, and the first row / column /
            # slice of the result is dropped (:437-438)
            n0, n1, n2 = parts[_OCTANTS[0]].shape
            parts[(0, 0, 0)] = Yl[:n0, :n1, :n2]
            for ax in (1, 0, 2):
                nxt = {}
                for key, A in parts.items():
                    if key[ax] == 0:
                        other = tuple(1 if i == ax else key[i] for i in range(3))
                        nxt[key] = _along(filt, A, ax, *lo) + _along(filt, parts[other], ax, *hi)
                parts = nxt
            return parts[(0, 0, 0)][1:, 1:, 1:]
        # merge axis 1, then axis 0, then axis 2 (transform3d.py:485-495); a merged
        # axis keeps key 0, so after three rounds only (0, 0, 0) is left
        for ax in (1, 0, 2):
            nxt = {}
            for key, A in parts.items():
                if key[ax] == 0:
                    other = tuple(1 if i == ax else key[i] for i in range(3))
                    nxt[key] = _along(filt, A, ax, *lo) + _along(filt, parts[other], ax, *hi)
            parts = nxt
        return parts[(0, 0, 0)]
