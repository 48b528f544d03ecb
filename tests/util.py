"""Shared helpers for the test-suite (fixtures, summaries, error measures)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

# tolerance stated by BASELINE.json north_star: 1e-5 relative to the reference, fp32
REL_TOL = 1e-5
# tolerance the reference uses against the MATLAB summaries (tests/test_againstmatlab.py:38)
MATLAB_ABS_TOL = 1e-5

_cache = {}


def golden(name):
    if name not in _cache:
        with np.load(os.path.join(GOLDEN, name + ".npz")) as d:
            _cache[name] = {k: d[k] for k in d.files}
    return _cache[name]


def rel_err(a, ref):
    """max|a - ref| / max|ref|  (per-array relative infinity-norm error)."""
    a, ref = np.asarray(a), np.asarray(ref)
    assert a.shape == ref.shape, (a.shape, ref.shape)
    scale = float(np.abs(ref).max())
    d = float(np.abs(a - ref).max())
    return d / scale if scale > 0 else d


def _kmean(a, axis):
    return np.expand_dims(np.mean(a, axis=axis), axis)


def summarise_mat(M, apron=8):
    """(2*apron+1)^2 summary: the four apron x apron corners verbatim, the four
    edge strips averaged along their long side, the centre averaged to one value.
    Restated from the reference's tests/util.py:46-60 (also matlab/verif_m_to_npz.py)."""
    a = apron
    top, mid, bot = M[:a], M[a:-a], M[-a:]

    def row(block, reduce_rows):
        left, centre, right = block[:, :a], block[:, a:-a], block[:, -a:]
        if reduce_rows:
            left, centre, right = _kmean(left, 0), _kmean(centre, 0), _kmean(right, 0)
        return np.concatenate((left, _kmean(centre, 1), right), axis=1)

    return np.concatenate((row(top, False), row(mid, True), row(bot, False)), axis=0)


def summarise_cube(M, apron=4):
    """summarise_mat applied to every axis-2 slice, stacked on axis 2 (tests/util.py:62-67)."""
    parts = [summarise_mat(M[:, :, i, ...], apron) for i in range(M.shape[2])]
    return np.dstack(parts)


def reg_frames(shape=(192, 256), seed=5):
    """A seeded synthetic frame pair for the registration tests: band-limited texture and a copy of it moved by a
    smooth, mostly translational flow of about (2, 3) pixels (the reference's examples/register_images.py rolls a
    frame; a smooth flow also exercises the affine terms).  float64 in [0, 1]."""
    rs = np.random.RandomState(seed)
    h, w = shape
    F = np.fft.rfft2(rs.randn(h, w))
    fy = np.fft.fftfreq(h)[:, None]
    fx = np.fft.rfftfreq(w)[None, :]
    F *= np.exp(-(fx ** 2 + fy ** 2) / (2 * 0.06 ** 2))
    tex = np.fft.irfft2(F, s=(h, w))
    tex = (tex - tex.min()) / (tex.max() - tex.min())
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    sx = xx - 3.0 - 1.5 * (yy / h - 0.5)
    sy = yy - 2.0 + 1.0 * (xx / w - 0.5)
    x0 = np.clip(np.floor(sx).astype(int), 0, w - 2)
    y0 = np.clip(np.floor(sy).astype(int), 0, h - 2)
    ax = np.clip(sx - x0, 0, 1)
    ay = np.clip(sy - y0, 0, 1)
    moved = ((1 - ay) * ((1 - ax) * tex[y0, x0] + ax * tex[y0, x0 + 1]) +
             ay * ((1 - ax) * tex[y0 + 1, x0] + ax * tex[y0 + 1, x0 + 1]))
    return tex, moved
