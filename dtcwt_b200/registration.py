"""DTCWT-based image registration on the GPU -- drop-in for ``dtcwt.registration``.

Mirrors the reference module (``dtcwt/registration.py``): ``estimatereg`` (:304), ``velocityfield`` (:374),
``warp`` (:410), ``warptransform`` (:275), ``warphighpass`` (:395), ``qtildematrices`` (:141), ``solvetransform``
(:214) with the same arguments and the same level schedule.  Pyramids are :class:`dtcwt_b200.Pyramid` objects (or
anything with ``highpasses``); a leading batch dimension is allowed everywhere, so frame PAIRS shard over GPUs like
every other batch (``dtcwt_b200.parallel``).  The per-pixel work -- confidence, phase gradients, the 27-element
Q~ vectors, box filter + rescale, the 6x6 solves (upper triangle only, as the reference does), velocity fields and
sub-band warping -- runs in the CUDA kernels of ``csrc/registration.cuh`` in float64 arithmetic.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib, _ops, sampling
from .common import Pyramid

__all__ = ["estimatereg", "velocityfield", "warp", "warptransform", "warphighpass", "qtildematrices", "solvetransform",
           "EXPECTED_SHIFTS"]

#: horizontal and vertical expected phase shifts per sub-band (reference registration.py:30)
EXPECTED_SHIFTS = np.array(((-1, -3), (-3, -3), (-3, -1), (-3, 1), (-3, 3), (-1, 3))) * np.pi / 2.15


def _level(pyr, level):
    """Sub-bands of one level as a complex tensor [n][h][w][6] on the device (+ whether a batch dim was added)."""
    hs = pyr.highpasses_t if isinstance(pyr, Pyramid) and "highpasses" not in pyr._np else pyr.highpasses
    h = _ops.as_complex_tensor(hs[level])
    batched = h.dim() == 4
    if not batched:
        h = h.unsqueeze(0)
    if h.dim() != 4 or h.shape[-1] != 6:
        raise ValueError("highpass arrays must be [h][w][6] (or [n][h][w][6])")
    return h, batched


def _strides(h):
    """complex strides (n, band, row, col) of a [n][h][w][6] view (planar storage or interleaved, both fine)"""
    return h.stride(0), h.stride(3), h.stride(1), h.stride(2)


def _qtilde(src, ref, reduce_out=None):
    """Q~ of one level: src, ref [n][h][w][6] complex.  -> [n][h][w][27] float64, or adds the image sums to reduce_out [n][27]."""
    if src.shape != ref.shape:
        raise ValueError("Subbands should have identical size")
    if src.dtype != ref.dtype:
        ref = ref.to(src.dtype)
    n, h, w, _ = src.shape
    suffix = "f32" if src.dtype == torch.complex64 else "f64"
    out = reduce_out if reduce_out is not None else torch.empty((n, h, w, 27), dtype=torch.float64, device=src.device)
    with _ops._on_device(src):
        _lib.call("reg_qtilde", suffix, _ops._ptr(src), _ops._ptr(ref), _ops._ptr(out), n, h, w, *_strides(src), *_strides(ref),
                  int(reduce_out is not None), _ops._stream(src))
    return out


def qtildematrices(t_ref, t_target, levels):
    """Q~ matrices of the given 0-based *levels* (reference registration.py:141-212): a list of ``[h][w][27]`` tensors."""
    out = []
    for level in levels:
        a, batched = _level(t_ref, level)
        b, _ = _level(t_target, level)
        q = _qtilde(a, b)
        out.append(q if batched else q[0])
    return out


def solvetransform(Qtilde_vec):
    """a = -Q^-1 q from 27-element Q~ vectors (reference registration.py:214-257; upper triangle of Q only)."""
    q = Qtilde_vec if isinstance(Qtilde_vec, torch.Tensor) else torch.from_numpy(np.asarray(Qtilde_vec, dtype=np.float64))
    q = _ops.to_device(q).to(torch.float64).contiguous()
    out = torch.empty(tuple(q.shape[:-1]) + (6,), dtype=torch.float64, device=q.device)
    with _ops._on_device(q):
        _lib.call("reg_solve", None, _ops._ptr(q), _ops._ptr(out), int(q.numel() // 27), 0, _ops._stream(q))
    return out


def _avecs_tensor(avecs):
    a = avecs if isinstance(avecs, torch.Tensor) else torch.from_numpy(np.asarray(avecs, dtype=np.float64))
    a = _ops.to_device(a).to(torch.float64).contiguous()
    batched = a.dim() == 4
    return (a if batched else a.unsqueeze(0)), batched


def _coords(avecs4, shape, mode):
    n, H, W, _ = avecs4.shape
    h, w = int(shape[0]), int(shape[1])
    xs = torch.empty((n, h, w), dtype=torch.float64, device=avecs4.device)
    ys = torch.empty_like(xs)
    with _ops._on_device(avecs4):
        _lib.call("reg_coords", None, _ops._ptr(avecs4), _ops._ptr(xs), _ops._ptr(ys), n, H, W, h, w, mode, _ops._stream(avecs4))
    return xs, ys


def velocityfield(avecs, shape, method=None):
    """(vxs, vys) of the affine-parameter grid *avecs* resampled to *shape* (reference registration.py:374-393).
    ``'bilinear'`` runs in one fused kernel; other methods rescale the two fields with :mod:`dtcwt_b200.sampling`."""
    a4, batched = _avecs_tensor(avecs)
    if method == "bilinear":
        vx, vy = _coords(a4, shape, 0)
        return (vx, vy) if batched else (vx[0], vy[0])
    n, H, W, _ = a4.shape
    px = (torch.arange(W, dtype=torch.float64, device=a4.device) / W).view(1, 1, W)
    py = (torch.arange(H, dtype=torch.float64, device=a4.device) / H).view(1, H, 1)
    fx = a4[..., 0] + a4[..., 2] * px + a4[..., 4] * py
    fy = a4[..., 1] + a4[..., 3] * px + a4[..., 5] * py
    vx = torch.stack([sampling.rescale(f, shape, method) for f in fx])
    vy = torch.stack([sampling.rescale(f, shape, method) for f in fy])
    return (vx, vy) if batched else (vx[0], vy[0])


def _sample_coords(avecs4, shape, method):
    """pixel coordinates (X + vx) * w, (Y + vy) * h the warps sample at (reference :401-404, :416-423)"""
    if method == "bilinear":
        return _coords(avecs4, shape, 1)
    vx, vy = velocityfield(avecs4, shape, method)
    h, w = int(shape[0]), int(shape[1])
    X = (torch.arange(w, dtype=torch.float64, device=vx.device) / w).view(1, 1, w)
    Y = (torch.arange(h, dtype=torch.float64, device=vx.device) / h).view(1, h, 1)
    return (X + vx) * w, (Y + vy) * h


def _sample_batch(im4, xs, ys, method, phase):
    """im4 [n][h][w][C] (real or complex) sampled at xs, ys [n][oh][ow] -> [n][oh][ow][C]"""
    n, h, w, C = im4.shape
    oh, ow = xs.shape[1], xs.shape[2]
    cplx = im4.is_complex()
    real = torch.view_as_real(im4) if cplx else im4
    k = 2 if cplx else 1
    suffix = "f32" if real.dtype == torch.float32 else "f64"
    out = torch.empty((n, oh, ow, C), dtype=im4.dtype, device=im4.device)
    if phase:
        kx, px = sampling._dptr(sampling.DTHETA_DX_2D)
        ky, py = sampling._dptr(sampling.DTHETA_DY_2D)
    else:
        px = py = ctypes.POINTER(ctypes.c_double)()
    with _ops._on_device(im4):
        _lib.call("sample", suffix, _ops._ptr(real), _ops._ptr(out), _ops._ptr(xs), _ops._ptr(ys), n, h, w, C, oh, ow,
                  im4.stride(0), im4.stride(1), im4.stride(2), im4.stride(3), oh * ow * C, ow * C, C, 1,
                  oh * ow, int(cplx), sampling._method(method), 0, px, py, _ops._stream(im4))
    return out


def warphighpass(Yh, avecs, method=None):
    """Warp a ``[h][w][6]`` sub-band array by the velocity field of *avecs*, de-rotating first (reference :395-408)."""
    a4, _ = _avecs_tensor(avecs)
    h = _ops.as_complex_tensor(Yh)
    batched = h.dim() == 4
    h4 = h if batched else h.unsqueeze(0)
    if a4.shape[0] != h4.shape[0]:
        a4 = a4.expand(h4.shape[0], -1, -1, -1).contiguous()
    xs, ys = _sample_coords(a4, h4.shape[1:3], method)
    out = _sample_batch(h4, xs, ys, method, True)
    return out if batched else out[0]


def warp(I, avecs, method=None):
    """Warp an image by the velocity field of *avecs* (reference registration.py:410-423)."""
    a4, _ = _avecs_tensor(avecs)
    im = I if isinstance(I, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(I)))
    if im.dtype not in (torch.float32, torch.float64):
        im = im.double()
    im = _ops.to_device(im)
    batched = im.dim() == 3
    im4 = (im if batched else im.unsqueeze(0)).unsqueeze(-1)
    if a4.shape[0] != im4.shape[0]:
        a4 = a4.expand(im4.shape[0], -1, -1, -1).contiguous()
    xs, ys = _sample_coords(a4, im4.shape[1:3], method)
    out = _sample_batch(im4.contiguous(), xs, ys, method, False)[..., 0]
    return out if batched else out[0]


def warptransform(t, avecs, levels, method=None):
    """A copy of pyramid *t* with the given 0-based *levels* warped (reference registration.py:275-302)."""
    hs = list(t.highpasses_t if isinstance(t, Pyramid) and "highpasses" not in t._np else t.highpasses)
    for l in levels:
        hs[l] = warphighpass(hs[l], avecs, method=method)
    lo = t.lowpass_t if isinstance(t, Pyramid) and "lowpass" not in t._np else t.lowpass
    sc = getattr(t, "scales_t", None) if isinstance(t, Pyramid) else getattr(t, "scales", None)
    return Pyramid(lo, tuple(hs), sc)


def estimatereg(source, reference, regshape=None, levels=None):
    """Affine-parameter grid mapping *source* onto *reference* (reference registration.py:304-372): ``[H][W][6]`` float64
    (``[n][H][W][6]`` for batched pyramids), H x W the size of level 4's sub-bands unless *regshape* is given."""
    nlevels = len(source.highpasses_t if isinstance(source, Pyramid) else source.highpasses)
    lvl3, batched = _level(source, 3)
    n = lvl3.shape[0]
    H, W = (lvl3.shape[1], lvl3.shape[2]) if regshape is None else (int(regshape[0]), int(regshape[1]))
    dev = lvl3.device
    if levels is None:                                   # reference :328-335
        levels = [[x for x in range(nlevels - 1, nlevels - 3, -1) if x >= 0]]
        for s in np.arange(nlevels - 1, 0, -0.5):
            refine = [int(np.floor(s)) - x for x in range(2) if s - x >= 2]
            if len(refine) >= 2:
                levels.append(refine)
    # global estimate: Q~ summed over every pixel of the coarse levels (:337-346)
    qsum = torch.zeros((n, 27), dtype=torch.float64, device=dev)
    for l in levels[0]:
        _qtilde(_level(source, l)[0], _level(reference, l)[0], reduce_out=qsum)
    a = solvetransform(qsum)                             # [n][6]
    avecs = a.view(n, 1, 1, 6).expand(n, H, W, 6).contiguous()
    # refinement (:348-370)
    for est_levels in levels[1:]:
        qts = torch.zeros((n, H, W, 27), dtype=torch.float64, device=dev)
        for l in est_levels:
            src, _ = _level(source, l)
            xs, ys = _coords(avecs, src.shape[1:3], 1)
            warped = _sample_batch(src, xs, ys, "bilinear", True)
            q = _qtilde(warped, _level(reference, l)[0])
            with _ops._on_device(q):
                _lib.call("reg_boxrescale", None, _ops._ptr(q), _ops._ptr(qts), n, q.shape[1], q.shape[2], H, W, 1,
                          _ops._stream(q))
        with _ops._on_device(qts):
            _lib.call("reg_solve", None, _ops._ptr(qts), _ops._ptr(avecs), n * H * W, 1, _ops._stream(qts))
    return avecs if batched else avecs[0]
