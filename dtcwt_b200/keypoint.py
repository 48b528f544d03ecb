"""Keypoint detection from DT-CWT sub-bands on the GPU -- drop-in for ``dtcwt.keypoint``.

Mirrors ``find_keypoints`` of the reference (``dtcwt/keypoint.py:9-141``): same arguments, same three energy methods
(``'fauqueur'`` default, ``'bendale'``, ``'kingsbury'``), same scale / position conventions, the same ``(P, 4)`` result
(x, y, scale, energy) sorted by decreasing energy.  The per-pixel work -- energy maps, 3x3 maxima, the quadratic
sub-pixel refinement -- runs in the CUDA kernels of ``csrc/keypoint.cuh`` (float64); selecting and sorting the
surviving points is done with torch on the device.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, _ops, sampling

__all__ = ["find_keypoints"]

_METHODS = {"fauqueur": 0, "bendale": 1, "kingsbury": 2}


def _energy(subband, method, alpha, beta, kappa, scale):
    h = _ops.as_complex_tensor(subband)
    if h.dim() != 3 or h.shape[-1] != 6:
        raise ValueError("highpass arrays must be [h][w][6]")
    e = torch.empty((1, h.shape[0], h.shape[1]), dtype=torch.float64, device=h.device)
    suffix = "f32" if h.dtype == torch.complex64 else "f64"
    with _ops._on_device(h):
        _lib.call("kp_energy", suffix, _ops._ptr(h), _ops._ptr(e), 1, h.shape[0], h.shape[1], 0, h.stride(2), h.stride(0),
                  h.stride(1), _METHODS[method], float(alpha) ** (scale + 1), float(beta), float(kappa), _ops._stream(h))
    return e[0]


def _maxima(X, threshold, refine):
    """-> (rows, cols, values) tensors of the kept local maxima of the energy map X (reference :201-260)"""
    X = X.contiguous()
    if threshold is None:
        threshold = float(X.min()) - 1
    out = torch.empty(tuple(X.shape) + (4,), dtype=torch.float64, device=X.device)
    with _ops._on_device(X):
        _lib.call("kp_maxima", None, _ops._ptr(X), _ops._ptr(out), 1, X.shape[0], X.shape[1], float(threshold), int(bool(refine)),
                  _ops._stream(X))
    keep = out[..., 0] > 0
    sel = out[keep]                      # row-major order, like numpy.nonzero
    return sel[:, 1], sel[:, 2], sel[:, 3]


def find_keypoints(highpass_highpasses, method=None, alpha=1.0, beta=0.4, kappa=1.0 / 6.0, threshold=None, max_points=None,
                   upsample_keypoint_energy=None, upsample_highpasses=None, refine_positions=True, skip_levels=1):
    """Keypoints of a pyramid's highpass tuple; see the reference docstring (``dtcwt/keypoint.py:9-80``).  Returns a
    ``(P, 4)`` float64 tensor: x, y, scale, energy."""
    if method is None:
        method = "fauqueur"
    if method == "gale":
        raise NotImplementedError("not implemented yet")
    if method not in _METHODS:
        raise ValueError("Unknown method: {0}".format(method))
    levels = list(highpass_highpasses)[skip_levels:]
    upsample_scale = 1
    if upsample_highpasses is not None:
        upsample_scale <<= 1
    if upsample_keypoint_energy is not None:
        upsample_scale <<= 1
    energies = []
    for scale, subband in enumerate(levels):
        if upsample_highpasses is not None:
            subband = sampling.upsample_highpass(subband, upsample_highpasses)
        e = _energy(subband, method, alpha, beta, kappa, scale)
        if upsample_keypoint_energy is not None:
            e = sampling.upsample(e, upsample_keypoint_energy)
        energies.append(e)
    parts = []
    for level_idx, e in enumerate(energies):
        kp_scale = 2 ** (level_idx + 1 + skip_levels) / float(upsample_scale)
        rows, cols, vals = _maxima(e, threshold, refine_positions)
        parts.append(torch.stack(((cols + 0.5) * kp_scale - 0.5, (rows + 0.5) * kp_scale - 0.5,
                                  torch.full_like(cols, kp_scale), vals), dim=1))
    if not parts:
        return torch.zeros((0, 4), dtype=torch.float64)
    kps = torch.cat(parts, dim=0)
    kps = kps[torch.argsort(kps[:, 3], descending=True, stable=True)]
    if max_points is not None:
        kps = kps[:max_points]
    return kps
