"""``colfilter`` / ``coldfilt`` / ``colifilt`` on the GPU.

Same names, argument meaning and error behaviour as the reference's
``dtcwt/numpy/lowlevel.py`` (``colfilter`` :47, ``coldfilt`` :82, ``colifilt``
:156): the filter runs along axis 0 of a 2-D array.  Inputs may be anything
``numpy.asarray`` accepts or a ``torch.Tensor``; the result is a ``torch.Tensor``
on the CUDA device.  Extension: ``axis`` selects the filtered axis of an n-d
tensor so that callers never need to transpose.
"""
from __future__ import annotations

import numpy as np

from . import _ops

__all__ = ["colfilter", "coldfilt", "colifilt"]


def _vec(h):
    return np.asarray(h, dtype=np.float64).reshape(-1)


def _check_pair(ha, hb):
    if np.asarray(ha).shape != np.asarray(hb).shape:
        raise ValueError("Shapes of ha and hb must be the same")
    if _vec(ha).shape[0] % 2 != 0:
        raise ValueError("Lengths of ha and hb must be even")


def colfilter(X, h, axis=0):
    """Filter along ``axis`` with ``h``, symmetric extension, no decimation.

    Odd-length ``h``: output shape == input shape; even length: one more sample
    along ``axis`` (reference lowlevel.py:49-52).
    """
    X = _ops.as_real_tensor(X)
    if X.dim() < 1:
        raise ValueError("X must have at least one dimension")
    return _ops.colfilter(X, _vec(h), axis % X.dim())


def coldfilt(X, ha, hb, axis=0):
    """2:1 decimating dual-tree filter pair along ``axis`` (reference lowlevel.py:82-154).

    Raises ValueError if the axis length is not a multiple of 4 or the taps are
    not two equal, even-length vectors (lowlevel.py:118-125).
    """
    X = _ops.as_real_tensor(X)
    axis = axis % X.dim()
    if X.shape[axis] % 4 != 0:
        raise ValueError("No. of rows in X must be a multiple of 4")
    _check_pair(ha, hb)
    return _ops.coldfilt(X, _vec(ha), _vec(hb), axis)


def colifilt(X, ha, hb, axis=0):
    """1:2 interpolating dual-tree filter pair along ``axis`` (reference lowlevel.py:156-260).

    Raises ValueError if the axis length is odd or the taps are not two equal,
    even-length vectors (lowlevel.py:189-196).
    """
    X = _ops.as_real_tensor(X)
    axis = axis % X.dim()
    if X.shape[axis] % 2 != 0:
        raise ValueError("No. of rows in X must be a multiple of 2")
    _check_pair(ha, hb)
    return _ops.colifilt(X, _vec(ha), _vec(hb), axis)
