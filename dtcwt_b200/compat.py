"""MATLAB-style tuple wrappers on the GPU backend -- drop-in for ``dtcwt.compat``.

Same names, argument order and defaults as the reference module (``dtcwt/compat.py:32-288``); the reference's own
wrappers are hard-wired to its numpy backend (``compat.py:17``), these call the transforms of this package.  The
forward wrappers return what the reference returns -- NumPy arrays (``Yl, Yh`` and, with *include_scale*, ``Yscale``);
the inverse wrappers return NumPy too.  Code that wants to stay on the device should use the Transform classes.
"""
from __future__ import annotations

from .common import Pyramid
from .defaults import DEFAULT_BIORT, DEFAULT_QSHIFT
from .transform1d import Transform1d
from .transform2d import Transform2d
from .transform3d import Transform3d

__all__ = ["dtwavexfm", "dtwaveifm", "dtwavexfm2", "dtwaveifm2", "dtwavexfm2b", "dtwaveifm2b", "dtwavexfm3", "dtwaveifm3"]


def _unpack(res, include_scale):
    if include_scale:
        return res.lowpass, res.highpasses, res.scales
    return res.lowpass, res.highpasses


def dtwavexfm(X, nlevels=3, biort=DEFAULT_BIORT, qshift=DEFAULT_QSHIFT, include_scale=False):
    """1-D forward transform of a column vector or of the columns of a matrix (reference compat.py:32-68)."""
    return _unpack(Transform1d(biort, qshift).forward(X, nlevels, include_scale), include_scale)


def dtwaveifm(Yl, Yh, biort=DEFAULT_BIORT, qshift=DEFAULT_QSHIFT, gain_mask=None):
    """1-D reconstruction (reference compat.py:70-105)."""
    return Transform1d(biort, qshift).inverse(Pyramid(Yl, Yh), gain_mask=gain_mask).cpu().numpy()


def dtwavexfm2(X, nlevels=3, biort=DEFAULT_BIORT, qshift=DEFAULT_QSHIFT, include_scale=False):
    """2-D forward transform (reference compat.py:107-143); the 6 / 12-tuple ``_bp`` families are accepted too."""
    return _unpack(Transform2d(biort, qshift).forward(X, nlevels, include_scale), include_scale)


def dtwaveifm2(Yl, Yh, biort=DEFAULT_BIORT, qshift=DEFAULT_QSHIFT, gain_mask=None):
    """2-D reconstruction (reference compat.py:145-181)."""
    return Transform2d(biort, qshift).inverse(Pyramid(Yl, Yh), gain_mask=gain_mask).cpu().numpy()


# the reference folds the ...b variants into the originals and keeps the names as aliases (compat.py:183-187)
dtwavexfm2b = dtwavexfm2
dtwaveifm2b = dtwaveifm2


def dtwavexfm3(X, nlevels=3, biort=DEFAULT_BIORT, qshift=DEFAULT_QSHIFT, include_scale=False, ext_mode=4,
               discard_level_1=False):
    """3-D forward transform (reference compat.py:189-246)."""
    res = Transform3d(biort, qshift, ext_mode).forward(X, nlevels, include_scale, discard_level_1)
    return _unpack(res, include_scale)


def dtwaveifm3(Yl, Yh, biort=DEFAULT_BIORT, qshift=DEFAULT_QSHIFT, ext_mode=4):
    """3-D reconstruction (reference compat.py:248-288); ``Yh[0]`` may be ``None`` (treated as zero)."""
    return Transform3d(biort, qshift, ext_mode).inverse(Pyramid(Yl, Yh)).cpu().numpy()
