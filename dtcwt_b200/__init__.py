"""dtcwt_b200 -- a Blackwell (B200, sm_100a) backend for the dual-tree complex wavelet transform.

Drop-in for the hot path of rjw57/dtcwt: ``Transform1d`` / ``Transform2d`` /
``Transform3d`` with ``forward`` / ``inverse`` and the ``Pyramid`` value type,
plus the ``colfilter`` / ``coldfilt`` / ``colifilt`` primitives, all running as
hand-written CUDA kernels behind the C ABI in ``include/dtcwt_b200.h``.

    import dtcwt, dtcwt_b200
    dtcwt_b200.register()            # adds the 'b200' entry to dtcwt's backend table
    dtcwt.push_backend('b200')
    pyramid = dtcwt.Transform2d().forward(image, nlevels=4)

Importing this package never touches the GPU; the first transform call loads
``libdtcwt_b200.so`` and raises ``RuntimeError`` if it (or a CUDA device) is
missing -- there is no CPU fallback.
"""
from .common import Pyramid
from .transform1d import Transform1d
from .transform2d import Transform2d
from .transform3d import Transform3d
from .backend import register, BACKEND_NAME
from . import coeffs, compat, graph, keypoint, lowlevel, registration, sampling
from .compat import (dtwavexfm, dtwaveifm, dtwavexfm2, dtwaveifm2, dtwavexfm2b, dtwaveifm2b, dtwavexfm3, dtwaveifm3)
from .coeffs import biort, qshift
from .lowlevel import colfilter, coldfilt, colifilt

__version__ = "0.1.0"

__all__ = ["Pyramid", "Transform1d", "Transform2d", "Transform3d", "register", "BACKEND_NAME",
           "coeffs", "compat", "graph", "keypoint", "lowlevel", "registration", "sampling", "dtwavexfm", "dtwaveifm", "dtwavexfm2", "dtwaveifm2",
           "dtwavexfm2b", "dtwaveifm2b", "dtwavexfm3", "dtwaveifm3", "biort", "qshift", "colfilter", "coldfilt", "colifilt"]
