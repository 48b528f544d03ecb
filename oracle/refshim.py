"""Import the UNMODIFIED reference (rjw57/dtcwt) under numpy >= 2 -- build-container only.

TEST INFRASTRUCTURE.  The reference checkout (``/root/reference`` or
``$DTCWT_REFERENCE``) does not exist on the GPU box; this helper is used only
by ``tests/golden/make_golden.py`` and by CPU tests that skip when it is
absent.  The reference's files are never edited: three removed numpy
attributes are re-added by monkey-patch before ``import dtcwt``
(they are used at ``dtcwt/utils.py:105,116-120`` and
``dtcwt/numpy/lowlevel.py:74,209,236``).
"""
import logging
import os
import sys

import numpy as np

REFERENCE_ROOT = os.environ.get("DTCWT_REFERENCE", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "dtcwt", "__init__.py"))


def load():
    """Return the reference ``dtcwt`` module (shimmed), or raise ImportError."""
    if not available():
        raise ImportError("reference checkout not found at %s" % REFERENCE_ROOT)
    if not hasattr(np, "asfarray"):
        def _asfarray(a, dtype=np.float64):
            dtype = np.dtype(dtype)
            if not np.issubdtype(dtype, np.inexact):
                dtype = np.dtype(np.float64)
            return np.asarray(a, dtype=dtype)
        np.asfarray = _asfarray
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "issubsctype"):
        np.issubsctype = np.issubdtype
    if not hasattr(logging, "warn"):
        logging.warn = logging.warning
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import dtcwt  # noqa: E402
    import dtcwt.numpy  # noqa: F401,E402
    return dtcwt
