"""Import the UNMODIFIED reference (rjw57/dtcwt) under numpy >= 2.

TEST / MEASUREMENT INFRASTRUCTURE -- only ``tests/``, ``__graft_entry__`` and
``bench.py``'s CPU legs import this; nothing under ``dtcwt_b200/`` does.

Where the reference comes from, first hit wins:
  1. ``$DTCWT_REFERENCE`` (a checkout),
  2. ``oracle/_ref/`` -- the pip ``--target`` install made by ``oracle/build_ref.py``; this is the
     copy that travels to the GPU box (git-ignored, not gpurun-ignored),
  3. ``/root/reference`` (the read-only checkout of the build container).

The reference's files are never edited: three removed numpy attributes are
re-added by monkey-patch before ``import dtcwt`` (they
are used at ``dtcwt/utils.py:105,116-120`` and ``dtcwt/numpy/lowlevel.py:74,209,236``),
and ``load_registration()`` rebinds the two ``dtcwt.registration`` functions that
index with a list / rely on the pre-numpy-2 ``solve`` broadcasting
(``registration.py:242,439,442``) to bodies that differ only in those expressions.
"""
import logging
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _candidates():
    env = os.environ.get("DTCWT_REFERENCE")
    if env:
        yield env
    yield os.path.join(HERE, "_ref")
    yield "/root/reference"


def reference_root():
    for root in _candidates():
        if os.path.isfile(os.path.join(root, "dtcwt", "__init__.py")):
            return root
    return None


REFERENCE_ROOT = reference_root()


def available():
    return reference_root() is not None


def kind():
    """'reference' -- what bench.py reports as cpu_baseline.kind when this module supplies the CPU leg."""
    return "reference"


def load():
    """Return the reference ``dtcwt`` module (shimmed), or raise ImportError."""
    root = reference_root()
    if root is None:
        raise ImportError("reference not found (run `python oracle/build_ref.py` where /root/reference exists)")
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
    if root not in sys.path:
        sys.path.insert(0, root)
    import warnings
    with warnings.catch_warnings():
        # the reference's docstrings hold '\p' style escapes: Python 3.12 reports each as a SyntaxWarning when it compiles them
        warnings.simplefilter("ignore", SyntaxWarning)
        import dtcwt  # noqa: E402
        import dtcwt.numpy  # noqa: F401,E402
    return dtcwt


def load_registration():
    """``dtcwt.registration`` of the reference, usable under numpy >= 2."""
    load()
    import dtcwt.registration as reg
    if getattr(reg, "_b200_shimmed", False):
        return reg

    def _boxfilter(X, kernel_size):
        # registration.py:425-446 with X[tuple(slices)] (list indexing was removed from numpy)
        if kernel_size % 2 == 0:
            raise ValueError('Kernel size must be odd')
        for axis_idx in range(2):
            slices = [slice(None), ] * len(X.shape)
            out = X
            for delta in range(1, 1 + (kernel_size - 1) // 2):
                slices[axis_idx] = reg.dtcwt.utils.reflect(np.arange(X.shape[axis_idx]) + delta, -0.5, X.shape[axis_idx] - 0.5)
                out = out + X[tuple(slices)]
                slices[axis_idx] = reg.dtcwt.utils.reflect(np.arange(X.shape[axis_idx]) - delta, -0.5, X.shape[axis_idx] - 0.5)
                out = out + X[tuple(slices)]
            X = out / kernel_size
        return X

    def solvetransform(Qtilde_vec):
        # registration.py:214-257: only the upper triangle of Q is filled (Q_TRIU_FLAT_INDICES) before the solve;
        # the one change is solve(Q, -q[..., None])[..., 0] -- numpy 2 reads a stacked (..., 6) right-hand side as matrices
        Q = np.zeros(Qtilde_vec.shape[:-1] + (6 * 6,))
        Q[..., reg.Q_TRIU_FLAT_INDICES] = Qtilde_vec[..., :21]
        q = Qtilde_vec[..., -6:]
        Q = np.reshape(Q, Qtilde_vec.shape[:-1] + (6, 6))
        return np.linalg.solve(Q, -q[..., None])[..., 0]

    reg._boxfilter = _boxfilter
    reg.solvetransform = solvetransform
    reg._b200_shimmed = True
    return reg
