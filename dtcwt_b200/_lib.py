"""ctypes binding of ``libdtcwt_b200.so`` (the C ABI declared in ``include/dtcwt_b200.h``).

There is exactly one compute path: the CUDA library.  If it has not been built,
is not a device build, was built against another version of the header, or no
CUDA device is present, the first call raises ``RuntimeError`` -- there is no CPU
fallback and no switch that selects one.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdtcwt_b200.so")

_LIB = None
ABI_VERSION = 201       # DTCWT_B200_VERSION of include/dtcwt_b200.h this binding was written against

_P, _I, _L, _D = c_void_p, c_int, c_int64, c_double
_TAPS = ctypes.POINTER(c_double)

# name (without _f32/_f64 suffix) -> argtypes; keep in step with include/dtcwt_b200.h
_TYPED = {
    "colfilter": [_P, _P, _L, _L, _L, _I, _I, _TAPS, _I, _I, _P],
    "coldfilt": [_P, _P, _L, _L, _L, _I, _I, _TAPS, _TAPS, _I, _I, _P],
    "colifilt": [_P, _P, _L, _L, _L, _I, _TAPS, _TAPS, _I, _I, _P],
    "q2c": [_P, _P, _L, _L, _L, _L, _L, _L, _L, _I, _I, _P],
    "c2q": [_P, _P, _L, _L, _L, _L, _L, _L, _L, _I, _I, _D, _D, _P],
    "pack1d": [_P, _P, _L, _L, _L, _P],
    "unpack1d": [_P, _P, _L, _L, _L, _D, _P],
    "cube2c": [_P, _P, _L, _L, _L, _L, _L, _L, _L, _L, _L, _I, _P],
    "c2cube": [_P, _P, _L, _L, _L, _L, _L, _L, _L, _L, _L, _I, _P],
    "reg_qtilde": [_P, _P, _P, _L, _L, _L, _L, _L, _L, _L, _L, _L, _L, _L, _I, _P],
    "kp_energy": [_P, _P, _L, _L, _L, _L, _L, _L, _L, _I, _D, _D, _D, _P],
    "sample": [_P, _P, _P, _P, _L, _L, _L, _L, _L, _L, _L, _L, _L, _L, _L, _L, _L, _L, _L, _I, _I, _I, _TAPS, _TAPS, _P],
}
_UNTYPED = {
    "l2_info": [ctypes.POINTER(c_int64)],
    "reg_boxrescale": [_P, _P, _L, _L, _L, _L, _L, _I, _P],
    "reg_solve": [_P, _P, _L, _I, _P],
    "reg_coords": [_P, _P, _P, _L, _L, _L, _L, _L, _I, _P],
    "kp_maxima": [_P, _P, _L, _L, _L, _D, _I, _P],
}
_F32_ONLY = {
    "fwd2d_level1": [_P, _P, _P, _L, _L, _L, _I, _I, _TAPS, _I, _TAPS, _I, _L, _L, _L, _P],
    "fwd2d_levelq": [_P, _P, _P, _L, _L, _L, _I, _I, _TAPS, _TAPS, _TAPS, _TAPS, _I, _L, _L, _L, _P],
    "inv2d_levelq": [_P, _P, _P, _L, _L, _L, _I, _I, _TAPS, _TAPS, _TAPS, _TAPS, _I, _TAPS, _L, _L, _L, _P],
    "inv2d_level1": [_P, _P, _P, _L, _L, _L, _TAPS, _I, _TAPS, _I, _TAPS, _L, _L, _L, _P],
    "fwd2d_level1_hh": [_P, _P, _L, _L, _L, _I, _I, _TAPS, _I, _L, _L, _L, _P],
    "fwd2d_levelq_hh": [_P, _P, _L, _L, _L, _I, _I, _TAPS, _TAPS, _I, _L, _L, _L, _P],
    "inv2d_levelq_hh": [_P, _P, _L, _L, _L, _I, _I, _TAPS, _TAPS, _I, _TAPS, _L, _L, _L, _P],
    "inv2d_level1_hh": [_P, _P, _L, _L, _L, _TAPS, _I, _TAPS, _L, _L, _L, _P],
    "fwd2d_level12": [_P, _P, _P, _P, _P, _L, _L, _L, _I, _I, _TAPS, _I, _TAPS, _I, _TAPS, _TAPS, _TAPS, _TAPS, _I,
                      _L, _L, _L, _L, _L, _L, _L, _I, _P],
    "inv2d_level21": [_P, _P, _P, _P, _P, _L, _L, _L, _I, _I, _TAPS, _TAPS, _TAPS, _TAPS, _I, _TAPS, _TAPS, _I, _TAPS, _I,
                      _TAPS, _L, _L, _L, _L, _L, _L, _L, _I, _P],
    "fwd3d_level1_lo": [_P, _P, _P, _L, _L, _L, _L, _TAPS, _I, _P],
    "inv3d_level1_lo": [_P, _P, _P, _L, _L, _L, _L, _TAPS, _I, _P],
    "fwd3d_level1": [_P, _P, _P, _P, _L, _L, _L, _L, _TAPS, _I, _TAPS, _I, _L, _L, _L, _L, _L, _P],
    "inv3d_level1": [_P, _P, _P, _P, _L, _L, _L, _L, _TAPS, _I, _TAPS, _I, _L, _L, _L, _L, _L, _P],
    "fwd3d_levelq": [_P, _P, _P, _P, _L, _L, _L, _L, _I, _I, _I, _TAPS, _TAPS, _TAPS, _TAPS, _I, _L, _L, _L, _L, _L, _P],
    "inv3d_levelq": [_P, _P, _P, _P, _L, _L, _L, _L, _I, _I, _I, _TAPS, _TAPS, _TAPS, _TAPS, _I, _L, _L, _L, _L, _L, _P],
}

EXPORTS = (["dtcwt_b200_version", "dtcwt_b200_error_string", "dtcwt_b200_is_device_build"]
           + ["dtcwt_b200_%s_%s" % (n, s) for n in _TYPED for s in ("f32", "f64")]
           + ["dtcwt_b200_%s_f32" % n for n in _F32_ONLY] + ["dtcwt_b200_%s" % n for n in _UNTYPED])


def _bind(path, check_version=True):
    lib = ctypes.CDLL(path)
    lib.dtcwt_b200_version.restype = c_int
    lib.dtcwt_b200_version.argtypes = []
    if check_version and lib.dtcwt_b200_version() != ABI_VERSION:
        raise RuntimeError("dtcwt_b200: %s reports ABI version %d, the Python binding expects %d -- rebuild it "
                           "(python build.py --force)" % (path, lib.dtcwt_b200_version(), ABI_VERSION))
    lib.dtcwt_b200_is_device_build.restype = c_int
    lib.dtcwt_b200_is_device_build.argtypes = []
    lib.dtcwt_b200_error_string.restype = c_char_p
    lib.dtcwt_b200_error_string.argtypes = [c_int]
    for name, args in _TYPED.items():
        for suf in ("f32", "f64"):
            fn = getattr(lib, "dtcwt_b200_%s_%s" % (name, suf))
            fn.restype = c_int
            fn.argtypes = args
    for name, args in _F32_ONLY.items():
        fn = getattr(lib, "dtcwt_b200_%s_f32" % name)
        fn.restype = c_int
        fn.argtypes = args
    for name, args in _UNTYPED.items():
        fn = getattr(lib, "dtcwt_b200_%s" % name)
        fn.restype = c_int
        fn.argtypes = args
    return lib


def lib():
    """The loaded CUDA library; raises RuntimeError when it cannot be used."""
    global _LIB
    if _LIB is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                "dtcwt_b200: %s has not been built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `python build.py`); there is no CPU fallback" % LIB_PATH)
        loaded = _bind(LIB_PATH)
        if not loaded.dtcwt_b200_is_device_build():
            raise RuntimeError("dtcwt_b200: %s is not a device build" % LIB_PATH)
        _LIB = loaded
    return _LIB


def check(code):
    if code != 0:
        msg = lib().dtcwt_b200_error_string(code).decode()
        if code < 0:
            raise ValueError(msg)
        raise RuntimeError("CUDA error %d: %s" % (code, msg))


_LAUNCH_HOOK = None


def set_launch_hook(hook):
    """Observer for C-ABI launches: ``hook(symbol, thunk)`` must call ``thunk()`` exactly once.
    Used by bench.py to count launches and bracket them with CUDA events; None removes it."""
    global _LAUNCH_HOOK
    _LAUNCH_HOOK = hook


E_UNSUPPORTED = -2


def call_optional(name, dtype_suffix, *args, launches=1):
    """Like :func:`call`, but a DTCWT_B200_EUNSUPPORTED answer returns False (nothing was launched)
    so the caller can compose the same result from the generic CUDA kernels.  launches: kernel launches the
    entry point makes (the chained entry points make two per chunk), reported to the launch hook."""
    symbol = "dtcwt_b200_%s_%s" % (name, dtype_suffix)
    fn = getattr(lib(), symbol)
    out = []

    def thunk():
        code = fn(*args)
        if code != E_UNSUPPORTED:
            check(code)
        out.append(code)

    if _LAUNCH_HOOK is None:
        thunk()
    elif launches == 1:
        _LAUNCH_HOOK(symbol, thunk)
    else:
        _LAUNCH_HOOK(symbol, thunk, launches)
    return out[0] == 0


def call(name, dtype_suffix, *args):
    symbol = "dtcwt_b200_%s_%s" % (name, dtype_suffix) if dtype_suffix else "dtcwt_b200_%s" % name
    fn = getattr(lib(), symbol)
    if _LAUNCH_HOOK is None:
        check(fn(*args))
    else:
        _LAUNCH_HOOK(symbol, lambda: check(fn(*args)))
