"""Thin tensor-level wrappers over the C ABI: shapes in, launches out.

Every function takes C-contiguous ``torch.Tensor`` arguments that already live on
the compute device, describes them to the library as ``[outer][len][inner]``
views (see ``include/dtcwt_b200.h``) and returns freshly allocated outputs.
PyTorch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

from . import _lib

_SUFFIX = {torch.float32: "f32", torch.float64: "f64"}
_COMPLEX = {torch.float32: torch.complex64, torch.float64: torch.complex128}
_REAL = {torch.complex64: torch.float32, torch.complex128: torch.float64}


# ----------------------------------------------------------------------------- coercion
def as_real_tensor(X, name="X"):
    """Array-like / tensor -> contiguous float32/float64 tensor on the compute device.

    dtype rules follow the reference's ``asfarray`` (utils.py:98-105): float32 and
    float64 are kept, everything else becomes float64 (float16/bfloat16 -> float32).
    """
    if not isinstance(X, torch.Tensor):
        X = np.asarray(X)
        if X.dtype not in (np.float32, np.float64):
            if np.issubdtype(X.dtype, np.complexfloating):
                raise ValueError("%s must be real" % name)
            X = X.astype(np.float64)
        X = torch.from_numpy(np.ascontiguousarray(X))
    if X.is_complex():
        raise ValueError("%s must be real" % name)
    if X.dtype in (torch.float16, torch.bfloat16):
        X = X.float()
    elif X.dtype not in (torch.float32, torch.float64):
        X = X.double()
    return to_device(X).contiguous()


def as_complex_tensor(Z, real_dtype=None):
    if not isinstance(Z, torch.Tensor):
        Z = np.asarray(Z)
        if not np.issubdtype(Z.dtype, np.complexfloating):
            Z = Z.astype(np.complex64 if Z.dtype == np.float32 else np.complex128)
        Z = torch.from_numpy(Z)   # may be non-contiguous; callers re-layout
    if not Z.is_complex():
        Z = Z.to(_COMPLEX.get(Z.dtype, torch.complex128))
    if real_dtype is not None and _REAL[Z.dtype] != real_dtype:
        Z = Z.to(_COMPLEX[real_dtype])
    return to_device(Z)


def to_device(t):
    if t.device.type == "cuda":
        return t
    if not torch.cuda.is_available():
        raise RuntimeError("dtcwt_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return t.cuda(non_blocking=False)


def complex_dtype(real_dtype):
    return _COMPLEX[real_dtype]


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)      # the handle without building a Stream object


def _stream(t):
    if t.device.type == "cuda":
        if _raw_stream is not None:
            return ctypes.c_void_p(_raw_stream(t.device.index if t.device.index is not None else torch.cuda.current_device()))
        return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)
    return ctypes.c_void_p(0)


class _on_device(object):
    """Make the tensor's device current for the duration of a launch."""

    def __init__(self, t):
        # entering a torch.cuda.device context costs ~10 us; skip it when the tensor's device is already current
        self.ctx = None
        if t.device.type == "cuda" and torch.cuda.current_device() != t.device.index:
            self.ctx = torch.cuda.device(t.device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)


_TAP_CACHE = {}     # id(array) -> (array kept alive, contiguous float64 copy, ctypes pointer, length)


def _taps(h):
    """Host tap vector -> (float64 array, ctypes pointer, length).  The transforms hand the same ndarray objects to
    every call, so the conversion is memoised per object (it was a third of the host time of a small transform)."""
    hit = _TAP_CACHE.get(id(h))
    if hit is not None and hit[0] is h:
        return hit[1], hit[2], hit[3]
    k = np.ascontiguousarray(np.asarray(h, dtype=np.float64).reshape(-1))
    out = (k, k.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), int(k.shape[0]))
    if isinstance(h, np.ndarray) and not h.flags.writeable:
        if len(_TAP_CACHE) > 256:
            _TAP_CACHE.clear()
        _TAP_CACHE[id(h)] = (h,) + out
    return out


_SCRATCH = {}       # (tag, shape, dtype, device, stream) -> tensor


def scratch(tag, shape, dtype, device):
    """Reusable device scratch for buffers that never leave a transform call (the four per-slice images of a 3-D
    level, LoLo intermediates of a 2-D forward without include_scale): one buffer per (purpose, shape, stream), so a
    steady-state loop of transforms allocates nothing.  Work on one stream is ordered, so reusing the buffer in the next
    call on that stream is safe; another stream gets its own."""
    if device.type != "cuda":
        stream = 0
    elif _raw_stream is not None:
        stream = _raw_stream(device.index if device.index is not None else torch.cuda.current_device())
    else:
        stream = torch.cuda.current_stream(device).cuda_stream
    key = (tag, tuple(int(s) for s in shape), dtype, str(device), stream)
    t = _SCRATCH.get(key)
    if t is None:
        if len(_SCRATCH) > 64:
            _SCRATCH.clear()
        t = torch.empty(key[1], dtype=dtype, device=device)
        _SCRATCH[key] = t
    return t


def release_scratch():
    """Drop the cached scratch buffers (they are otherwise kept for the life of the process)."""
    _SCRATCH.clear()


def _view(shape, axis):
    outer = int(np.prod(shape[:axis], dtype=np.int64))
    inner = int(np.prod(shape[axis + 1:], dtype=np.int64))
    return outer, int(shape[axis]), inner


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


# ----------------------------------------------------------------------------- filters
def colfilter(x, h, axis, pad=(0, 0), out=None, accumulate=False):
    hk, hp, m = _taps(h)
    outer, n, inner = _view(x.shape, axis)
    L = n + pad[0] + pad[1]
    shape = list(x.shape)
    shape[axis] = L if m % 2 else L + 1
    y = _out(x, shape, out, accumulate)
    with _on_device(x):
        _lib.call("colfilter", _SUFFIX[x.dtype], _ptr(x), _ptr(y), outer, n, inner, pad[0], pad[1],
                  hp, m, int(accumulate), _stream(x))
    return y


def coldfilt(x, ha, hb, axis, pad=(0, 0), out=None, accumulate=False):
    ka, pa, m = _taps(ha)
    kb, pb, mb = _taps(hb)
    outer, n, inner = _view(x.shape, axis)
    L = n + pad[0] + pad[1]
    shape = list(x.shape)
    shape[axis] = L // 2
    y = _out(x, shape, out, accumulate)
    with _on_device(x):
        _lib.call("coldfilt", _SUFFIX[x.dtype], _ptr(x), _ptr(y), outer, n, inner, pad[0], pad[1],
                  pa, pb, m, int(accumulate), _stream(x))
    return y


def colifilt(x, ha, hb, axis, crop=0, out=None, accumulate=False):
    ka, pa, m = _taps(ha)
    kb, pb, mb = _taps(hb)
    outer, n, inner = _view(x.shape, axis)
    shape = list(x.shape)
    shape[axis] = 2 * n - 2 * crop
    y = _out(x, shape, out, accumulate)
    with _on_device(x):
        _lib.call("colifilt", _SUFFIX[x.dtype], _ptr(x), _ptr(y), outer, n, inner, crop,
                  pa, pb, m, int(accumulate), _stream(x))
    return y


def _out(x, shape, out, accumulate):
    if out is None:
        if accumulate:
            raise ValueError("accumulate needs an output tensor")
        return torch.empty(shape, dtype=x.dtype, device=x.device)
    if list(out.shape) != list(shape) or out.dtype != x.dtype or not out.is_contiguous():
        raise ValueError("output tensor has the wrong shape/dtype: %s vs %s" % (list(out.shape), list(shape)))
    return out


# ----------------------------------------------------------------------------- 2-D packing
def new_highpass(n, bands, spatial, real_dtype, device):
    """Planar sub-band storage [n][bands][*spatial] (complex)."""
    return torch.empty((n, bands) + tuple(spatial), dtype=_COMPLEX[real_dtype], device=device)


def q2c(y, z, band0, band1):
    """y real [n][2h][2w] -> bands band0/band1 of planar z [n][6][h][w]."""
    n, h2, w2 = y.shape
    h, w = h2 // 2, w2 // 2
    assert z.shape[0] == n and tuple(z.shape[2:]) == (h, w) and z.is_contiguous()
    with _on_device(y):
        _lib.call("q2c", _SUFFIX[y.dtype], _ptr(y), _ptr(z), n, h, w,
                  z.stride(0), z.stride(1), z.stride(2), z.stride(3), band0, band1, _stream(y))


def c2q(z, band0, band1, gain0, gain1):
    """bands band0/band1 of planar z [n][6][h][w] -> real [n][2h][2w]."""
    n, _, h, w = z.shape
    y = torch.empty((n, 2 * h, 2 * w), dtype=_REAL[z.dtype], device=z.device)
    with _on_device(z):
        _lib.call("c2q", _SUFFIX[y.dtype], _ptr(z), _ptr(y), n, h, w,
                  z.stride(0), z.stride(1), z.stride(2), z.stride(3), band0, band1,
                  float(gain0), float(gain1), _stream(z))
    return y


# ----------------------------------------------------------------------------- fused 2-D levels
FUSED_ENABLED = True      # tests flip this to compare the fused kernels with the generic composition


def _fused_ok(*tensors):
    return FUSED_ENABLED and all(t.dtype in (torch.float32, torch.complex64) for t in tensors)


def fwd2d_level1(x, h0o, h1o, pad_hi, internal=False):
    """Fused level 1 of the 2-D forward transform: x [n][H][W] -> (LoLo [n][H'][W'], Yh planar [n][6][H'/2][W'/2])
    or None when the fused kernels do not cover the request.  internal: LoLo is only the next level's input (not
    returned to the user) and lives in reusable scratch."""
    if not _fused_ok(x):
        return None
    n, r, c = x.shape
    Lr, Lc = r + pad_hi[0], c + pad_hi[1]
    k0, p0, m0 = _taps(h0o)
    k1, p1, m1 = _taps(h1o)
    lolo = scratch("lolo1", (n, Lr, Lc), x.dtype, x.device) if internal else torch.empty((n, Lr, Lc), dtype=x.dtype, device=x.device)
    yh = new_highpass(n, 6, (Lr // 2, Lc // 2), x.dtype, x.device)
    with _on_device(x):
        ok = _lib.call_optional("fwd2d_level1", "f32", _ptr(x), _ptr(lolo), _ptr(yh), n, r, c, pad_hi[0], pad_hi[1],
                                p0, m0, p1, m1, yh.stride(0), yh.stride(1), yh.stride(2), _stream(x))
    return (lolo, yh) if ok else None


def fwd2d_levelq(x, lo_a, lo_b, hi_a, hi_b, pad, internal=None):
    """Fused level >= 2 of the 2-D forward transform; pad = (pad_r, pad_c) in {0, 1} (one sample each side).
    internal: a tag when LoLo is only the next level's input (reusable scratch)."""
    if not _fused_ok(x):
        return None
    n, r, c = x.shape
    Lr, Lc = r + 2 * pad[0], c + 2 * pad[1]
    taps = [_taps(h) for h in (lo_a, lo_b, hi_a, hi_b)]
    if len({t[2] for t in taps}) != 1:
        return None
    if internal:
        lolo = scratch(internal, (n, Lr // 2, Lc // 2), x.dtype, x.device)
    else:
        lolo = torch.empty((n, Lr // 2, Lc // 2), dtype=x.dtype, device=x.device)
    yh = new_highpass(n, 6, (Lr // 4, Lc // 4), x.dtype, x.device)
    with _on_device(x):
        ok = _lib.call_optional("fwd2d_levelq", "f32", _ptr(x), _ptr(lolo), _ptr(yh), n, r, c, pad[0], pad[1],
                                taps[0][1], taps[1][1], taps[2][1], taps[3][1], taps[0][2],
                                yh.stride(0), yh.stride(1), yh.stride(2), _stream(x))
    return (lolo, yh) if ok else None


def chain_mode():
    """How levels 1 and 2 of a batched 2-D transform are launched: 0 = one launch per level over the whole batch;
    1..100 = chained chunk by chunk with the level-1 lowpass L2-persisting (hit ratio in percent); -1 = chained
    without the cache policy.  DTCWT_B200_CHAIN overrides the default (read per call: an experiment switch)."""
    try:
        return int(os.environ.get("DTCWT_B200_CHAIN", CHAIN_DEFAULT))
    except ValueError:
        return CHAIN_DEFAULT


CHAIN_DEFAULT = 0
CHAIN_BUDGET_MB = 64      # LoLo1 of one chunk: half of the B200's 126 MB L2


def chain_chunk(n, rows, cols):
    """Images per chunk of the chained levels (0: do not chain -- one image, or images too large for L2 / too small
    to fill the GPU one chunk at a time)."""
    budget = int(os.environ.get("DTCWT_B200_CHAIN_MB", CHAIN_BUDGET_MB)) << 20
    min_pix = int(os.environ.get("DTCWT_B200_CHAIN_MIN_PIX", 1 << 20))      # tests chain small images
    per = 4 * rows * cols
    if n < 2 or per > budget or rows * cols < min_pix:
        return 0
    return max(1, min(n, budget // per))


def fwd2d_level12(x, h0o, h1o, lo_a, lo_b, hi_a, hi_b, pad_hi, internal2=None):
    """Levels 1 and 2 of the forward transform chained chunk by chunk (LoLo1 never returned): x [n][H][W] ->
    (LoLo2, Yh1 planar, Yh2 planar), or None when not applicable (then the caller runs the levels one by one)."""
    mode = chain_mode()
    if mode == 0 or not _fused_ok(x):
        return None
    n, r, c = x.shape
    Lr, Lc = r + pad_hi[0], c + pad_hi[1]
    chunk = chain_chunk(n, Lr, Lc)
    if chunk == 0:
        return None
    k0, p0, m0 = _taps(h0o)
    k1, p1, m1 = _taps(h1o)
    taps = [_taps(h) for h in (lo_a, lo_b, hi_a, hi_b)]
    if len({t[2] for t in taps}) != 1:
        return None
    pr, pc = (1 if Lr % 4 else 0), (1 if Lc % 4 else 0)
    r2, c2 = (Lr + 2 * pr) // 2, (Lc + 2 * pc) // 2
    lolo1 = scratch("lolo1c", (chunk, Lr, Lc), x.dtype, x.device)
    lolo2 = scratch(internal2, (n, r2, c2), x.dtype, x.device) if internal2 else torch.empty((n, r2, c2), dtype=x.dtype, device=x.device)
    yh1 = new_highpass(n, 6, (Lr // 2, Lc // 2), x.dtype, x.device)
    yh2 = new_highpass(n, 6, (r2 // 2, c2 // 2), x.dtype, x.device)
    with _on_device(x):
        ok = _lib.call_optional("fwd2d_level12", "f32", _ptr(x), _ptr(lolo1), _ptr(lolo2), _ptr(yh1), _ptr(yh2), n, r, c,
                                pad_hi[0], pad_hi[1], p0, m0, p1, m1, taps[0][1], taps[1][1], taps[2][1], taps[3][1], taps[0][2],
                                yh1.stride(0), yh1.stride(1), yh1.stride(2), yh2.stride(0), yh2.stride(1), yh2.stride(2),
                                chunk, max(mode, 0), _stream(x), launches=2 * ((n + chunk - 1) // chunk))
    return (lolo2, yh1, yh2) if ok else None


def inv2d_level21(z2, yh2, yh1, lo_a, lo_b, hi_a, hi_b, gain2, g0o, g1o, gain1, crop):
    """Levels 2 and 1 of the inverse transform chained chunk by chunk; None when not applicable."""
    mode = chain_mode()
    if mode == 0 or not _fused_ok(z2, yh2, yh1) or not yh2.is_contiguous() or not yh1.is_contiguous():
        return None
    n, r, c = z2.shape
    r1, c1 = 2 * r - 2 * crop[0], 2 * c - 2 * crop[1]
    chunk = chain_chunk(n, r1, c1)
    if chunk == 0:
        return None
    taps = [_taps(h) for h in (lo_a, lo_b, hi_a, hi_b)]
    if len({t[2] for t in taps}) != 1:
        return None
    k0, p0, m0 = _taps(g0o)
    k1, p1, m1 = _taps(g1o)
    g2k, g2p = _gain6(gain2)
    g1k, g1p = _gain6(gain1)
    z1 = scratch("z1c", (chunk, r1, c1), z2.dtype, z2.device)
    out = torch.empty((n, r1, c1), dtype=z2.dtype, device=z2.device)
    with _on_device(z2):
        ok = _lib.call_optional("inv2d_level21", "f32", _ptr(z2), _ptr(yh2), _ptr(yh1), _ptr(z1), _ptr(out), n, r, c,
                                crop[0], crop[1], taps[0][1], taps[1][1], taps[2][1], taps[3][1], taps[0][2], g2p,
                                p0, m0, p1, m1, g1p, yh2.stride(0), yh2.stride(1), yh2.stride(2),
                                yh1.stride(0), yh1.stride(1), yh1.stride(2), chunk, max(mode, 0), _stream(z2),
                                launches=2 * ((n + chunk - 1) // chunk))
    return out if ok else None


def _gain6(gain):
    g = np.ascontiguousarray(np.asarray(gain, dtype=np.float64).reshape(6))
    return g, g.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def inv2d_levelq(z, yh, lo_a, lo_b, hi_a, hi_b, gain, crop):
    """Fused level >= 2 of the 2-D inverse: z [n][r][c], yh planar [n][6][r/2][c/2] -> [n][2r-2crop_r][2c-2crop_c]."""
    if not _fused_ok(z, yh) or not yh.is_contiguous():
        return None
    n, r, c = z.shape
    taps = [_taps(h) for h in (lo_a, lo_b, hi_a, hi_b)]
    if len({t[2] for t in taps}) != 1:
        return None
    gk, gp = _gain6(gain)
    out = torch.empty((n, 2 * r - 2 * crop[0], 2 * c - 2 * crop[1]), dtype=z.dtype, device=z.device)
    with _on_device(z):
        ok = _lib.call_optional("inv2d_levelq", "f32", _ptr(z), _ptr(yh), _ptr(out), n, r, c, crop[0], crop[1],
                                taps[0][1], taps[1][1], taps[2][1], taps[3][1], taps[0][2], gp,
                                yh.stride(0), yh.stride(1), yh.stride(2), _stream(z))
    return out if ok else None


def inv2d_level1(z, yh, g0o, g1o, gain):
    if not _fused_ok(z, yh) or not yh.is_contiguous():
        return None
    n, r, c = z.shape
    k0, p0, m0 = _taps(g0o)
    k1, p1, m1 = _taps(g1o)
    gk, gp = _gain6(gain)
    out = torch.empty((n, r, c), dtype=z.dtype, device=z.device)
    with _on_device(z):
        ok = _lib.call_optional("inv2d_level1", "f32", _ptr(z), _ptr(yh), _ptr(out), n, r, c, p0, m0, p1, m1, gp,
                                yh.stride(0), yh.stride(1), yh.stride(2), _stream(z))
    return out if ok else None


# `_bp` families: the second launch of a level (bands 1 and 4 from the band-pass pair h2 / g2)
def fwd2d_level1_hh(x, yh, h2o, pad_hi):
    if not _fused_ok(x):
        return False
    n, r, c = x.shape
    k, p, m = _taps(h2o)
    with _on_device(x):
        return _lib.call_optional("fwd2d_level1_hh", "f32", _ptr(x), _ptr(yh), n, r, c, pad_hi[0], pad_hi[1], p, m,
                                  yh.stride(0), yh.stride(1), yh.stride(2), _stream(x))


def fwd2d_levelq_hh(x, yh, h2_a, h2_b, pad):
    if not _fused_ok(x):
        return False
    n, r, c = x.shape
    ta, tb = _taps(h2_a), _taps(h2_b)
    if ta[2] != tb[2]:
        return False
    with _on_device(x):
        return _lib.call_optional("fwd2d_levelq_hh", "f32", _ptr(x), _ptr(yh), n, r, c, pad[0], pad[1], ta[1], tb[1], ta[2],
                                  yh.stride(0), yh.stride(1), yh.stride(2), _stream(x))


def inv2d_levelq_hh(yh, out, rows, cols, g2_a, g2_b, gain, crop):
    """out += H:g2(V:g2(c2q(bands 1, 4))) for a level >= 2; rows x cols is the level's lowpass size."""
    if not _fused_ok(yh, out) or not yh.is_contiguous():
        return False
    ta, tb = _taps(g2_a), _taps(g2_b)
    if ta[2] != tb[2]:
        return False
    gk, gp = _gain6(gain)
    with _on_device(out):
        return _lib.call_optional("inv2d_levelq_hh", "f32", _ptr(yh), _ptr(out), yh.shape[0], rows, cols, crop[0], crop[1],
                                  ta[1], tb[1], ta[2], gp, yh.stride(0), yh.stride(1), yh.stride(2), _stream(out))


def inv2d_level1_hh(yh, out, g2o, gain):
    if not _fused_ok(yh, out) or not yh.is_contiguous():
        return False
    n, r, c = out.shape
    k, p, m = _taps(g2o)
    gk, gp = _gain6(gain)
    with _on_device(out):
        return _lib.call_optional("inv2d_level1_hh", "f32", _ptr(yh), _ptr(out), n, r, c, p, m, gp,
                                  yh.stride(0), yh.stride(1), yh.stride(2), _stream(out))


# ----------------------------------------------------------------------------- 1-D packing
def pack1d(hi):
    """real [2k][c] -> complex [k][c]: even rows real part, odd rows imaginary part."""
    k2, c = hi.shape
    z = torch.empty((k2 // 2, c), dtype=_COMPLEX[hi.dtype], device=hi.device)
    with _on_device(hi):
        _lib.call("pack1d", _SUFFIX[hi.dtype], _ptr(hi), _ptr(z), 1, k2 // 2, c, _stream(hi))
    return z


def unpack1d(z, gain):
    k, c = z.shape
    hi = torch.empty((2 * k, c), dtype=_REAL[z.dtype], device=z.device)
    with _on_device(z):
        _lib.call("unpack1d", _SUFFIX[hi.dtype], _ptr(z), _ptr(hi), 1, k, c, float(gain), _stream(z))
    return hi


# ----------------------------------------------------------------------------- 3-D packing
def cube2c(y, z, chan0):
    """y real [n][2a][2b][2c] -> channels chan0..chan0+3 of planar z [n][28][a][b][c]."""
    n = y.shape[0]
    a, b, c = (s // 2 for s in y.shape[1:])
    assert tuple(z.shape[2:]) == (a, b, c) and z.is_contiguous()
    with _on_device(y):
        _lib.call("cube2c", _SUFFIX[y.dtype], _ptr(y), _ptr(z), n, a, b, c,
                  z.stride(0), z.stride(1), z.stride(2), z.stride(3), z.stride(4), chan0, _stream(y))


def c2cube(z, chan0):
    n, _, a, b, c = z.shape
    y = torch.empty((n, 2 * a, 2 * b, 2 * c), dtype=_REAL[z.dtype], device=z.device)
    with _on_device(z):
        _lib.call("c2cube", _SUFFIX[y.dtype], _ptr(z), _ptr(y), n, a, b, c,
                  z.stride(0), z.stride(1), z.stride(2), z.stride(3), z.stride(4), chan0, _stream(z))
    return y


# ----------------------------------------------------------------------------- fused 3-D levels
def _chan_strides(yh):
    return yh.stride(0), yh.stride(1), yh.stride(2), yh.stride(3), yh.stride(4)


def lowpass3d(x, h, inverse=False):
    """Level 1 of the 3-D transform without highpasses: colfilter(h) along all three axes of x [n][d0][d1][d2]
    (reference transform3d.py:291-315 forward with h0o, :442-456 inverse with g0o); None when not covered."""
    if not _fused_ok(x):
        return None
    n, d0, d1, d2 = x.shape
    k, p, m = _taps(h)
    y = torch.empty_like(x)
    scr = scratch("low3d", x.shape, x.dtype, x.device)
    with _on_device(x):
        ok = _lib.call_optional("inv3d_level1_lo" if inverse else "fwd3d_level1_lo", "f32", _ptr(x), _ptr(y), _ptr(scr),
                                n, d0, d1, d2, p, m, _stream(x))
    return y if ok else None


def fwd3d_level1(x, h0o, h1o):
    """Fused level 1 of the 3-D forward transform (reference _level1_xfm :208-289): -> (LLL, Yh planar [n][28][...])."""
    if not _fused_ok(x):
        return None
    n, d0, d1, d2 = x.shape
    if d0 % 2 or d1 % 2 or d2 % 2:
        return None
    k0, p0, m0 = _taps(h0o)
    k1, p1, m1 = _taps(h1o)
    lll = torch.empty_like(x)
    yh = new_highpass(n, 28, (d0 // 2, d1 // 2, d2 // 2), x.dtype, x.device)
    scr = scratch("fwd3d1", (4,) + tuple(x.shape), x.dtype, x.device)
    with _on_device(x):
        ok = _lib.call_optional("fwd3d_level1", "f32", _ptr(x), _ptr(lll), _ptr(yh), _ptr(scr), n, d0, d1, d2,
                                p0, m0, p1, m1, *_chan_strides(yh), _stream(x))
    return (lll, yh) if ok else None


def fwd3d_levelq(x, lo_a, lo_b, hi_a, hi_b, pads):
    """Fused level >= 2 of the 3-D forward transform (reference _level2_xfm :317-383); pads = replicated samples per
    side of each axis.  -> (LLL [n][L0/2][L1/2][L2/2], Yh planar [n][28][L0/4][L1/4][L2/4]) or None."""
    if not _fused_ok(x):
        return None
    n, d0, d1, d2 = x.shape
    L = [d0 + 2 * pads[0], d1 + 2 * pads[1], d2 + 2 * pads[2]]
    if any(v % 4 for v in L):
        return None
    taps = [_taps(h) for h in (lo_a, lo_b, hi_a, hi_b)]
    if len({t[2] for t in taps}) != 1:
        return None
    lll = torch.empty((n, L[0] // 2, L[1] // 2, L[2] // 2), dtype=x.dtype, device=x.device)
    yh = new_highpass(n, 28, (L[0] // 4, L[1] // 4, L[2] // 4), x.dtype, x.device)
    scr = scratch("fwd3dq", (n * d0 * L[1] * L[2],), x.dtype, x.device)
    with _on_device(x):
        ok = _lib.call_optional("fwd3d_levelq", "f32", _ptr(x), _ptr(lll), _ptr(yh), _ptr(scr), n, d0, d1, d2,
                                pads[0], pads[1], pads[2], taps[0][1], taps[1][1], taps[2][1], taps[3][1], taps[0][2],
                                *_chan_strides(yh), _stream(x))
    return (lll, yh) if ok else None


def inv3d_levelq(yl, yh, lo_a, lo_b, hi_a, hi_b, crops):
    """Fused level >= 2 of the 3-D inverse (reference _level2_ifm :458-526); crops = samples dropped per end of each axis."""
    if not _fused_ok(yl, yh) or not yh.is_contiguous():
        return None
    n, a0, a1, a2 = yl.shape
    taps = [_taps(h) for h in (lo_a, lo_b, hi_a, hi_b)]
    if len({t[2] for t in taps}) != 1:
        return None
    od = (2 * a0 - 2 * crops[0], 2 * a1 - 2 * crops[1], 2 * a2 - 2 * crops[2])
    out = torch.empty((n,) + od, dtype=yl.dtype, device=yl.device)
    scr = scratch("inv3dq", (4 * n * od[0] * a1 * a2,), yl.dtype, yl.device)
    with _on_device(yl):
        ok = _lib.call_optional("inv3d_levelq", "f32", _ptr(yl), _ptr(yh), _ptr(out), _ptr(scr), n, a0, a1, a2,
                                crops[0], crops[1], crops[2], taps[0][1], taps[1][1], taps[2][1], taps[3][1], taps[0][2],
                                *_chan_strides(yh), _stream(yl))
    return out if ok else None


def inv3d_level1(yl, yh, g0o, g1o):
    """Fused level 1 of the 3-D inverse (reference _level1_ifm :385-440)."""
    if not _fused_ok(yl, yh) or not yh.is_contiguous():
        return None
    n, a0, a1, a2 = yl.shape
    k0, p0, m0 = _taps(g0o)
    k1, p1, m1 = _taps(g1o)
    out = torch.empty_like(yl)
    scr = scratch("inv3d1", (4,) + tuple(yl.shape), yl.dtype, yl.device)
    with _on_device(yl):
        ok = _lib.call_optional("inv3d_level1", "f32", _ptr(yl), _ptr(yh), _ptr(out), _ptr(scr), n, a0, a1, a2,
                                p0, m0, p1, m1, *_chan_strides(yh), _stream(yl))
    return out if ok else None
