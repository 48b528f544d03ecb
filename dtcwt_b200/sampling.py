"""Re-sampling of images and complex sub-bands on the GPU -- drop-in for ``dtcwt.sampling``.

Mirrors the reference module (``dtcwt/sampling.py``): ``sample`` (:105), ``rescale`` (:131),
``sample_highpass`` (:192), ``rescale_highpass`` (:224), ``upsample`` (:343), ``upsample_highpass`` (:372),
same argument order, same ``method`` strings (``'lanczos'`` default, ``'bilinear'``, ``'nearest'``), same
pixel-centre convention and symmetric extension.  Arrays may be NumPy or ``torch.Tensor``; results are
``torch.Tensor`` on the GPU.  One CUDA kernel (``csrc/registration.cuh: SampleElem``) does the tap gathering,
the weights and -- for sub-bands -- the phase un-rolling / re-rolling in float64 arithmetic.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib, _ops

__all__ = ("sample", "sample_highpass", "rescale", "rescale_highpass", "upsample", "upsample_highpass",
           "DTHETA_DX_2D", "DTHETA_DY_2D")

_W0 = -3 * np.pi / 2.15
_W1 = -np.pi / 2.15
#: expected phase advance per sub-band along x and y (reference sampling.py:26-32)
DTHETA_DX_2D = np.array((_W1, _W0, _W0, _W0, _W0, _W1))
DTHETA_DY_2D = np.array((_W0, _W0, _W1, -_W1, -_W0, -_W0))

_METHODS = {"nearest": 0, "bilinear": 1, "lanczos": 2}


def _method(method):
    if method is None:
        method = "lanczos"
    if method not in _METHODS:
        raise NotImplementedError('Sampling method "{0}" is not implemented.'.format(method))
    return _METHODS[method]


def _up_method(method):
    """upsample()'s Lanczos kernel has seven un-windowed taps (reference sampling.py:312-320), sample()'s has six"""
    m = _method(method)
    return 3 if m == 2 else m


def _as_image(im):
    """-> (tensor [h][w][C] on the device (real or complex, float32/64 based), squeeze flag)"""
    if not isinstance(im, torch.Tensor):
        im = np.asarray(im)
        if im.dtype not in (np.float32, np.float64, np.complex64, np.complex128):
            im = im.astype(np.complex128 if np.iscomplexobj(im) else np.float64)
        im = torch.from_numpy(np.ascontiguousarray(im))
    if im.dtype not in (torch.float32, torch.float64, torch.complex64, torch.complex128):
        im = im.to(torch.complex128 if im.is_complex() else torch.float64)
    im = _ops.to_device(im)
    if im.dim() == 1:
        im = im.unsqueeze(0)             # np.atleast_2d
    squeeze = im.dim() == 2
    if squeeze:
        im = im.unsqueeze(-1)
    if im.dim() != 3:
        raise ValueError("images must be [h][w] or [h][w][C]")
    return im, squeeze


def _coords(v, device):
    if not isinstance(v, torch.Tensor):
        v = torch.from_numpy(np.ascontiguousarray(np.asarray(v, dtype=np.float64)))
    return _ops.to_device(v).to(torch.float64).contiguous()


def _dptr(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _run(im, out_hw, xs, ys, method, coords_mode, wx=None, wy=None):
    """im [h][w][C] -> [oh][ow][C] through dtcwt_b200_sample_*."""
    h, w, C = im.shape
    oh, ow = out_hw
    cplx = im.is_complex()
    real = torch.view_as_real(im.contiguous()) if cplx else im.contiguous()
    suffix = "f32" if real.dtype == torch.float32 else "f64"
    out = torch.empty((oh, ow, C), dtype=im.dtype, device=im.device)
    null = ctypes.c_void_p(0)
    keep = []
    if wx is not None:
        kx, px = _dptr(wx)
        ky, py = _dptr(wy)
        keep += [kx, ky]
    else:
        px = py = ctypes.POINTER(ctypes.c_double)()
    with _ops._on_device(im):
        _lib.call("sample", suffix, _ops._ptr(real), _ops._ptr(out), _ops._ptr(xs) if xs is not None else null,
                  _ops._ptr(ys) if ys is not None else null, 1, h, w, C, oh, ow,
                  0, w * C, C, 1, 0, ow * C, C, 1, 0, int(cplx), method, coords_mode, px, py, _ops._stream(im))
    return out


def sample(im, xs, ys, method=None):
    """Sample *im* at the positions (*xs*, *ys*) (reference sampling.py:105-129)."""
    m = _method(method)
    im, squeeze = _as_image(im)
    xs, ys = _coords(xs, im.device), _coords(ys, im.device)
    if xs.shape != ys.shape:
        raise ValueError("Shape of xs and ys must match")
    shape = tuple(xs.shape)
    out = _run(im, (1, int(xs.numel())), xs.reshape(1, -1), ys.reshape(1, -1), m, 0)
    out = out.reshape(shape + (im.shape[2],))
    return out[..., 0] if squeeze else out


def rescale(im, shape, method=None):
    """Resample *im* to *shape* over the same extent (reference sampling.py:131-165)."""
    m = _method(method)
    im, squeeze = _as_image(im)
    out = _run(im, (int(shape[0]), int(shape[1])), None, None, m, 1)
    return out[..., 0] if squeeze else out


def _subbands(im, sbs):
    im, _ = _as_image(im)
    if not im.is_complex():
        im = im.to(torch.complex64 if im.dtype == torch.float32 else torch.complex128)
    sbs = np.arange(6) if sbs is None else np.asarray(sbs)
    if len(sbs) != im.shape[2] or not np.array_equal(sbs, np.arange(im.shape[2])):
        im = im[:, :, torch.as_tensor(sbs, device=im.device)]
    return im, DTHETA_DX_2D[sbs], DTHETA_DY_2D[sbs]


def sample_highpass(im, xs, ys, method=None, sbs=None):
    """As :func:`sample` for a ``[h][w][6]`` sub-band array: phase un-rolled to DC, sampled, re-rolled (:192-222)."""
    m = _method(method)
    im, wx, wy = _subbands(im, sbs)
    xs, ys = _coords(xs, im.device), _coords(ys, im.device)
    if xs.shape != ys.shape:
        raise ValueError("Shape of xs and ys must match")
    shape = tuple(xs.shape)
    out = _run(im, (1, int(xs.numel())), xs.reshape(1, -1), ys.reshape(1, -1), m, 0, wx, wy)
    return out.reshape(shape + (im.shape[2],))


def rescale_highpass(im, shape, method=None, sbs=None):
    """As :func:`rescale` for sub-bands (reference sampling.py:224-278)."""
    m = _method(method)
    im, wx, wy = _subbands(im, sbs)
    return _run(im, (int(shape[0]), int(shape[1])), None, None, m, 1, wx, wy)


def upsample(image, method=None):
    """Upsample rows and columns by two (reference sampling.py:343-370).  The reference's separable convolution with
    the kernels sampled at -1/4 and +1/4 lands on the rescale grid x(d) = (d + 0.5) / 2 - 0.5 of a doubled shape; its
    Lanczos variant keeps one tap more than sample() does, which the kernel reproduces (method 3)."""
    im, squeeze = _as_image(image)
    out = _run(im, (2 * im.shape[0], 2 * im.shape[1]), None, None, _up_method(method), 1)
    return out[..., 0] if squeeze else out


def upsample_highpass(im, method=None):
    """As :func:`upsample` with phase rolling (reference sampling.py:372-391)."""
    im, wx, wy = _subbands(im, None)
    return _run(im, (2 * im.shape[0], 2 * im.shape[1]), None, None, _up_method(method), 1, wx, wy)
