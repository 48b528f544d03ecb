"""Parity of the dtcwt_b200 host layer + kernels with the reference.

Every test takes the ``backend`` fixture and therefore runs twice: against the
host emulator of the kernel bodies in the CPU suite, and against the CUDA library
on a B200 (``-m gpu``).  Expected values are (a) the committed outputs of the
unmodified reference (``tests/golden/ref_outputs.npz``), (b) the MATLAB golden
summaries, (c) the CPU oracle.  Tolerances: 1e-5 relative for float32 (the
north-star tolerance), 1e-12 for float64 (the reference's own, test_ifm2.py:8).
"""
import logging

import numpy as np
import pytest
import torch

import dtcwt_b200
import dtcwt_oracle as O
from dtcwt_b200 import coeffs, lowlevel
from util import MATLAB_ABS_TOL, REL_TOL, golden, rel_err, summarise_cube, summarise_mat

logging.disable(logging.WARNING)
G = golden("ref_outputs")


def tol(dtype):
    return REL_TOL if np.dtype(dtype) in (np.float32, np.complex64) else 1e-12


def npy(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def _cases(prefix):
    return sorted({k.split("/")[1] for k in G if k.startswith(prefix + "/")})


# ----------------------------------------------------------------------------- low-level filters
def test_lowlevel_vs_reference(backend):
    n = 0
    for key, want in G.items():
        parts = key.split("/")
        if parts[0] != "ll" or len(parts) != 5:
            continue
        _, fn, fam, tap, bits = parts
        X = G["ll/X" + bits]
        if fn == "colfilter":
            h = dict(zip(("h0o", "g0o", "h1o", "g1o"), coeffs.biort(fam)))[tap]
            got = lowlevel.colfilter(X, h)
        else:
            q = dict(zip(("h0a", "h0b", "g0a", "g0b", "h1a", "h1b", "g1a", "g1b"), coeffs.qshift(fam)))
            got = getattr(lowlevel, fn)(X, q[tap + "b"], q[tap + "a"])
        got = npy(got)
        assert got.dtype == want.dtype, key
        assert rel_err(got, want) < tol(want.dtype), key
        n += 1
    assert n > 100


def test_lowlevel_axis_and_tiny(backend):
    rs = np.random.RandomState(5)
    X = rs.rand(3, 8, 12).astype(np.float32)
    h = coeffs.biort("near_sym_b")[2]
    ha, hb = coeffs.qshift("qshift_b")[1], coeffs.qshift("qshift_b")[0]
    for ax in range(3):
        want = np.moveaxis(O.colfilter(np.moveaxis(X, ax, 0), h), 0, ax)
        assert rel_err(npy(lowlevel.colfilter(X, h, axis=ax)), want) < REL_TOL
    for ax in (1, 2):
        want = np.moveaxis(O.coldfilt(np.moveaxis(X, ax, 0), ha, hb), 0, ax)
        assert rel_err(npy(lowlevel.coldfilt(X, ha, hb, axis=ax)), want) < REL_TOL
        want = np.moveaxis(O.colifilt(np.moveaxis(X, ax, 0), ha, hb), 0, ax)
        assert rel_err(npy(lowlevel.colifilt(X, ha, hb, axis=ax)), want) < REL_TOL
    # fewer rows than taps: the reflection wraps several times (reference lowlevel.py:73)
    T = rs.rand(2, 5).astype(np.float32)
    assert rel_err(npy(lowlevel.colfilter(T, h)), O.colfilter(T, h)) < REL_TOL
    assert rel_err(npy(lowlevel.colifilt(T, ha, hb)), O.colifilt(T, ha, hb)) < REL_TOL
    T4 = rs.rand(4, 5).astype(np.float32)
    assert rel_err(npy(lowlevel.coldfilt(T4, ha, hb)), O.coldfilt(T4, ha, hb)) < REL_TOL


def test_lowlevel_contracts(backend):
    # reference tests/test_colfilter.py:23-50, test_coldfilt.py:20-40, test_colifilt.py:20-53
    X = np.zeros((8, 3), np.float32)
    assert tuple(lowlevel.colfilter(X, np.ones(5)).shape) == (8, 3)
    assert tuple(lowlevel.colfilter(X, np.ones(4)).shape) == (9, 3)
    assert tuple(lowlevel.coldfilt(X, np.ones(6), np.ones(6)).shape) == (4, 3)
    assert tuple(lowlevel.colifilt(X, np.ones(6), np.ones(6)).shape) == (16, 3)
    assert float(lowlevel.colfilter(X, np.ones(5)).abs().max()) == 0.0
    assert tuple(lowlevel.colfilter([[1, 2], [3, 4]], [1, 2, 1]).shape) == (2, 2)   # lists accepted
    assert lowlevel.colfilter(np.arange(12).reshape(4, 3), np.ones(3)).dtype == torch.float64
    with pytest.raises(ValueError):
        lowlevel.coldfilt(np.zeros((6, 3)), np.ones(6), np.ones(6))
    with pytest.raises(ValueError):
        lowlevel.coldfilt(X, np.ones(5), np.ones(5))
    with pytest.raises(ValueError):
        lowlevel.coldfilt(X, np.ones(6), np.ones(4))
    with pytest.raises(ValueError):
        lowlevel.colifilt(np.zeros((7, 3)), np.ones(6), np.ones(6))
    with pytest.raises(ValueError):
        lowlevel.colifilt(X, np.ones(5), np.ones(5))
    with pytest.raises(ValueError):
        lowlevel.colifilt(X, np.ones(6), np.ones(4))


# ----------------------------------------------------------------------------- 2-D
def _check_pyramid(p, pre, nlev, t):
    assert rel_err(p.lowpass, G[pre + "Yl"]) < t
    assert p.lowpass.dtype == G[pre + "Yl"].dtype
    for i in range(nlev):
        want = G[pre + "Yh%d" % i]
        assert p.highpasses[i].shape == want.shape
        assert p.highpasses[i].dtype == want.dtype
        assert rel_err(p.highpasses[i], want) < t
        assert rel_err(p.scales[i], G[pre + "Ys%d" % i]) < t


@pytest.mark.parametrize("tag", _cases("t2"))
def test_transform2d_vs_reference(backend, tag):
    pre = "t2/%s/" % tag
    bn, qn, nlev = G[pre + "meta"]
    nlev = int(nlev)
    X = G[pre + "X"]
    t = tol(X.dtype)
    xf = dtcwt_b200.Transform2d(bn, qn)
    p = xf.forward(X, nlev, include_scale=True)
    _check_pyramid(p, pre, nlev, t)
    assert rel_err(npy(xf.inverse(p)), G[pre + "Z"]) < t
    assert rel_err(npy(xf.inverse(p, G[pre + "gain_mask"])), G[pre + "Zgain"]) < t
    # the inverse also accepts the reference's own (NumPy, interleaved) pyramid
    ref_p = O.Pyramid(G[pre + "Yl"], tuple(G[pre + "Yh%d" % i] for i in range(nlev)))
    assert rel_err(npy(xf.inverse(ref_p)), G[pre + "Z"]) < t


@pytest.mark.parametrize("biort,qshift,suffix", [("near_sym_a", "qshift_a", ""), ("near_sym_b_bp", "qshift_b_bp", "b")])
def test_transform2d_vs_matlab(backend, biort, qshift, suffix):
    # reference tests/test_againstmatlab.py:84-102
    v, mandrill = golden("verification_subset"), golden("inputs")["mandrill"]
    p = dtcwt_b200.Transform2d(biort, qshift).forward(mandrill, 4, include_scale=True)
    assert np.abs(summarise_mat(p.lowpass) - v["mandrill_Yl" + suffix]).max() < MATLAB_ABS_TOL
    for i in range(4):
        assert np.abs(summarise_mat(p.highpasses[i]) - v["mandrill_Yh%s_%d" % (suffix, i)]).max() < MATLAB_ABS_TOL
        assert np.abs(summarise_mat(p.scales[i]) - v["mandrill_Yscale%s_%d" % (suffix, i)]).max() < MATLAB_ABS_TOL


def test_lowlevel_vs_matlab(backend):
    # reference tests/test_againstmatlab.py:72-82 (qshift_d on mandrill)
    v, mandrill = golden("verification_subset"), golden("inputs")["mandrill"]
    h0a, h0b, g0a, g0b, h1a, h1b, g1a, g1b = coeffs.qshift("qshift_d")
    assert np.abs(summarise_mat(npy(lowlevel.coldfilt(mandrill, h1b, h1a))) - v["mandrill_coldfilt"]).max() < MATLAB_ABS_TOL
    assert np.abs(summarise_mat(npy(lowlevel.colifilt(mandrill, g0b, g0a))) - v["mandrill_colifilt"]).max() < MATLAB_ABS_TOL


def test_config2_mandrill_near_sym_b(backend):
    """BASELINE.json config 2: 512x512 mandrill, 4 levels, near_sym_b + qshift_b.  verification.npz has
    no 2-D entry for this pair, so: full arrays vs the oracle + reconstruction."""
    mandrill = golden("inputs")["mandrill"]
    xf = dtcwt_b200.Transform2d("near_sym_b", "qshift_b")
    p = xf.forward(mandrill, 4)
    po = O.Transform2d(coeffs.biort("near_sym_b"), coeffs.qshift("qshift_b")).forward(mandrill, 4)
    assert rel_err(p.lowpass, po.lowpass) < REL_TOL
    for a, b in zip(p.highpasses, po.highpasses):
        assert a.dtype == np.complex64 and rel_err(a, b) < REL_TOL
    Z = npy(xf.inverse(p))
    assert Z.dtype == np.float32
    assert rel_err(Z, mandrill) < REL_TOL


@pytest.mark.parametrize("shape", [(1, 30), (2, 40), (17, 21), (30, 22), (36, 44)])
def test_transform2d_shapes_vs_oracle(backend, shape):
    # odd sizes (test_xfm2.py:41-57), 1-row input (:27), non-multiple-of-4 levels (test_ifm2.py:13,27-31)
    rs = np.random.RandomState(sum(shape))
    X = rs.rand(*shape).astype(np.float32)
    gm = rs.rand(6, 3)
    xf = dtcwt_b200.Transform2d("near_sym_b", "qshift_b")
    to = O.Transform2d(coeffs.biort("near_sym_b"), coeffs.qshift("qshift_b"))
    p, po = xf.forward(X, 3, include_scale=True), to.forward(X, 3, include_scale=True)
    assert rel_err(p.lowpass, po.lowpass) < REL_TOL
    for a, b in zip(p.highpasses + p.scales, po.highpasses + po.scales):
        assert a.shape == b.shape and rel_err(a, b) < REL_TOL
    assert rel_err(npy(xf.inverse(p, gm)), to.inverse(po, gm)) < REL_TOL


def test_transform2d_api_contracts(backend):
    xf = dtcwt_b200.Transform2d()
    X = np.random.RandomState(0).rand(16, 16)
    with pytest.raises(ValueError):                      # test_xfm2.py:30-32
        xf.forward(np.zeros((4, 4, 4)))
    p0 = xf.forward(X, 0)                                # test_xfm2.py:64-73
    assert p0.highpasses == () and np.array_equal(p0.lowpass, X)
    p0 = xf.forward(X[:15, :13], 0, include_scale=True)
    assert p0.lowpass.shape == (16, 14) and p0.scales == ()
    pi = xf.forward(np.arange(64).reshape(8, 8), 2)       # integer input -> float64 (test_xfm2.py:75-87)
    assert pi.lowpass.dtype == np.float64 and pi.highpasses[0].dtype == np.complex128
    assert rel_err(npy(xf.inverse(pi)), np.arange(64).reshape(8, 8)) < 1e-12
    p32 = xf.forward(X.astype(np.float32), 2)             # float32 stays float32 (test_xfm2.py:89-93)
    assert p32.lowpass.dtype == np.float32 and p32.highpasses[1].dtype == np.complex64
    assert npy(xf.inverse(p32)).dtype == np.float32       # test_ifm2.py:39-46
    p = xf.forward(X, 3)
    bad = dtcwt_b200.Pyramid(p.lowpass_t[:-2], p.highpasses_t)
    with pytest.raises(ValueError):                      # transform2d.py:270-271
        xf.inverse(bad)
    with pytest.raises(ValueError):
        dtcwt_b200.Transform2d(biort=(1, 2, 3)).forward(X)
    # custom taps given as tuples (test_xfm2.py:95-103)
    xt = dtcwt_b200.Transform2d(coeffs.biort("antonini"), coeffs.qshift("qshift_06"))
    assert rel_err(npy(xt.inverse(xt.forward(X, 3))), X) < 1e-12


def test_transform2d_batch(backend):
    rs = np.random.RandomState(11)
    X = rs.rand(3, 24, 20).astype(np.float32)
    gm = rs.rand(6, 3)
    xf = dtcwt_b200.Transform2d("near_sym_b", "qshift_b")
    pb = xf.forward_channels(X, "nhw", 3, include_scale=True)
    assert pb.highpasses[0].shape == (3, 12, 10, 6)
    Zb = npy(xf.inverse_channels(pb, "nhw", gm))
    for i in range(3):
        ps = xf.forward(X[i], 3, include_scale=True)
        assert np.array_equal(pb.lowpass[i], ps.lowpass)
        for a, b in zip(pb.highpasses, ps.highpasses):
            assert np.array_equal(a[i], b)
        assert np.array_equal(Zb[i], npy(xf.inverse(ps, gm)))
    p4 = xf.forward_channels(X.reshape(3, 1, 24, 20), "nchw", 2)
    assert p4.lowpass.shape == (3, 1, 12, 10) and p4.highpasses[1].shape == (3, 1, 6, 5, 6)
    assert rel_err(npy(xf.inverse_channels(p4, "nchw")), X.reshape(3, 1, 24, 20)) < REL_TOL


# ----------------------------------------------------------------------------- 1-D
@pytest.mark.parametrize("tag", _cases("t1"))
def test_transform1d_vs_reference(backend, tag):
    pre = "t1/%s/" % tag
    bn, qn, nlev = G[pre + "meta"]
    nlev = int(nlev)
    X = G[pre + "X"]
    t = tol(X.dtype)
    xf = dtcwt_b200.Transform1d(bn, qn)
    p = xf.forward(X, nlev, include_scale=True)
    assert rel_err(p.lowpass, G[pre + "Yl"]) < t
    for i in range(nlev):
        assert p.highpasses[i].dtype == G[pre + "Yh%d" % i].dtype
        assert rel_err(p.highpasses[i], G[pre + "Yh%d" % i]) < t
        assert rel_err(p.scales[i], G[pre + "Ys%d" % i]) < t
    Z = npy(xf.inverse(p))
    assert Z.shape == G[pre + "Z"].shape        # 1-D in -> 1-D out (transform1d.py:177-180)
    assert Z.dtype == X.dtype                   # the reference drifts to float64 under numpy 2
    assert rel_err(Z, G[pre + "Z"]) < t
    assert rel_err(npy(xf.inverse(p, G[pre + "gain_mask"])), G[pre + "Zgain"]) < t


def test_transform1d_contracts(backend):
    xf = dtcwt_b200.Transform1d()
    with pytest.raises(ValueError):              # transform1d.py:70-71
        xf.forward(np.zeros(7))
    x = np.random.RandomState(4).randn(50)       # 25 -> not a multiple of 4 at level 2: pad + crop
    assert rel_err(npy(xf.inverse(xf.forward(x, 4))), x) < 1e-12
    p0 = xf.forward(x, 0)
    assert p0.highpasses == ()


# ----------------------------------------------------------------------------- 3-D
@pytest.mark.parametrize("tag", _cases("t3"))
def test_transform3d_vs_reference(backend, tag):
    pre = "t3/%s/" % tag
    bn, qn, nlev, em, disc = G[pre + "meta"]
    nlev, em, disc = int(nlev), int(em), bool(int(disc))
    X = G[pre + "X"]
    t = tol(X.dtype)
    xf = dtcwt_b200.Transform3d(bn, qn, ext_mode=em)
    p = xf.forward(X, nlev, include_scale=True, discard_level_1=disc)
    assert rel_err(p.lowpass, G[pre + "Yl"]) < t
    for i in range(nlev):
        if disc and i == 0:
            assert p.highpasses[0] is None
        else:
            want = G[pre + "Yh%d" % i]
            assert p.highpasses[i].shape == want.shape and p.highpasses[i].dtype == want.dtype
            assert rel_err(p.highpasses[i], want) < t
        assert rel_err(p.scales[i], G[pre + "Ys%d" % i]) < t
    want = G[pre + "Z"]
    if disc:   # reference quirk, transform3d.py:452-454: its result has axes 0 and 2 swapped
        want = want.transpose(2, 1, 0)
    assert rel_err(npy(xf.inverse(p)), want) < 2 * t


def test_transform3d_vs_matlab(backend):
    # reference tests/test_againstmatlab.py:115-124 -- 128^3 qbgn, near_sym_b / qshift_b, 3 levels
    v, qbgn = golden("verification_subset"), golden("inputs")["qbgn"]
    p = dtcwt_b200.Transform3d("near_sym_b", "qshift_b").forward(qbgn, 3, include_scale=True)
    assert np.abs(summarise_cube(p.lowpass) - v["qbgn_Yl"]).max() < MATLAB_ABS_TOL
    for i in range(3):
        assert np.abs(summarise_cube(p.highpasses[i]) - v["qbgn_Yh_%d" % i]).max() < MATLAB_ABS_TOL
        assert np.abs(summarise_cube(p.scales[i]) - v["qbgn_Yscale_%d" % i]).max() < MATLAB_ABS_TOL


def test_transform3d_ext_modes_and_batch(backend):
    rs = np.random.RandomState(8)
    V = rs.rand(30, 22, 26)                       # test_xfm3.py:109-121 style non-cubic crop, ext_mode 4
    xf = dtcwt_b200.Transform3d("near_sym_a", "qshift_a", ext_mode=4)
    assert rel_err(npy(xf.inverse(xf.forward(V, 3))), V) < 1e-12
    V8 = rs.rand(28, 20, 36)                      # test_xfm3.py:95-107, ext_mode 8
    xf8 = dtcwt_b200.Transform3d("near_sym_a", "qshift_a", ext_mode=8)
    assert rel_err(npy(xf8.inverse(xf8.forward(V8, 3))), V8) < 1e-12
    with pytest.raises(ValueError):
        xf.forward(np.zeros((5, 4, 4)))
    with pytest.raises(ValueError):
        dtcwt_b200.Transform3d(ext_mode=3).forward(V)
    B = rs.rand(2, 16, 12, 20).astype(np.float32)
    pb = xf.forward_channels(B, 2, discard_level_1=True)
    assert pb.highpasses[0] is None and pb.highpasses[1].shape == (2, 4, 3, 5, 28)
    Zb = npy(xf.inverse(pb))
    for i in range(2):
        ps = xf.forward(B[i], 2, discard_level_1=True)
        assert np.array_equal(pb.highpasses[1][i], ps.highpasses[1])
        assert np.array_equal(Zb[i], npy(xf.inverse(ps)))


@pytest.mark.parametrize("data_format", ["nhw", "chw", "hwn", "hwc", "nchw", "nhwc"])
def test_forward_inverse_channels_data_formats(backend, data_format):
    """data_format strings of the reference's TensorFlow backend (dtcwt/tf/transform2d.py:179-330): every layout gives
    the per-image pyramids in the documented axis order and inverse_channels returns the input layout."""
    rs = np.random.RandomState(5)
    N, C, H, W = 2, 3, 40, 56
    base = rs.rand(N, C, H, W).astype(np.float32)
    xf = dtcwt_b200.Transform2d("near_sym_a", "qshift_a")
    if data_format in ("nhw", "chw"):
        X, pick, ax = base[:, 0], (lambda A, i, j: A[i]), [(i, 0) for i in range(N)]
    elif data_format in ("hwn", "hwc"):
        X, pick, ax = np.ascontiguousarray(base[:, 0].transpose(1, 2, 0)), (lambda A, i, j: A[:, :, i]), [(i, 0) for i in range(N)]
    elif data_format == "nchw":
        X, pick, ax = base, (lambda A, i, j: A[i, j]), [(i, j) for i in range(N) for j in range(C)]
    else:
        X, pick, ax = np.ascontiguousarray(base.transpose(0, 2, 3, 1)), (lambda A, i, j: A[:, :, :, j][i]), \
            [(i, j) for i in range(N) for j in range(C)]
    p = xf.forward_channels(X, data_format, nlevels=2, include_scale=True)
    for i, j in ax:
        ref = xf.forward(base[i, j], 2, include_scale=True)
        assert np.array_equal(pick(p.lowpass, i, j), ref.lowpass)
        for a, b in zip(p.highpasses, ref.highpasses):
            assert np.array_equal(pick(a, i, j), b)
        for a, b in zip(p.scales, ref.scales):
            assert np.array_equal(pick(a, i, j), b)
    Z = xf.inverse_channels(p, data_format)
    Z = Z.cpu().numpy() if hasattr(Z, "cpu") else np.asarray(Z)
    assert Z.shape == X.shape and np.abs(Z - X).max() < 1e-5
    with pytest.raises(ValueError):
        xf.forward_channels(X, "whn", nlevels=1)
    with pytest.raises(ValueError):
        xf.forward_channels(X[0], data_format, nlevels=1)


def test_empty_batches(backend):
    """Zero images / volumes / columns: every transform returns correctly shaped empty results without a launch."""
    p = dtcwt_b200.Transform2d("near_sym_b", "qshift_b").forward_channels(torch.empty(0, 64, 48), "nhw", nlevels=2)
    assert tuple(p.lowpass_t.shape) == (0, 32, 24) and [tuple(h.shape) for h in p.highpasses_t] == [(0, 32, 24, 6), (0, 16, 12, 6)]
    assert tuple(dtcwt_b200.Transform2d("near_sym_b", "qshift_b").inverse_channels(p, "nhw").shape) == (0, 64, 48)
    x3 = dtcwt_b200.Transform3d("near_sym_b", "qshift_b")
    p3 = x3.forward_channels(torch.empty(0, 16, 16, 16), nlevels=2)
    assert tuple(p3.lowpass_t.shape) == (0, 8, 8, 8) and tuple(x3.inverse(p3).shape) == (0, 16, 16, 16)
    x1 = dtcwt_b200.Transform1d("near_sym_b", "qshift_b")
    p1 = x1.forward(np.zeros((64, 0), np.float32), 2)
    assert p1.lowpass.shape == (32, 0) and npy(x1.inverse(p1)).shape == (64, 0)


# ----------------------------------------------------------------------------- Pyramid: one source of truth
def test_pyramid_numpy_edits_reach_inverse(backend, monkeypatch):
    """Code written for the reference edits ``pyramid.highpasses[l]`` / ``.lowpass`` in place before ``inverse``
    (reference numpy Pyramid: plain mutable attributes, dtcwt/numpy/common.py:5-32).  On a CUDA device the NumPy
    attributes are COPIES of the device tensors; force that on the CPU run too."""
    from dtcwt_b200 import common
    orig = common._to_numpy
    monkeypatch.setattr(common, "_to_numpy", lambda t: None if t is None else np.array(orig(t), copy=True))
    X = np.random.RandomState(8).rand(64, 48).astype(np.float32)
    xf = dtcwt_b200.Transform2d("near_sym_a", "qshift_a")
    p = xf.forward(X, nlevels=2)
    untouched = xf.inverse(xf.forward(X, nlevels=2)).cpu().numpy()
    assert np.abs(untouched - X).max() < 1e-5
    p.highpasses[0][:] = 0
    p.lowpass[:] *= 0.5
    want = O.Transform2d(coeffs.biort("near_sym_a"), coeffs.qshift("qshift_a"))
    po = want.forward(X, 2)
    po.highpasses[0][:] = 0
    po.lowpass[:] *= 0.5
    ref = want.inverse(po)
    got = xf.inverse(p).cpu().numpy()
    assert rel_err(got, ref) < REL_TOL
    assert np.abs(got - X).max() > 1e-2            # the edit really changed the result
    # assignment works like the reference's plain attributes, tensors or arrays
    p.lowpass = np.zeros_like(po.lowpass)
    p.highpasses = tuple(np.zeros_like(h) for h in po.highpasses)
    assert float(xf.inverse(p).abs().max()) == 0.0
    q = xf.forward(X, nlevels=2)
    q.lowpass                                      # reading alone must not change anything
    assert np.abs(xf.inverse(q).cpu().numpy() - X).max() < 1e-5


# ----------------------------------------------------------------------------- fused 3-D levels (fused3d.cuh)
def _launched(fn):
    """Run fn() and return (result, set of C-ABI symbols it launched)."""
    from dtcwt_b200 import _lib
    seen = []

    def hook(symbol, thunk):
        seen.append(symbol)
        thunk()

    _lib.set_launch_hook(hook)
    try:
        out = fn()
    finally:
        _lib.set_launch_hook(None)
    return out, set(seen)


@pytest.mark.parametrize("shape,em,disc,names", [
    ((32, 40, 48), 4, True, ("near_sym_b", "qshift_b")),       # config-4 shape family: level 1 lowpass only
    ((32, 40, 48), 4, False, ("near_sym_b", "qshift_b")),      # full level 1 (28 channels at n/2)
    ((38, 42, 34), 4, False, ("near_sym_a", "qshift_a")),      # level 2 input not a multiple of 4: pad 1 / crop 1 on every axis
    ((40, 36, 44), 8, True, ("antonini", "qshift_06")),        # ext_mode 8: pad 2 / crop 2 (transform3d.py:329-335, 515-524)
    ((32, 64, 32), 4, True, ("legall", "qshift_d")),           # 18-tap q-shift pair
])
def test_fused3d_levels_vs_oracle(backend, shape, em, disc, names):
    """The fused 3-D level kernels (slices + depth pass, packers in registers) against the CPU oracle, and the launch
    record shows that the fused entry points -- not the per-axis composition -- produced the result."""
    rs = np.random.RandomState(12)
    X = rs.rand(2, *shape).astype(np.float32)
    xf = dtcwt_b200.Transform3d(*names, ext_mode=em)
    p, fwd_syms = _launched(lambda: xf.forward_channels(X, 2, discard_level_1=disc))
    Z, inv_syms = _launched(lambda: xf.inverse(p))
    assert "dtcwt_b200_fwd3d_levelq_f32" in fwd_syms and "dtcwt_b200_inv3d_levelq_f32" in inv_syms
    assert ("dtcwt_b200_fwd3d_level1_lo_f32" if disc else "dtcwt_b200_fwd3d_level1_f32") in fwd_syms
    assert ("dtcwt_b200_inv3d_level1_lo_f32" if disc else "dtcwt_b200_inv3d_level1_f32") in inv_syms
    assert not any(s.startswith(("dtcwt_b200_col", "dtcwt_b200_cube2c", "dtcwt_b200_c2cube")) for s in fwd_syms | inv_syms)
    to = O.Transform3d(coeffs.biort(names[0]), coeffs.qshift(names[1]), ext_mode=em)
    for i in range(2):
        po = to.forward(X[i], 2, discard_level_1=disc)
        assert rel_err(p.lowpass[i], po.lowpass) < REL_TOL
        for l in range(2):
            if po.highpasses[l] is None:
                assert p.highpasses[l] is None
            else:
                assert p.highpasses[l][i].shape == po.highpasses[l].shape
                assert rel_err(p.highpasses[l][i], po.highpasses[l]) < REL_TOL
        assert rel_err(npy(Z)[i], to.inverse(po)) < 2 * REL_TOL
    if not disc:
        assert np.abs(npy(Z) - X).max() < 1e-5


def test_fused3d_equals_per_axis_composition(backend):
    from dtcwt_b200 import _ops
    X = np.random.RandomState(13).rand(1, 32, 32, 64).astype(np.float32)
    xf = dtcwt_b200.Transform3d("near_sym_b", "qshift_b")
    p = xf.forward_channels(X, 2)
    Z = npy(xf.inverse(p))
    _ops.FUSED_ENABLED = False
    try:
        q = xf.forward_channels(X, 2)
        Zq = npy(xf.inverse(q))
    finally:
        _ops.FUSED_ENABLED = True
    assert rel_err(p.lowpass, q.lowpass) < REL_TOL
    for a, b in zip(p.highpasses, q.highpasses):
        assert rel_err(a, b) < REL_TOL
    assert rel_err(Z, Zq) < REL_TOL


def test_transform3d_even_length_biort(backend):
    """Haar level-1 filters in 3-D (reference tests/test_xfm3.py:42-58; transform3d.py:223-251, 437-438): n+1 lowpass
    samples per axis, highpasses at the original size, the inverse drops the first sample of every axis."""
    h0 = np.array((1.0, 1.0))
    g0 = h0 / h0.sum()
    h0 = h0 / h0.sum()
    h1 = g0 * np.cumprod(-np.ones_like(g0))
    g1 = -h0 * np.cumprod(-np.ones_like(h0))
    haar = tuple(np.asarray(h).reshape(-1, 1) for h in (h0, g0, h1, g1))
    X = np.random.RandomState(21).rand(12, 16, 20)
    xf = dtcwt_b200.Transform3d(haar, "qshift_a")
    p = xf.forward(X, 1)
    po = O.Transform3d(haar, coeffs.qshift("qshift_a")).forward(X, 1)
    assert p.lowpass.shape == (13, 17, 21) and p.highpasses[0].shape == (6, 8, 10, 28)
    assert rel_err(p.lowpass, po.lowpass) < 1e-12 and rel_err(p.highpasses[0], po.highpasses[0]) < 1e-12
    Z = npy(xf.inverse(p))
    assert Z.shape == X.shape and rel_err(Z, O.Transform3d(haar, coeffs.qshift("qshift_a")).inverse(po)) < 1e-12
    # the reference's own test: an ellipsoid that vanishes at the borders is reconstructed exactly
    g = np.arange(-16, 16) / 16.0
    r = np.sqrt(g[:, None, None] ** 2 * 2 + g[None, :, None] ** 2 * 3 + g[None, None, :] ** 2 * 1.5)
    E = np.where(r < 0.9, 1.0, 0.0)
    Yl, Yh = dtcwt_b200.dtwavexfm3(E, 1, biort=haar)
    assert np.abs(dtcwt_b200.dtwaveifm3(Yl, Yh, biort=haar) - E).max() < 1e-12
    with pytest.raises(ValueError):
        xf.forward(X, 1, discard_level_1=True)


def test_compat_wrappers(backend):
    """dtcwt.compat's MATLAB-style tuple API (reference compat.py:32-288) on this backend."""
    rs = np.random.RandomState(17)
    v = rs.rand(64, 3)
    Yl, Yh, Ys = dtcwt_b200.dtwavexfm(v, 3, include_scale=True)
    po = O.Transform1d(coeffs.biort("near_sym_a"), coeffs.qshift("qshift_a")).forward(v, 3, True)
    assert rel_err(Yl, po.lowpass) < 1e-12 and len(Yh) == 3 and len(Ys) == 3
    assert rel_err(dtcwt_b200.dtwaveifm(Yl, Yh), v) < 1e-10
    X = rs.rand(48, 40).astype(np.float32)
    Yl, Yh = dtcwt_b200.dtwavexfm2(X, 2, "near_sym_b", "qshift_b")
    po = O.Transform2d(coeffs.biort("near_sym_b"), coeffs.qshift("qshift_b")).forward(X, 2)
    assert isinstance(Yl, np.ndarray) and rel_err(Yl, po.lowpass) < REL_TOL and rel_err(Yh[1], po.highpasses[1]) < REL_TOL
    Z = dtcwt_b200.dtwaveifm2(Yl, Yh, "near_sym_b", "qshift_b", gain_mask=np.ones((6, 2)))
    assert isinstance(Z, np.ndarray) and np.abs(Z - X).max() < 1e-5
    assert dtcwt_b200.dtwavexfm2b is dtcwt_b200.dtwavexfm2 and dtcwt_b200.dtwaveifm2b is dtcwt_b200.dtwaveifm2
    Ylb, Yhb = dtcwt_b200.dtwavexfm2b(X, 2, "near_sym_b_bp", "qshift_b_bp")
    pb = O.Transform2d(coeffs.biort("near_sym_b_bp"), coeffs.qshift("qshift_b_bp")).forward(X, 2)
    assert rel_err(Yhb[0], pb.highpasses[0]) < REL_TOL and rel_err(Yhb[1], pb.highpasses[1]) < REL_TOL
    V = rs.rand(16, 16, 16)
    Yl, Yh = dtcwt_b200.dtwavexfm3(V, 2, ext_mode=4, discard_level_1=True)
    assert Yh[0] is None and Yh[1].shape == (4, 4, 4, 28)
    assert dtcwt_b200.dtwaveifm3(Yl, Yh).shape == V.shape
