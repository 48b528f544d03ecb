"""Pin the CPU oracle: MATLAB golden summaries, reference-generated full arrays, live reference."""
import logging

import numpy as np
import pytest

import dtcwt_oracle as O
from dtcwt_b200 import coeffs
from util import MATLAB_ABS_TOL, golden, rel_err, summarise_cube, summarise_mat

logging.disable(logging.WARNING)


def _maxdiff(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max())


# ----------------------------------------------------------------------------- summaries helper
def test_summarise_mat_shape():
    M = np.arange(40 * 50, dtype=float).reshape(40, 50)
    S = summarise_mat(M, 8)
    assert S.shape == (17, 17)
    assert S[0, 0] == M[0, 0] and S[-1, -1] == M[-1, -1]
    assert np.isclose(S[8, 8], M[8:-8, 8:-8].mean())


# ----------------------------------------------------------------------------- MATLAB golden vectors
# mirrors the reference's tests/test_againstmatlab.py:72-124
def test_matlab_coldfilt_colifilt():
    v, mandrill = golden("verification_subset"), golden("inputs")["mandrill"]
    h0a, h0b, g0a, g0b, h1a, h1b, g1a, g1b = coeffs.qshift("qshift_d")
    assert _maxdiff(summarise_mat(O.coldfilt(mandrill, h1b, h1a)), v["mandrill_coldfilt"]) < MATLAB_ABS_TOL
    assert _maxdiff(summarise_mat(O.colifilt(mandrill, g0b, g0a)), v["mandrill_colifilt"]) < MATLAB_ABS_TOL


@pytest.mark.parametrize("biort,qshift,suffix", [("near_sym_a", "qshift_a", ""), ("near_sym_b_bp", "qshift_b_bp", "b")])
def test_matlab_dtwavexfm2(biort, qshift, suffix):
    v, mandrill = golden("verification_subset"), golden("inputs")["mandrill"]
    p = O.Transform2d(coeffs.biort(biort), coeffs.qshift(qshift)).forward(mandrill, 4, include_scale=True)
    assert _maxdiff(summarise_mat(p.lowpass), v["mandrill_Yl" + suffix]) < MATLAB_ABS_TOL
    for i in range(4):
        assert _maxdiff(summarise_mat(p.highpasses[i]), v["mandrill_Yh%s_%d" % (suffix, i)]) < MATLAB_ABS_TOL
        assert _maxdiff(summarise_mat(p.scales[i]), v["mandrill_Yscale%s_%d" % (suffix, i)]) < MATLAB_ABS_TOL


def test_matlab_transform3d():
    v, qbgn = golden("verification_subset"), golden("inputs")["qbgn"]
    p = O.Transform3d(coeffs.biort("near_sym_b"), coeffs.qshift("qshift_b")).forward(qbgn, 3, include_scale=True)
    assert _maxdiff(summarise_cube(p.lowpass), v["qbgn_Yl"]) < MATLAB_ABS_TOL
    for i in range(3):
        assert _maxdiff(summarise_cube(p.highpasses[i]), v["qbgn_Yh_%d" % i]) < MATLAB_ABS_TOL
        assert _maxdiff(summarise_cube(p.scales[i]), v["qbgn_Yscale_%d" % i]) < MATLAB_ABS_TOL


# ----------------------------------------------------------------------------- reference-generated arrays
G = golden("ref_outputs")


def _cases(prefix):
    return sorted({k.split("/")[1] for k in G if k.startswith(prefix + "/")})


def test_ref_lowlevel():
    n = 0
    for key, want in G.items():
        parts = key.split("/")
        if parts[0] != "ll" or len(parts) != 5:
            continue
        _, fn, fam, tap, bits = parts
        X = G["ll/X" + bits]
        if fn == "colfilter":
            h = dict(zip(("h0o", "g0o", "h1o", "g1o"), coeffs.biort(fam)))[tap]
            got = O.colfilter(X, h)
        else:
            q = dict(zip(("h0a", "h0b", "g0a", "g0b", "h1a", "h1b", "g1a", "g1b"), coeffs.qshift(fam)))
            got = getattr(O, fn)(X, q[tap + "b"], q[tap + "a"])
        assert got.dtype == want.dtype
        assert _maxdiff(got, want) == 0.0, key   # same summation order -> bit-identical
        n += 1
    assert n > 100


@pytest.mark.parametrize("tag", _cases("t2"))
def test_ref_transform2d(tag):
    pre = "t2/%s/" % tag
    bn, qn, nlev = G[pre + "meta"]
    nlev = int(nlev)
    t = O.Transform2d(coeffs.biort(bn), coeffs.qshift(qn))
    p = t.forward(G[pre + "X"], nlev, include_scale=True)
    assert rel_err(p.lowpass, G[pre + "Yl"]) < 1e-6
    for i in range(nlev):
        assert p.highpasses[i].dtype == G[pre + "Yh%d" % i].dtype
        assert rel_err(p.highpasses[i], G[pre + "Yh%d" % i]) < 1e-6
        assert rel_err(p.scales[i], G[pre + "Ys%d" % i]) < 1e-6
    assert rel_err(t.inverse(p), G[pre + "Z"]) < 1e-6
    assert rel_err(t.inverse(p, G[pre + "gain_mask"]), G[pre + "Zgain"]) < 1e-6


@pytest.mark.parametrize("tag", _cases("t1"))
def test_ref_transform1d(tag):
    pre = "t1/%s/" % tag
    bn, qn, nlev = G[pre + "meta"]
    nlev = int(nlev)
    t = O.Transform1d(coeffs.biort(bn), coeffs.qshift(qn))
    p = t.forward(G[pre + "X"], nlev, include_scale=True)
    assert rel_err(p.lowpass, G[pre + "Yl"]) < 1e-6
    for i in range(nlev):
        assert rel_err(p.highpasses[i], G[pre + "Yh%d" % i]) < 1e-6
        assert rel_err(p.scales[i], G[pre + "Ys%d" % i]) < 1e-6
    # numpy-2 drift: the reference returns float64 here for float32 data
    assert rel_err(t.inverse(p), G[pre + "Z"]) < 2e-6
    assert rel_err(t.inverse(p, G[pre + "gain_mask"]), G[pre + "Zgain"]) < 2e-6


@pytest.mark.parametrize("tag", _cases("t3"))
def test_ref_transform3d(tag):
    pre = "t3/%s/" % tag
    bn, qn, nlev, em, disc = G[pre + "meta"]
    nlev, em, disc = int(nlev), int(em), bool(int(disc))
    t = O.Transform3d(coeffs.biort(bn), coeffs.qshift(qn), ext_mode=em)
    p = t.forward(G[pre + "X"], nlev, include_scale=True, discard_level_1=disc)
    assert rel_err(p.lowpass, G[pre + "Yl"]) < 1e-6
    for i in range(nlev):
        if disc and i == 0:
            assert p.highpasses[0] is None
        else:
            assert rel_err(p.highpasses[i], G[pre + "Yh%d" % i]) < 1e-6
        assert rel_err(p.scales[i], G[pre + "Ys%d" % i]) < 1e-6
    Z = t.inverse(p)
    want = G[pre + "Z"]
    if disc:
        # reference quirk (transform3d.py:452-454): its result has axes 0 and 2 swapped
        want = want.transpose(2, 1, 0)
    assert rel_err(Z, want) < 2e-6


# ----------------------------------------------------------------------------- contracts (shapes / errors)
def test_filter_contracts():
    X = np.zeros((8, 3), np.float32)
    assert O.colfilter(X, np.ones(5)).shape == (8, 3)       # test_colfilter.py:23-33
    assert O.colfilter(X, np.ones(4)).shape == (9, 3)
    assert O.coldfilt(X, np.ones(6), np.ones(6)).shape == (4, 3)     # test_coldfilt.py:38-40
    assert O.colifilt(X, np.ones(6), np.ones(6)).shape == (16, 3)    # test_colifilt.py:39-53
    with pytest.raises(ValueError):
        O.coldfilt(np.zeros((6, 3)), np.ones(6), np.ones(6))
    with pytest.raises(ValueError):
        O.coldfilt(X, np.ones(5), np.ones(5))
    with pytest.raises(ValueError):
        O.coldfilt(X, np.ones(6), np.ones(4))
    with pytest.raises(ValueError):
        O.colifilt(np.zeros((7, 3)), np.ones(6), np.ones(6))
    assert O.colfilter(np.arange(12).reshape(4, 3), np.ones(3)).dtype == np.float64


def test_reflect_index():
    # reference tests/test_reflect.py semantics with (-0.5, r-0.5)
    r = 5
    got = O.reflect_index(np.arange(-12, 17), r)
    period = list(range(r)) + list(range(r - 1, -1, -1))
    want = [period[i % (2 * r)] for i in range(-12, 17)]
    assert list(got) == want


def test_perfect_reconstruction_f64():
    rs = np.random.RandomState(3)
    X = rs.rand(48, 40)
    t = O.Transform2d(coeffs.biort("near_sym_b"), coeffs.qshift("qshift_b"))
    assert _maxdiff(t.inverse(t.forward(X, 4)), X) < 1e-12          # test_ifm2.py:8,25
    V = rs.rand(16, 24, 20)
    t3 = O.Transform3d(coeffs.biort("near_sym_a"), coeffs.qshift("qshift_a"))
    assert _maxdiff(t3.inverse(t3.forward(V, 3)), V) < 1e-12        # test_xfm3.py:9,40
    x = rs.randn(64)
    t1 = O.Transform1d(coeffs.biort("antonini"), coeffs.qshift("qshift_c"))
    assert _maxdiff(t1.inverse(t1.forward(x, 4)), x) < 1e-12


# ----------------------------------------------------------------------------- live reference (container only)
def _live():
    import refshim
    if not refshim.available():
        pytest.skip("reference checkout not present")
    refshim.load()
    import dtcwt.numpy as dn
    return dn


@pytest.mark.parametrize("shape", [(32, 32), (30, 22), (17, 21), (2, 16)])
@pytest.mark.parametrize("wave", [("near_sym_a", "qshift_a"), ("near_sym_b", "qshift_b"), ("antonini", "qshift_32")])
def test_live_reference_2d(shape, wave):
    dn = _live()
    X = np.random.RandomState(sum(shape)).rand(*shape).astype(np.float32)
    gm = np.random.RandomState(1).rand(6, 3)
    tr = dn.Transform2d(*wave)
    to = O.Transform2d(coeffs.biort(wave[0]), coeffs.qshift(wave[1]))
    pr, po = tr.forward(X, 3, include_scale=True), to.forward(X, 3, include_scale=True)
    assert _maxdiff(pr.lowpass, po.lowpass) == 0
    for a, b in zip(pr.highpasses, po.highpasses):
        assert a.dtype == b.dtype and rel_err(b, a) < 1e-6
    assert rel_err(to.inverse(po, gm), tr.inverse(pr, gm)) < 1e-6


def test_live_reference_coeffs():
    _live()
    from dtcwt import coeffs as rc
    for n in coeffs.BIORT_NAMES:
        for a, b in zip(coeffs.biort(n), rc.biort(n)):
            assert a.shape == b.shape and np.array_equal(a, b)
    for n in coeffs.QSHIFT_NAMES:
        for a, b in zip(coeffs.qshift(n), rc.qshift(n)):
            assert a.shape == b.shape and np.array_equal(a, b)
    with pytest.raises(IOError):
        coeffs.biort("nonsuch")
    with pytest.raises(ValueError):
        coeffs.biort("qshift_a")
    with pytest.raises(ValueError):
        coeffs.qshift("near_sym_a")
