#!/usr/bin/env python
"""Generate the committed golden fixtures from the UNMODIFIED reference.

Build-container only (needs ``/root/reference``; see ``oracle/refshim.py`` for
the numpy-2 attribute shims).  Writes, next to this script:

* ``inputs.npz``               -- the reference's own test inputs, verbatim:
                                  ``mandrill`` (512x512 float32, tests/mandrill.npz) and
                                  ``qbgn`` (128^3 uint8, tests/qbgn.npz).
* ``verification_subset.npz``  -- the MATLAB-toolbox golden summaries of
                                  ``tests/verification.npz`` that pin this path
                                  (tests/test_againstmatlab.py:72-124), verbatim.
* ``ref_outputs.npz``          -- FULL arrays produced here by the reference's numpy
                                  backend (``dtcwt.numpy``) on seeded inputs; each case
                                  stores its inputs too, so tests need nothing else.

Usage:  python tests/golden/make_golden.py
"""
import logging
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import refshim  # noqa: E402

VERIF_KEYS_PREFIX = ("mandrill_", "qbgn_")


def main():
    dtcwt = refshim.load()
    import dtcwt.numpy as dn
    from dtcwt import coeffs
    from dtcwt.numpy import lowlevel as LL
    logging.disable(logging.WARNING)
    tests = os.path.join(refshim.REFERENCE_ROOT, "tests")

    mandrill = np.load(os.path.join(tests, "mandrill.npz"))["mandrill"]
    qbgn = np.load(os.path.join(tests, "qbgn.npz"))["qbgn"]
    np.savez_compressed(os.path.join(HERE, "inputs.npz"), mandrill=mandrill, qbgn=qbgn)

    verif = np.load(os.path.join(tests, "verification.npz"))
    keep = {k: verif[k] for k in verif.files if k.startswith(VERIF_KEYS_PREFIX) and k != "mandrill_upsample"}
    np.savez_compressed(os.path.join(HERE, "verification_subset.npz"), **keep)

    out = {}
    rs = np.random.RandomState(20261017)

    # ---- low-level filters (lowlevel.py:47,82,156), float32 and float64
    X32 = rs.rand(24, 7).astype(np.float32)
    X64 = rs.rand(16, 5)
    out["ll/X32"], out["ll/X64"] = X32, X64
    for fam in ("near_sym_a", "near_sym_b", "antonini", "legall"):
        for n, h in zip(("h0o", "g0o", "h1o", "g1o"), coeffs.biort(fam)):
            out["ll/colfilter/%s/%s/32" % (fam, n)] = LL.colfilter(X32, h)
            out["ll/colfilter/%s/%s/64" % (fam, n)] = LL.colfilter(X64, h)
    for fam in ("qshift_06", "qshift_a", "qshift_b", "qshift_c", "qshift_d", "qshift_32"):
        h0a, h0b, g0a, g0b, h1a, h1b, g1a, g1b = coeffs.qshift(fam)
        for n, (a, b) in (("h0", (h0b, h0a)), ("h1", (h1b, h1a)), ("g0", (g0b, g0a)), ("g1", (g1b, g1a))):
            out["ll/coldfilt/%s/%s/32" % (fam, n)] = LL.coldfilt(X32, a, b)
            out["ll/colifilt/%s/%s/32" % (fam, n)] = LL.colifilt(X32, a, b)
            out["ll/coldfilt/%s/%s/64" % (fam, n)] = LL.coldfilt(X64, a, b)
            out["ll/colifilt/%s/%s/64" % (fam, n)] = LL.colifilt(X64, a, b)

    # ---- 2-D transform cases (transform2d.py:40,190)
    cases2d = [
        ("a", "near_sym_a", "qshift_a", (40, 36), 3, np.float32),
        ("b", "near_sym_b", "qshift_b", (64, 48), 4, np.float32),
        ("bp", "near_sym_b_bp", "qshift_b_bp", (48, 40), 3, np.float32),
        ("odd", "near_sym_b", "qshift_b", (33, 27), 3, np.float32),
        ("pad", "antonini", "qshift_c", (36, 44), 3, np.float32),   # 18x22 at level 2 -> %4 padding
        ("f64", "legall", "qshift_06", (24, 20), 2, np.float64),
    ]
    for tag, bn, qn, shape, nlev, dt in cases2d:
        X = rs.rand(*shape).astype(dt)
        gm = rs.rand(6, nlev)
        t = dn.Transform2d(bn, qn)
        p = t.forward(X, nlevels=nlev, include_scale=True)
        pre = "t2/%s/" % tag
        out[pre + "meta"] = np.array([bn, qn, str(nlev)])
        out[pre + "X"], out[pre + "gain_mask"] = X, gm
        out[pre + "Yl"] = p.lowpass
        for i, (h, s) in enumerate(zip(p.highpasses, p.scales)):
            out[pre + "Yh%d" % i], out[pre + "Ys%d" % i] = h, s
        out[pre + "Z"] = t.inverse(p)
        out[pre + "Zgain"] = t.inverse(p, gm)

    # ---- 1-D (transform1d.py:26,112); "c1" is BASELINE.json config 1
    cases1d = [
        ("c1", "near_sym_a", "qshift_a", (256,), 3, np.float32),
        ("cols", "near_sym_b", "qshift_d", (100, 3), 3, np.float32),
        ("f64", "antonini", "qshift_c", (36, 2), 3, np.float64),
    ]
    for tag, bn, qn, shape, nlev, dt in cases1d:
        X = (np.random.RandomState(0).randn(*shape) if tag == "c1" else rs.randn(*shape)).astype(dt)
        gm = rs.rand(nlev)
        t = dn.Transform1d(bn, qn)
        p = t.forward(X, nlevels=nlev, include_scale=True)
        pre = "t1/%s/" % tag
        out[pre + "meta"] = np.array([bn, qn, str(nlev)])
        out[pre + "X"], out[pre + "gain_mask"] = X, gm
        out[pre + "Yl"] = p.lowpass
        for i, (h, s) in enumerate(zip(p.highpasses, p.scales)):
            out[pre + "Yh%d" % i], out[pre + "Ys%d" % i] = h, s
        out[pre + "Z"] = np.asarray(t.inverse(p))
        out[pre + "Zgain"] = np.asarray(t.inverse(p, gm))

    # ---- 3-D (transform3d.py:37,133)
    cases3d = [
        ("a", "near_sym_a", "qshift_a", (16, 16, 16), 2, 4, False, np.float32),
        ("b", "near_sym_b", "qshift_b", (20, 12, 28), 3, 4, False, np.float32),
        ("ext8", "near_sym_a", "qshift_b", (24, 16, 40), 3, 8, False, np.float64),
        ("disc", "near_sym_b", "qshift_b", (16, 16, 16), 2, 4, True, np.float32),
    ]
    for tag, bn, qn, shape, nlev, em, disc, dt in cases3d:
        X = rs.rand(*shape).astype(dt)
        t = dn.Transform3d(bn, qn, ext_mode=em)
        p = t.forward(X, nlevels=nlev, include_scale=True, discard_level_1=disc)
        pre = "t3/%s/" % tag
        out[pre + "meta"] = np.array([bn, qn, str(nlev), str(em), str(int(disc))])
        out[pre + "X"] = X
        out[pre + "Yl"] = p.lowpass
        for i, (h, s) in enumerate(zip(p.highpasses, p.scales)):
            if h is not None:
                out[pre + "Yh%d" % i] = h.astype(np.complex64 if dt == np.float32 else np.complex128)
            out[pre + "Ys%d" % i] = s
        Z = t.inverse(p)
        # reference quirk: the discard_level_1 inverse comes back with axes 0 and 2
        # swapped (transform3d.py:452-454); stored exactly as the reference returned it.
        out[pre + "Z"] = Z.astype(dt)

    np.savez_compressed(os.path.join(HERE, "ref_outputs.npz"), **out)
    for f in ("inputs.npz", "verification_subset.npz", "ref_outputs.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
