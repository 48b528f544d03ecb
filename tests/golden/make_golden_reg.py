#!/usr/bin/env python
"""Golden fixtures for the registration / re-sampling rows, produced by the UNMODIFIED reference
(``dtcwt.registration``, ``dtcwt.sampling`` through ``oracle/refshim.py``) on seeded synthetic frames.

Writes ``tests/golden/reg_outputs.npz``: the frame pair, ``estimatereg`` of their 5-level pyramids, the Q~ matrices of
level 3, a warped frame, a velocity field, and sample / rescale / sample_highpass / upsample results for every method.
Usage:  python tests/golden/make_golden_reg.py
"""
import logging
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refshim  # noqa: E402
from util import reg_frames  # noqa: E402


def main():
    logging.disable(logging.WARNING)
    d = refshim.load()
    reg = refshim.load_registration()
    import dtcwt.sampling as S
    f1, f2 = reg_frames()
    xf = d.numpy.Transform2d()
    p1, p2 = xf.forward(f1, nlevels=5), xf.forward(f2, nlevels=5)
    out = {"f1": f1.astype(np.float32), "f2": f2.astype(np.float32)}
    avecs = reg.estimatereg(p1, p2)
    out["avecs"] = avecs
    out["qt3"] = reg.qtildematrices(p1, p2, [3])[0]
    out["warp_bilinear"] = reg.warp(f1, avecs, method="bilinear")
    vx, vy = reg.velocityfield(avecs, (40, 56), method="bilinear")
    out["vx"], out["vy"] = vx, vy
    out["warphp2"] = reg.warphighpass(p1.highpasses[2], avecs, method="bilinear")
    rs = np.random.RandomState(77)
    xs = rs.uniform(-4, f1.shape[1] + 3, size=(9, 11))
    ys = rs.uniform(-4, f1.shape[0] + 3, size=(9, 11))
    out["xs"], out["ys"] = xs, ys
    hp = p1.highpasses[1]
    hx = rs.uniform(-2, hp.shape[1] + 1, size=(7, 5))
    hy = rs.uniform(-2, hp.shape[0] + 1, size=(7, 5))
    out["hx"], out["hy"] = hx, hy
    for m in ("nearest", "bilinear", "lanczos"):
        out["sample/" + m] = S.sample(f1, xs, ys, m)
        out["rescale/" + m] = S.rescale(f1, (37, 53), m)
        out["sample_highpass/" + m] = S.sample_highpass(hp, hx, hy, m)
        out["rescale_highpass/" + m] = S.rescale_highpass(hp, (30, 21), m)
        out["upsample/" + m] = S.upsample(f1[:24, :20], m)
        out["upsample_highpass/" + m] = S.upsample_highpass(hp[:10, :12], m)
    np.savez_compressed(os.path.join(HERE, "reg_outputs.npz"), **out)
    print("wrote reg_outputs.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
