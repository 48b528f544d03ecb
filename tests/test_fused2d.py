"""Fused per-level 2-D kernels (dtcwt_b200/csrc/fused2d.cuh) against the CPU oracle.

Runs on the host emulator of the kernel phases in the CPU suite and on the CUDA
library under ``-m gpu``.  Each case also checks that the fused entry points were
the ones launched (no silent use of the generic composition), and that fused and
generic paths agree with each other.
"""
import logging

import numpy as np
import pytest
import torch

import dtcwt_b200
import dtcwt_oracle as O
from dtcwt_b200 import _lib, _ops, coeffs
from util import REL_TOL, golden, rel_err

logging.disable(logging.WARNING)


class Launches(object):
    def __init__(self):
        self.names = []

    def __call__(self, symbol, thunk, launches=1):
        self.names.append(symbol)
        thunk()

    def __enter__(self):
        _lib.set_launch_hook(self)
        return self

    def __exit__(self, *exc):
        _lib.set_launch_hook(None)

    def only_fused(self):
        return self.names and all("2d_level" in n for n in self.names)


def npy(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def check_roundtrip(X, biort, qshift, nlevels, gain=None, expect_fused=True):
    xf = dtcwt_b200.Transform2d(biort, qshift)
    to = O.Transform2d(coeffs.biort(biort), coeffs.qshift(qshift))
    with Launches() as L:
        p = xf.forward(X, nlevels, include_scale=True)
        Z = npy(xf.inverse(p, gain))
    if expect_fused:
        assert L.only_fused(), L.names
        assert len(L.names) == 2 * nlevels
    po = to.forward(X, nlevels, include_scale=True)
    assert p.lowpass.dtype == np.float32 and rel_err(p.lowpass, po.lowpass) < REL_TOL
    for a, b in zip(p.highpasses, po.highpasses):
        assert a.dtype == np.complex64 and a.shape == b.shape
        assert rel_err(a, b) < REL_TOL
    for a, b in zip(p.scales, po.scales):
        assert a.shape == b.shape and rel_err(a, b) < REL_TOL
    Zo = to.inverse(po, gain)
    assert Z.shape == Zo.shape and rel_err(Z, Zo) < REL_TOL
    return p


@pytest.mark.parametrize("shape,nlevels", [
    ((64, 64), 2),        # one tile, every edge patched
    ((128, 192), 2),      # several tiles, exact multiples
    ((72, 100), 2),       # ragged last tile; level 2 input 72x100 -> 100 % 4 == 0, 72 % 4 == 0
    ((70, 90), 2),        # level-2 input not a multiple of 4 in either axis: pad + crop (transform2d.py:134-140)
    ((65, 131), 2),       # odd sizes: last row / column repeated (transform2d.py:86-94)
    ((160, 136), 3),      # three levels, level 3 input 40x34 (pad columns)
    ((32, 32), 1),        # smallest image the fused kernels accept
    ((64, 1056), 2),      # wide enough for interior (no symmetric-extension) tiles of every kernel
    ((840, 48), 1),       # tall: several runs of the streaming level-1 kernels, interior periods, ragged last run
    ((410, 300), 1),      # two column strips, last period of the run partly below the image
    ((1100, 72), 2),      # level-2 inverse streams 550 rows: two runs of the q-shift streaming kernel, 550 % 4 != 0 -> crop
    ((160, 1000), 3),     # wide: three column strips at level 2, ragged strips at level 3
])
def test_fused_vs_oracle_shapes(backend, shape, nlevels):
    rs = np.random.RandomState(shape[0] * 1000 + shape[1])
    X = rs.rand(*shape).astype(np.float32)
    gain = rs.rand(6, nlevels)
    check_roundtrip(X, "near_sym_b", "qshift_b", nlevels, gain)


@pytest.mark.parametrize("shape", [(64, 64), (840, 200), (410, 300), (65, 131)])
def test_streaming_forward_level1(backend, monkeypatch, shape):
    """The streaming level-1 forward kernel (stream2d.cuh FwdS1; the default forward is the tile kernel, which measured
    faster) is kept selectable with DTCWT_B200_FWD_STREAM=1 and must give the same pyramid."""
    rs = np.random.RandomState(shape[0] + shape[1])
    X = rs.rand(*shape).astype(np.float32)
    for biort, qshift in (("near_sym_b", "qshift_b"), ("near_sym_a", "qshift_a"), ("antonini", "qshift_06")):
        xf = dtcwt_b200.Transform2d(biort, qshift)
        p_tile = xf.forward(X, 1)
        monkeypatch.setenv("DTCWT_B200_FWD_STREAM", "1")
        with Launches() as L:
            p_stream = xf.forward(X, 1)
        monkeypatch.delenv("DTCWT_B200_FWD_STREAM")
        assert L.only_fused()
        po = O.Transform2d(coeffs.biort(biort), coeffs.qshift(qshift)).forward(X, 1)
        assert rel_err(p_stream.lowpass, po.lowpass) < REL_TOL and rel_err(p_tile.lowpass, po.lowpass) < REL_TOL
        assert rel_err(p_stream.highpasses[0], po.highpasses[0]) < REL_TOL
        assert rel_err(p_stream.highpasses[0], p_tile.highpasses[0]) < REL_TOL


@pytest.mark.parametrize("shape,nlevels", [((64, 64), 2), ((1100, 72), 2), ((160, 1000), 3), ((70, 90), 2)])
def test_streaming_inverse_qshift(backend, monkeypatch, shape, nlevels):
    """The streaming q-shift inverse (stream2d.cuh InvSq; the default is the tile kernel) is selected with
    DTCWT_B200_INV_STREAM=1 and must reconstruct like the oracle, gains and crops included."""
    rs = np.random.RandomState(shape[0] + 7 * shape[1])
    X = rs.rand(*shape).astype(np.float32)
    gain = rs.rand(6, nlevels)
    for biort, qshift in (("near_sym_b", "qshift_b"), ("near_sym_a", "qshift_a")):
        xf = dtcwt_b200.Transform2d(biort, qshift)
        to = O.Transform2d(coeffs.biort(biort), coeffs.qshift(qshift))
        p = xf.forward(X, nlevels)
        monkeypatch.setenv("DTCWT_B200_INV_STREAM", "1")
        with Launches() as L:
            Z = npy(xf.inverse(p, gain))
        monkeypatch.delenv("DTCWT_B200_INV_STREAM")
        assert L.only_fused()
        Zt = npy(xf.inverse(p, gain))
        Zo = to.inverse(to.forward(X, nlevels), gain)
        assert Z.shape == Zo.shape and rel_err(Z, Zo) < REL_TOL and rel_err(Z, Zt) < REL_TOL


@pytest.mark.parametrize("biort,qshift", [
    ("near_sym_a", "qshift_a"),      # library defaults: 5/7-tap level 1, 10-tap q-shift
    ("antonini", "qshift_06"),       # 9/7 taps zero-padded into the 13/19 instance; 10-tap q-shift
    ("legall", "qshift_d"),          # 5/3 taps; 18-tap q-shift
    ("near_sym_b", "qshift_b"),
])
def test_fused_wavelet_families(backend, biort, qshift):
    rs = np.random.RandomState(len(biort) + len(qshift))
    X = rs.rand(96, 80).astype(np.float32)
    check_roundtrip(X, biort, qshift, 2)


def test_fused_unsupported_falls_back_to_generic_kernels(backend):
    """qshift_32 has 32 taps (no fused instance), 24x24 is below the fused minimum: the generic CUDA kernels must produce
    the result instead."""
    rs = np.random.RandomState(3)
    X = rs.rand(64, 64).astype(np.float32)
    for biort, qshift in (("near_sym_b", "qshift_32"),):
        xf = dtcwt_b200.Transform2d(biort, qshift)
        to = O.Transform2d(coeffs.biort(biort), coeffs.qshift(qshift))
        with Launches() as L:
            p = xf.forward(X, 2)
        assert any("coldfilt" in n for n in L.names)
        po = to.forward(X, 2)
        assert rel_err(p.lowpass, po.lowpass) < REL_TOL
        assert rel_err(npy(xf.inverse(p)), to.inverse(po)) < REL_TOL
    with Launches() as L:
        dtcwt_b200.Transform2d("near_sym_b", "qshift_b").forward(X[:24, :24], 1)
    assert any("colfilter" in n for n in L.names)


def test_fused_bp_families(backend):
    """near_sym_b_bp / qshift_b_bp (6- and 12-tuples, reference transform2d.py:116-127, 145-157, 254-262, 279-292): every
    level is two fused launches -- the ordinary one and the band-pass one for sub-bands 1 and 4 -- nothing else."""
    rs = np.random.RandomState(29)
    X = rs.rand(2, 264, 256).astype(np.float32)
    xf = dtcwt_b200.Transform2d("near_sym_b_bp", "qshift_b_bp")
    to = O.Transform2d(coeffs.biort("near_sym_b_bp"), coeffs.qshift("qshift_b_bp"))
    gm = rs.rand(6, 3)
    with Launches() as L:
        p = xf.forward_channels(X, "nhw", 3)
        Z = npy(xf.inverse_channels(p, "nhw"))
        Zg = npy(xf.inverse_channels(p, "nhw", gm))
    assert L.only_fused(), L.names
    assert sum(n.endswith("_hh_f32") for n in L.names) == 9 and len(L.names) == 18
    for i in range(2):
        po = to.forward(X[i], 3)
        assert rel_err(p.lowpass[i], po.lowpass) < REL_TOL
        for a, b in zip(p.highpasses, po.highpasses):
            assert rel_err(a[i], b) < REL_TOL
        assert rel_err(Zg[i], to.inverse(po, gm)) < REL_TOL
        assert rel_err(Z[i], to.inverse(po)) < REL_TOL      # (the band-pass families are not a perfect-reconstruction set)


def test_fused_16_tap_qshift(backend):
    """qshift_c: 16 taps, m / 2 even -- colifilt's second index scheme (reference lowlevel.py:205-231) -- runs on the
    fused level kernels too."""
    rs = np.random.RandomState(23)
    X = rs.rand(2, 256, 264).astype(np.float32)          # level 3 works on 64 x 66 -> 32 x 33, still above the fused minimum
    xf = dtcwt_b200.Transform2d("near_sym_b", "qshift_c")
    to = O.Transform2d(coeffs.biort("near_sym_b"), coeffs.qshift("qshift_c"))
    with Launches() as L:
        p = xf.forward_channels(X, "nhw", 3)
        Z = npy(xf.inverse_channels(p, "nhw"))
    assert L.only_fused(), L.names
    for i in range(2):
        po = to.forward(X[i], 3)
        assert rel_err(p.lowpass[i], po.lowpass) < REL_TOL
        for a, b in zip(p.highpasses, po.highpasses):
            assert rel_err(a[i], b) < REL_TOL
    assert np.abs(Z - X).max() < 1e-5


def test_fused_batch_and_generic_agree(backend):
    rs = np.random.RandomState(17)
    X = rs.rand(3, 160, 136).astype(np.float32)
    xf = dtcwt_b200.Transform2d("near_sym_b", "qshift_b")
    gm = rs.rand(6, 3)
    with Launches() as L:
        pf = xf.forward_channels(X, "nhw", 3)
        Zf = npy(xf.inverse_channels(pf, "nhw", gm))
    assert L.only_fused() and len(L.names) == 6
    _ops.FUSED_ENABLED = False
    try:
        pg = xf.forward_channels(X, "nhw", 3)
        Zg = npy(xf.inverse_channels(pg, "nhw", gm))
    finally:
        _ops.FUSED_ENABLED = True
    assert rel_err(pf.lowpass, pg.lowpass) < REL_TOL
    for a, b in zip(pf.highpasses, pg.highpasses):
        assert rel_err(a, b) < REL_TOL
    assert rel_err(Zf, Zg) < REL_TOL
    # every image of the batch equals its single-image transform bit for bit
    p1 = xf.forward(X[1], 3)
    assert np.array_equal(p1.lowpass, pf.lowpass[1])
    assert np.array_equal(p1.highpasses[0], pf.highpasses[0][1])


def test_fused_mandrill_config2(backend):
    """BASELINE config 2 through the fused kernels: 512x512 mandrill, 4 levels, near_sym_b + qshift_b."""
    mandrill = golden("inputs")["mandrill"]
    p = check_roundtrip(mandrill, "near_sym_b", "qshift_b", 4)
    Z = npy(dtcwt_b200.Transform2d("near_sym_b", "qshift_b").inverse(p))
    assert rel_err(Z, mandrill) < REL_TOL


def test_fused_accepts_reference_layout_pyramid(backend):
    """The inverse takes the reference's interleaved (h, w, 6) NumPy pyramid as well as ours."""
    rs = np.random.RandomState(23)
    X = rs.rand(64, 96).astype(np.float32)
    to = O.Transform2d(coeffs.biort("near_sym_a"), coeffs.qshift("qshift_a"))
    po = to.forward(X, 2)
    Z = npy(dtcwt_b200.Transform2d("near_sym_a", "qshift_a").inverse(po))
    assert rel_err(Z, to.inverse(po)) < REL_TOL


@pytest.mark.gpu
def test_tma_and_plain_staging_agree():
    """GPU only: a width that is a multiple of 4 is staged by TMA, 4k+2 by plain loads; both match the oracle
    (checked above); here the same image is run through both and compared bit for bit via a column-padded copy."""
    rs = np.random.RandomState(31)
    X = rs.rand(2, 128, 256).astype(np.float32)
    xf = dtcwt_b200.Transform2d("near_sym_b", "qshift_b")
    import emu_seam
    emu_seam.install(None)
    p_tma = xf.forward_channels(torch.from_numpy(X).cuda(), "nhw", 2)
    # misalign the base pointer by one float so that TMA cannot be used
    buf = torch.empty(X.size + 1, dtype=torch.float32, device="cuda")
    Xm = buf[1:].view(2, 128, 256)
    Xm.copy_(torch.from_numpy(X))
    p_ld = xf.forward_channels(Xm, "nhw", 2)
    assert torch.equal(p_tma.lowpass_t, p_ld.lowpass_t)
    for a, b in zip(p_tma.highpasses_t, p_ld.highpasses_t):
        assert torch.equal(a, b)


def test_staged_level1_inverse_matches(backend, monkeypatch):
    """The bulk-copy staged level-1 inverse (InvS1T, opt-in: DTCWT_B200_INV_STAGED=1) gives the same result as the
    default per-thread-load kernel, on interior strips, image borders (mirrored quad columns) and short last runs."""
    rs = np.random.RandomState(41)
    for shape, names in (((2, 96, 520), ("near_sym_b", "qshift_b")), ((1, 200, 72), ("near_sym_a", "qshift_a")),
                         ((1, 64, 256), ("antonini", "qshift_b"))):
        X = rs.rand(*shape).astype(np.float32)
        xf = dtcwt_b200.Transform2d(*names)
        p = xf.forward_channels(X, "nhw", 1)
        monkeypatch.setenv("DTCWT_B200_INV_STAGED", "0")
        Z0 = npy(xf.inverse_channels(p, "nhw", np.array([[1.0], [0.5], [2.0], [1.5], [0.25], [3.0]])))
        monkeypatch.setenv("DTCWT_B200_INV_STAGED", "1")
        Z1 = npy(xf.inverse_channels(p, "nhw", np.array([[1.0], [0.5], [2.0], [1.5], [0.25], [3.0]])))
        monkeypatch.delenv("DTCWT_B200_INV_STAGED")
        assert np.abs(Z0 - Z1).max() < 1e-6 * np.abs(Z0).max()
        to = O.Transform2d(coeffs.biort(names[0]), coeffs.qshift(names[1]))
        for i in range(shape[0]):
            po = to.forward(X[i], 1)
            assert rel_err(Z1[i], to.inverse(po, np.array([[1.0], [0.5], [2.0], [1.5], [0.25], [3.0]]))) < REL_TOL


def test_symmetric_sum_column_pass_matches(backend, monkeypatch):
    """Level-1 forward with the shared symmetric sums in the column pass (kFwdSym, DTCWT_B200_FWD_SYM=1): same results
    as the scatter form to rounding, oracle parity at the stated tolerance; odd sizes and image borders included."""
    rs = np.random.RandomState(43)
    for shape in ((2, 96, 200), (1, 67, 131)):
        X = rs.rand(*shape).astype(np.float32)
        xf = dtcwt_b200.Transform2d("near_sym_b", "qshift_b")
        monkeypatch.setenv("DTCWT_B200_FWD_SYM", "0")
        p0 = xf.forward_channels(X, "nhw", 1)
        monkeypatch.setenv("DTCWT_B200_FWD_SYM", "1")
        p1 = xf.forward_channels(X, "nhw", 1)
        monkeypatch.delenv("DTCWT_B200_FWD_SYM")
        assert rel_err(p1.lowpass, p0.lowpass) < 1e-6 and rel_err(p1.highpasses[0], p0.highpasses[0]) < 1e-6
        to = O.Transform2d(coeffs.biort("near_sym_b"), coeffs.qshift("qshift_b"))
        for i in range(shape[0]):
            po = to.forward(X[i], 1)
            assert rel_err(p1.lowpass[i], po.lowpass) < REL_TOL and rel_err(p1.highpasses[0][i], po.highpasses[0]) < REL_TOL


@pytest.mark.parametrize("mode", ["-1", "100"])
@pytest.mark.parametrize("shape,nlevels", [((3, 96, 128), 3), ((5, 130, 150), 2), ((2, 65, 131), 4)])
def test_chained_levels_match_per_level_launches(backend, monkeypatch, mode, shape, nlevels):
    """Levels 1 and 2 chained chunk by chunk through the L2-resident scratch (dtcwt_b200_fwd2d_level12_f32 /
    inv2d_level21_f32) give bit-identical results to one launch per level: same kernels, only the order changes.
    Covers a chunk that does not divide the batch, the level-2 edge padding (130 % 4 != 0) and odd sizes."""
    rng = np.random.RandomState(5)
    X = rng.rand(*shape).astype(np.float32)
    gain = rng.rand(6, nlevels)
    xf = dtcwt_b200.Transform2d("near_sym_b", "qshift_b")
    monkeypatch.setenv("DTCWT_B200_CHAIN", "0")
    p0 = xf.forward_channels(X, "nhw", nlevels=nlevels)
    Z0 = npy(xf.inverse_channels(p0, "nhw", gain_mask=gain))
    ref = [npy(p0.lowpass_t)] + [npy(h) for h in p0.highpasses_t]
    monkeypatch.setenv("DTCWT_B200_CHAIN", mode)
    monkeypatch.setenv("DTCWT_B200_CHAIN_MIN_PIX", "1")
    per = 4 * (shape[1] + shape[1] % 2) * (shape[2] + shape[2] % 2)
    monkeypatch.setenv("DTCWT_B200_CHAIN_MB", "1")               # 1 MiB budget: 2 .. 21 images per chunk
    assert _ops.chain_chunk(shape[0], shape[1] + shape[1] % 2, shape[2] + shape[2] % 2) == min(shape[0], (1 << 20) // per)
    with Launches() as L:
        p1 = xf.forward_channels(X, "nhw", nlevels=nlevels)
        Z1 = npy(xf.inverse_channels(p1, "nhw", gain_mask=gain))
    assert "dtcwt_b200_fwd2d_level12_f32" in L.names and "dtcwt_b200_inv2d_level21_f32" in L.names, L.names
    assert "dtcwt_b200_fwd2d_level1_f32" not in L.names and "dtcwt_b200_inv2d_level1_f32" not in L.names
    got = [npy(p1.lowpass_t)] + [npy(h) for h in p1.highpasses_t]
    for a, b in zip(got, ref):
        assert a.shape == b.shape and np.array_equal(a, b)
    assert np.array_equal(Z0, Z1)


@pytest.mark.parametrize("variant", ["1", "2"])
def test_row_pair_inverse_qshift_matches(backend, monkeypatch, variant):
    """The q-shift inverse tile kernel with its row pass on interleaved row pairs (opt-in: DTCWT_B200_INVQ_VARIANT=1 runtime
    taps, 2 baked qshift_b immediates; 2-D and the 3-D slice mode) agrees with the default single-row row pass to rounding --
    cropped levels (130 % 4 != 0), image borders and gains included -- and with the oracle at the stated tolerance."""
    rs = np.random.RandomState(47)
    gain = rs.rand(6, 2)
    for shape, names in (((2, 96, 200), ("near_sym_b", "qshift_b")), ((1, 130, 150), ("near_sym_b", "qshift_b")),
                         ((1, 72, 264), ("near_sym_a", "qshift_a"))):
        X = rs.rand(*shape).astype(np.float32)
        xf = dtcwt_b200.Transform2d(*names)
        p = xf.forward_channels(X, "nhw", 2)
        monkeypatch.setenv("DTCWT_B200_INVQ_VARIANT", "0")
        Z0 = npy(xf.inverse_channels(p, "nhw", gain))
        monkeypatch.setenv("DTCWT_B200_INVQ_VARIANT", variant)
        with Launches() as L:
            Z1 = npy(xf.inverse_channels(p, "nhw", gain))
        monkeypatch.delenv("DTCWT_B200_INVQ_VARIANT")
        assert L.only_fused(), L.names
        assert np.abs(Z0 - Z1).max() < 2e-6 * np.abs(Z0).max()
        to = O.Transform2d(coeffs.biort(names[0]), coeffs.qshift(names[1]))
        for i in range(shape[0]):
            assert rel_err(Z1[i], to.inverse(to.forward(X[i], 2), gain)) < REL_TOL
    V = rs.rand(1, 32, 40, 48).astype(np.float32)
    x3 = dtcwt_b200.Transform3d("near_sym_b", "qshift_b")
    p3 = x3.forward(V[0], 2)
    monkeypatch.setenv("DTCWT_B200_INVQ_VARIANT", "0")
    W0 = npy(x3.inverse(p3))
    monkeypatch.setenv("DTCWT_B200_INVQ_VARIANT", variant)
    W1 = npy(x3.inverse(p3))
    monkeypatch.delenv("DTCWT_B200_INVQ_VARIANT")
    assert np.abs(W0 - W1).max() < 2e-6 * np.abs(W0).max() and np.abs(W1 - V[0]).max() < 1e-5


@pytest.mark.parametrize("switch", ["DTCWT_B200_INVQ_ASYNC=4", "DTCWT_B200_INVQ_ASYNC=2", "DTCWT_B200_INV_ASYNC=0", "DTCWT_B200_INV_ASYNC=4",
                                    "DTCWT_B200_Z3_SPLIT=1", "DTCWT_B200_Z3_ASYNC=0", "DTCWT_B200_Z3_ASYNC=3", "DTCWT_B200_AXIS_NG=8",
                                    "DTCWT_B200_FWD_PREFETCH=0", "DTCWT_B200_FWD_PREFETCH=100", "DTCWT_B200_FWDQ_VARIANT=4",
                                    "DTCWT_B200_FWDQ_VARIANT=5", "DTCWT_B200_INV_UNI=0"])
def test_experiment_switches_agree_with_the_defaults(backend, monkeypatch, switch):
    """Every kernel variant that stays selectable by an environment switch (the measured-slower experiments of
    profiles/r3_01_experiments.md and the former defaults) gives the default kernels' results to rounding: a 2-D
    forward + inverse with gains and a cropped level, a 3-D forward + inverse, and a 3-D transform without level-1
    highpasses on a volume whose depth takes the long-axis pass."""
    rs = np.random.RandomState(53)
    X = rs.rand(2, 130, 200).astype(np.float32)
    gain = rs.rand(6, 2)
    V = rs.rand(32, 40, 48).astype(np.float32)
    W = rs.rand(128, 32, 40).astype(np.float32)
    x2 = dtcwt_b200.Transform2d("near_sym_b", "qshift_b")
    x3 = dtcwt_b200.Transform3d("near_sym_b", "qshift_b")

    def run():
        p = x2.forward_channels(X, "nhw", 2)
        out = [npy(p.lowpass_t)] + [npy(h) for h in p.highpasses_t] + [npy(x2.inverse_channels(p, "nhw", gain))]
        p3 = x3.forward(V, 2)
        out += [npy(p3.lowpass)] + [npy(h) for h in p3.highpasses] + [npy(x3.inverse(p3))]
        p4 = x3.forward(W, 2, discard_level_1=True)
        out += [npy(p4.lowpass), npy(p4.highpasses[1]), npy(x3.inverse(p4))]
        return out

    ref = run()
    name, value = switch.split("=")
    monkeypatch.setenv(name, value)
    got = run()
    monkeypatch.delenv(name)
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        assert a.shape == b.shape
        assert np.abs(a - b).max() <= 3e-6 * max(np.abs(b).max(), 1e-30)
