"""dtcwt.push_backend('b200') -- the drop-in boundary (reference dtcwt/__init__.py:29-48, 97-143;
semantics from the reference's tests/test_switchbackends.py)."""
import numpy as np
import pytest

import dtcwt_b200


@pytest.fixture()
def dtcwt():
    import refshim
    if not refshim.available():
        pytest.skip("reference checkout not present")
    mod = refshim.load()
    yield mod
    while len(mod._BACKEND_STACK) > 1:
        mod.pop_backend()
    mod._AVAILABLE_BACKENDS.pop("b200", None)


def test_register_push_pop(dtcwt):
    with pytest.raises(ValueError):
        dtcwt.push_backend("b200")                 # unknown until registered
    dtcwt_b200.register(dtcwt)
    assert set(dtcwt._AVAILABLE_BACKENDS["b200"]) == set(dtcwt._AVAILABLE_BACKENDS["numpy"])
    dtcwt.push_backend("b200")
    assert dtcwt.backend_name == "b200"
    assert dtcwt.Transform2d is dtcwt_b200.Transform2d
    assert dtcwt.Transform1d is dtcwt_b200.Transform1d
    assert dtcwt.Transform3d is dtcwt_b200.Transform3d
    assert dtcwt.Pyramid is dtcwt_b200.Pyramid
    dtcwt.pop_backend()
    assert dtcwt.backend_name == "numpy"
    assert dtcwt.Transform2d is dtcwt.numpy.Transform2d
    with pytest.raises(IndexError):
        dtcwt.pop_backend()


def test_preserve_backend_stack(dtcwt):
    dtcwt_b200.register(dtcwt)
    with dtcwt.preserve_backend_stack():
        dtcwt.push_backend("b200")
        assert dtcwt.backend_name == "b200"
    assert dtcwt.backend_name == "numpy"


def test_same_results_through_the_registry(dtcwt, backend):
    """User code written against the reference gets the same numbers after push_backend('b200')."""
    X = np.random.RandomState(2).rand(32, 24).astype(np.float32)
    ref = dtcwt.Transform2d("near_sym_b", "qshift_b").forward(X, nlevels=3)
    dtcwt_b200.register(dtcwt)
    dtcwt.push_backend("b200")
    xf = dtcwt.Transform2d("near_sym_b", "qshift_b")
    ours = xf.forward(X, nlevels=3)
    assert np.abs(ours.lowpass - ref.lowpass).max() < 1e-5 * np.abs(ref.lowpass).max()
    for a, b in zip(ours.highpasses, ref.highpasses):
        assert a.shape == b.shape and np.abs(a - b).max() < 1e-5 * np.abs(b).max()
    # a pyramid made by the numpy backend is accepted by our inverse
    Z = xf.inverse(ref)
    assert np.abs(Z.cpu().numpy() - X).max() < 1e-5
    # and the reference's registration module can read our pyramid's NumPy views
    assert isinstance(ours.highpasses[2], np.ndarray)
