import os
import sys

import pytest

ROOT = os.path.normpath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # a GPU test on a machine without a GPU is a skip, never a silent CPU pass
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


# ----------------------------------------------------------------------------- compute backends for parity tests
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_LIB = os.path.join(EMU_DIR, "libdtcwt_b200_emu.so")


def _build_emulator():
    """g++ build of the kernel bodies for the host (tests/emu/emu.cpp); rebuilt when sources change."""
    import glob
    import shutil
    import subprocess
    srcs = [os.path.join(EMU_DIR, "emu.cpp")] + glob.glob(os.path.join(ROOT, "dtcwt_b200", "csrc", "*")) \
        + [os.path.join(ROOT, "include", "dtcwt_b200.h")]
    if os.path.isfile(EMU_LIB) and all(os.path.getmtime(EMU_LIB) >= os.path.getmtime(s) for s in srcs):
        return EMU_LIB
    if shutil.which("g++") is None:
        return None
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", EMU_LIB,
                           os.path.join(EMU_DIR, "emu.cpp")])
    return EMU_LIB


@pytest.fixture(scope="session")
def emulator_path():
    p = _build_emulator()
    if p is None:
        pytest.skip("g++ not available to build the kernel-logic emulator")
    return p


@pytest.fixture(params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def backend(request):
    """Run a parity test twice: on the host emulator of the kernel bodies (CPU suite) and on the
    real CUDA library (`-m gpu`).  Test code is identical; only the loaded library differs."""
    from dtcwt_b200 import _lib
    import emu_seam
    if request.param == "emu":
        emu_seam.install(request.getfixturevalue("emulator_path"))
        yield "emu"
        emu_seam.install(None)
    else:
        emu_seam.install(None)
        assert _lib.lib().dtcwt_b200_is_device_build() == 1
        yield "gpu"
