"""The C-ABI library loads without a GPU and exports every symbol include/dtcwt_b200.h declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from dtcwt_b200 import _lib

HEADER = os.path.join(ROOT, "include", "dtcwt_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dtcwt_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.EXPORTS)


def test_library_exports_every_symbol():
    if not os.path.isfile(_lib.LIB_PATH):
        import build
        try:
            build.build()
        except RuntimeError as e:
            pytest.skip(str(e))
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for sym in declared_symbols():
        assert hasattr(handle, sym), sym
    handle.dtcwt_b200_version.restype = ctypes.c_int
    assert handle.dtcwt_b200_version() == _lib.ABI_VERSION
    assert handle.dtcwt_b200_is_device_build() == 1
    handle.dtcwt_b200_error_string.restype = ctypes.c_char_p
    assert b"invalid" in handle.dtcwt_b200_error_string(-1)


def test_no_cpu_fallback():
    """Without a CUDA device the product path refuses to run (it must never fall back to the CPU)."""
    import numpy as np
    import torch
    import dtcwt_b200
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    import emu_seam
    emu_seam.install(None)
    with pytest.raises(RuntimeError):
        dtcwt_b200.Transform2d().forward(np.zeros((8, 8), np.float32), 1)


def test_emulator_cannot_masquerade(emulator_path):
    handle = ctypes.CDLL(emulator_path)
    assert handle.dtcwt_b200_is_device_build() == 0
    import emu_seam
    with pytest.raises(RuntimeError):
        emu_seam.install(_lib.LIB_PATH)   # a device build is refused as emulator
    emu_seam.install(None)


def test_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "dtcwt_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl")):
                src = open(os.path.join(dirpath, f)).read()
                assert "dtcwt_oracle" not in src and "refshim" not in src, f
                if f.endswith(".py"):      # the CPU seam lives in tests/ only (the kernel bodies keep their DTCWT_EMU build)
                    assert "emu_seam" not in src and "emulator" not in src.lower(), f


def test_graph_helper_needs_a_cuda_tensor():
    """dtcwt_b200.graph.Graphed is CUDA-graph capture: a host tensor is refused up front (no CPU path to fall back to)."""
    import pytest
    import torch
    import dtcwt_b200
    with pytest.raises(ValueError):
        dtcwt_b200.graph.Graphed(lambda x: x, torch.zeros(4, 4))
