"""TEST INFRASTRUCTURE: run the package's host logic against tests/emu's HOST build of the kernel bodies.

The product package has exactly one compute path (the CUDA library on a CUDA device) and no switch
to select anything else.  The CPU test-suite exercises the host layer (shapes, padding rules, level
loops, layouts, exceptions) and every index map of the kernels by monkeypatching, from here:

  * ``dtcwt_b200._lib._LIB``        <- ctypes binding of tests/emu/libdtcwt_b200_emu.so
  * ``dtcwt_b200._ops.to_device``   <- keep tensors on the CPU

``install(None)`` restores both.  Nothing under ``dtcwt_b200/`` refers to this module.
"""
from dtcwt_b200 import _lib, _ops

_ORIG_TO_DEVICE = _ops.to_device


def _to_cpu(t):
    return t.cpu()


def installed():
    return _ops.to_device is _to_cpu


def install(path):
    """path of the emulator library, or None to restore the product behaviour."""
    if path is None:
        if installed():
            _lib._LIB = None
        _ops.to_device = _ORIG_TO_DEVICE
        return
    loaded = _lib._bind(path, check_version=True)
    if loaded.dtcwt_b200_is_device_build():
        raise RuntimeError("refusing to install a device build as the emulator")
    _lib._LIB = loaded
    _ops.to_device = _to_cpu
