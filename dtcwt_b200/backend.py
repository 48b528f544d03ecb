"""Registration with the reference package: ``dtcwt.push_backend('b200')``.

The reference has no public registration call; its registry is the module-level
dict ``dtcwt._AVAILABLE_BACKENDS`` (``dtcwt/__init__.py:29-48``) that
``push_backend`` indexes (``:112-116``).  ``register()`` adds one entry with the
four keys every backend provides.  See INTEGRATION.md for the one-line patch a
reference maintainer would add instead.
"""
from __future__ import annotations

BACKEND_NAME = "b200"


def backend_dict():
    from . import Pyramid, Transform1d, Transform2d, Transform3d
    return {"Transform1d": Transform1d, "Transform2d": Transform2d,
            "Transform3d": Transform3d, "Pyramid": Pyramid}


def register(dtcwt_module=None, name=BACKEND_NAME):
    """Make ``dtcwt.push_backend(name)`` select this backend.  Returns the ``dtcwt`` module."""
    if dtcwt_module is None:
        import dtcwt as dtcwt_module
    dtcwt_module._AVAILABLE_BACKENDS[name] = backend_dict()
    return dtcwt_module
