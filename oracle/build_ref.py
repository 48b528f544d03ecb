#!/usr/bin/env python
"""Recipe: install the UNMODIFIED reference (rjw57/dtcwt) into ``oracle/_ref/``.

TEST / MEASUREMENT INFRASTRUCTURE -- never imported by the product package.

The reference is pure Python (no build step of its own), so "building" it is a
``pip install --target`` of the checkout at ``/root/reference`` (or
``$DTCWT_REFERENCE``).  pip wants to write ``*.egg-info`` next to ``setup.py`` and
the checkout is read-only, so the install runs from a throw-away copy under
``/tmp``.  Nothing from the reference is committed: ``oracle/_ref/`` is listed
in ``.gitignore`` (not in ``.gpurunignore``), so it travels to the GPU box with
the snapshot exactly like our own built ``.so``.

Who uses it (all through ``oracle/refshim.py``, which adds the three NumPy-2
attribute shims of SURVEY.md appendix C before ``import dtcwt``):
  * ``bench.py --impl reference`` and the ``cpu_baseline`` leg: ``dtcwt.numpy`` timed
    on the box's host cores (``kind: "reference"``);
  * ``tests/test_backend_registry.py``: ``dtcwt.push_backend('b200')`` on the GPU box;
  * the parity check inside ``bench.py`` (full arrays at 4096 x 4096).

Usage:  python oracle/build_ref.py [--force]
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("DTCWT_REFERENCE", "/root/reference")
WHEELHOUSE = "/opt/wheelhouse"


def installed():
    return os.path.isfile(os.path.join(TARGET, "dtcwt", "__init__.py"))


def source_available():
    return os.path.isfile(os.path.join(SOURCE, "setup.py")) and os.path.isdir(os.path.join(SOURCE, "dtcwt"))


def build(force=False):
    """-> path of oracle/_ref, or None when there is neither an install nor a reference checkout."""
    if installed() and not force:
        return TARGET
    if not source_available():
        return TARGET if installed() else None
    tmp = tempfile.mkdtemp(prefix="dtcwt_ref_")
    try:
        copy = os.path.join(tmp, "src")
        shutil.copytree(SOURCE, copy, ignore=shutil.ignore_patterns(".git", "*.pyc", "__pycache__"))
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
               "--no-compile", "--target", TARGET, copy]
        if os.path.isdir(WHEELHOUSE):
            cmd[cmd.index("--target"):cmd.index("--target")] = ["--find-links", WHEELHOUSE]
        subprocess.check_call(cmd)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    if not installed():
        raise RuntimeError("pip reported success but %s/dtcwt is missing" % TARGET)
    return TARGET


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
