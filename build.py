#!/usr/bin/env python
"""Build libdtcwt_b200.so (sm_100a) in-tree with nvcc.  Usage: python build.py [--force] [--ptxas-v]"""
import glob
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "dtcwt_b200", "csrc")
OUT = os.path.join(ROOT, "dtcwt_b200", "libdtcwt_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]


def _nvcc():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.isfile(c):
            return c
    return None


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


STAMP = OUT + ".flags"      # the flag set the library was built with (a diagnosis build must never pass for the default one)


def _flag_stamp(extra):
    return " ".join(NVCC_FLAGS + list(extra))


def up_to_date(extra=()):
    if not os.path.isfile(OUT):
        return False
    deps = glob.glob(os.path.join(CSRC, "*")) + [os.path.join(ROOT, "include", "dtcwt_b200.h")]
    if not all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return False
    if _nvcc() is None:
        return True              # GPU box without a toolkit: the shipped library is what there is
    try:
        with open(STAMP) as f:
            return f.read() == _flag_stamp(extra)
    except OSError:
        return False


def build(force=False, verbose=False, extra=()):
    if not force and up_to_date(extra):
        return OUT
    nvcc = _nvcc()
    if nvcc is None:
        if os.path.isfile(OUT):
            return OUT      # GPU box without a toolkit on PATH: use the library shipped in the snapshot
        raise RuntimeError("nvcc not found and %s does not exist" % OUT)
    # five translation units (DTCWT_PART, csrc/common.cuh) compiled in parallel, then one link
    objdir = os.path.join(ROOT, "build")
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f not in ("-shared",)]
    jobs, objs = [], []
    for src in sources():
        for part in (0, 1, 2, 3, 4):
            obj = os.path.join(objdir, "%s.part%d.o" % (os.path.splitext(os.path.basename(src))[0], part))
            cmd = [nvcc] + flags + list(extra) + ["-DDTCWT_PART=%d" % part, "-I", os.path.join(ROOT, "include"), "-c", "-o", obj, src]
            if verbose:
                print(" ".join(cmd))
            jobs.append((cmd, subprocess.Popen(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
            objs.append(obj)
    failed = False
    for cmd, proc in jobs:
        out, _ = proc.communicate()
        if out and (verbose or proc.returncode):
            print(out)
        failed = failed or proc.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    link = [nvcc] + NVCC_FLAGS + ["-o", OUT] + objs
    if verbose:
        print(" ".join(link))
    subprocess.check_call(link, cwd=ROOT)
    with open(STAMP, "w") as f:
        f.write(_flag_stamp(extra))
    return OUT


if __name__ == "__main__":
    extra = ["-Xptxas", "-v"] if "--ptxas-v" in sys.argv else []
    if "--diagnosis" in sys.argv:      # adds the memory-only / arithmetic-only builds of InvS1 (DTCWT_B200_INV_VARIANT=1|2)
        extra += ["-DDTCWT_DIAGNOSIS"]
    print(build(force="--force" in sys.argv or bool(extra), verbose=True, extra=extra))
