#!/usr/bin/env python
"""Generate dtcwt_b200/data/wavelets.npz from the reference's tap tables.

Provenance: the biort / qshift filter taps are published constants
(N. G. Kingsbury's DT-CWT filter designs).  The reference ships them as one
``.npz`` per family under ``dtcwt/data/`` (read by ``dtcwt/coeffs.py:13-25``).
This script repacks those arrays -- values untouched, float64 -- into ONE
archive keyed ``<family>/<tap name>`` so that the package is self-contained on
machines where the reference checkout does not exist (the GPU box).

Run in the build container only:  python tools/gen_wavelets.py
"""
import glob
import os
import sys

import numpy as np

REF = os.environ.get("DTCWT_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(__file__), "..", "dtcwt_b200", "data", "wavelets.npz")


def main():
    files = sorted(glob.glob(os.path.join(REF, "dtcwt", "data", "*.npz")))
    if not files:
        sys.exit("reference data directory not found under %s" % REF)
    table = {}
    for f in files:
        family = os.path.splitext(os.path.basename(f))[0]
        d = np.load(f)
        for k in d.files:
            if k.startswith("__") or k == "param":
                continue
            table["%s/%s" % (family, k)] = np.asarray(d[k], dtype=np.float64).reshape(-1)
    np.savez_compressed(OUT, **table)
    print("wrote %s: %d arrays from %d families" % (os.path.normpath(OUT), len(table), len(files)))


if __name__ == "__main__":
    main()
