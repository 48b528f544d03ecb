#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, share.

Usage: python tools/summarize_launches.py gpurun_out/launches.csv [--skip N] > profiles/xxx.md
(per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes)
"""
import csv
import re
import sys


def short(name):
    name = re.sub(r"\(.*$", "", name)
    m = re.search(r"(generic_1d_kernel<dtcwt::\w+<\w+>\s*>|dtcwt::\w+(<[^(]*>)?)", name)
    if m:
        return m.group(1)
    name = re.sub(r"^void\s+", "", name)
    return name[:70]


def main():
    path = sys.argv[1]
    skip = int(sys.argv[sys.argv.index("--skip") + 1]) if "--skip" in sys.argv else 0
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        rows.append((int(r["ID"]), short(r["Kernel Name"]), us, r["Grid Size"], r["Block Size"]))
    rows = [r for r in rows if r[0] >= skip]
    agg = {}
    for _, k, us, g, b in rows:
        t, n = agg.get(k, (0.0, 0))
        agg[k] = (t + us, n + 1)
    tot = sum(t for t, _ in agg.values())
    print("| kernel | launches | total us | avg us | share |")
    print("|---|---:|---:|---:|---:|")
    for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print("| `%s` | %d | %.1f | %.1f | %.1f%% |" % (k, n, t, t / n, 100 * t / tot))
    print("\ntotal %.1f us over %d launches (launch IDs >= %d)" % (tot, len(rows), skip))


if __name__ == "__main__":
    main()
