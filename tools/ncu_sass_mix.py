#!/usr/bin/env python
"""Opcode mix, stall reasons and hottest SASS lines of one kernel of an .ncu-rep captured with --import-source on.
usage: ncu_sass_mix.py REPORT LAUNCH_INDEX [TOP_N]"""
import collections
import csv
import subprocess
import sys


def main(path, skip, top=14):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(skip),
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    print(rows[0][1][:160])
    hdr = rows[1]
    ia, isamp, iexe = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    data = [r for r in rows[2:] if len(r) == len(hdr) and r[isamp].isdigit()]
    tot_s = sum(int(r[isamp]) for r in data) or 1
    tot_e = sum(int(r[iexe]) for r in data) or 1
    print("SASS lines %d, samples %d, warp instructions %d" % (len(data), tot_s, tot_e))
    mix, smp = collections.Counter(), collections.Counter()
    for r in data:
        t = r[ia].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        mix[op] += int(r[iexe])
        smp[op] += int(r[isamp])
    for op, c in mix.most_common(int(top)):
        print("  %-10s executed %5.1f%%   stall samples %5.1f%%" % (op, 100.0 * c / tot_e, 100.0 * smp[op] / tot_s))
    st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    for r in data:
        for h in st:
            tot[h] += int(r[hdr.index(h)])
    s = sum(tot.values()) or 1
    print("  stalls: " + ", ".join("%s %.1f%%" % (h[6:], 100.0 * c / s) for h, c in tot.most_common(8)))


if __name__ == "__main__":
    main(*sys.argv[1:])
