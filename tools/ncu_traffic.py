#!/usr/bin/env python
"""DRAM traffic of each of our kernels from an `ncu --set full` report of `bench.py --images N`:
writes profiles/ncu_traffic.json = {abi symbol: {"dram_bytes_per_pixel": ..., ...}} where a pixel is one
pixel of the level-1 image (the q-shift entry is the level-2 launch, the largest).  bench.py scales it by
the pixels of its own launch to fill roofline.traffic.   usage: ncu_traffic.py REPORT N_IMAGES SIDE OUT.json"""
import csv, json, subprocess, sys

def symbol(name):
    if "fwds1_kernel" in name or ("fwd2d_kernel" in name and "SpecCol" in name): return "dtcwt_b200_fwd2d_level1_f32"
    if "invs1_kernel" in name or ("inv2d_kernel" in name and "SpecCol" in name): return "dtcwt_b200_inv2d_level1_f32"
    if "fwd2d_kernel" in name: return "dtcwt_b200_fwd2d_levelq_f32"
    if "inv2d_kernel" in name: return "dtcwt_b200_inv2d_levelq_f32"
    return None

def main(path, nimg, side, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, u = rows[0], rows[1]
    ix = {k: h.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")}
    def to_bytes(v, unit):
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    px = float(nimg) * float(side) ** 2
    best = {}
    for r in rows[2:]:
        s = symbol(r[ix["Kernel Name"]])
        if s is None: continue
        rd = to_bytes(r[ix["dram__bytes_read.sum"]], u[ix["dram__bytes_read.sum"]])
        wr = to_bytes(r[ix["dram__bytes_write.sum"]], u[ix["dram__bytes_write.sum"]])
        if s not in best or rd + wr > best[s]["dram_bytes"]:
            best[s] = {"kernel": r[ix["Kernel Name"]][:160], "dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr,
                       "time_us_under_ncu": float(r[ix["gpu__time_duration.sum"]]),
                       "dram_bytes_per_pixel": (rd + wr) / px, "images": int(nimg), "side": int(side), "report": path}
    json.dump(best, open(out, "w"), indent=1, sort_keys=True)
    for k, v in sorted(best.items()): print(k, "%.2f B/pixel" % v["dram_bytes_per_pixel"])

if __name__ == "__main__": main(*sys.argv[1:])
