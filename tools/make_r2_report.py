#!/usr/bin/env python
"""Assemble profiles/<PREFIX>_end_of_round.md from the files a `tools/gpu/r2z.sh` / `r3z.sh` run left in gpurun_out/ (bench lines,
launch lists, ncu reports or their on-box summaries) and copy the evidence next to it.
usage: python tools/make_r2_report.py [TAG [PREFIX]]   (defaults: r2z r2_04)"""
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r2z"
PREFIX = sys.argv[2] if len(sys.argv) > 2 else "r2_04"


def load(name):
    try:
        with open(os.path.join(G, name)) as f:
            return json.load(f)
    except Exception:
        return None


def run(cmd):
    return subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT).stdout


def main():
    out = ["# Round 2, end of round — measured state (%s)" % TAG, "",
           "All numbers: 1x B200 of this pool unless stated, CUDA events, `tools/gpu/%s.sh` (smoke, `pytest -m gpu`, compute-sanitizer, "
           "the three bench lines with CPU baseline and e2e, launch lists, `ncu --set full`)." % TAG, ""]
    copies = {"bench_%s.json" % TAG: PREFIX + "_bench_2d.json", "bench_%s_3d.json" % TAG: PREFIX + "_bench_3d.json",
              "bench_%s_reg.json" % TAG: PREFIX + "_bench_reg.json", "bench_%s_defaults.json" % TAG: PREFIX + "_bench_near_sym_a_qshift_a.json",
              "bench_%s_bp.json" % TAG: PREFIX + "_bench_near_sym_b_bp.json", "bench_%s_ref.json" % TAG: PREFIX + "_reference_arm_2d.json",
              "launches_%s.csv" % TAG: PREFIX + "_launches_2d.csv", "launches_%s_3d.csv" % TAG: PREFIX + "_launches_3d.csv",
              "sanitizer_memcheck_%s.log" % TAG: PREFIX + "_sanitizer_memcheck.log", "sanitizer_racecheck_%s.log" % TAG: PREFIX + "_sanitizer_racecheck.log"}
    for src, dst in copies.items():
        if os.path.isfile(os.path.join(G, src)):
            shutil.copy(os.path.join(G, src), os.path.join(P, dst))
    for key, title in (("bench_%s.json" % TAG, "2-D headline (`python bench.py`, 16 x 4096² fp32 per step, near_sym_b + qshift_b)"),
                       ("bench_%s_3d.json" % TAG, "3-D, config 4 (`python bench.py --workload 3d`, 8 x 256³ per step)"),
                       ("bench_%s_reg.json" % TAG, "Registration, config 5 (`python bench.py --workload reg`, 32 pairs of 1080 x 1920 per step)")):
        d = load(key)
        if not d:
            continue
        out += ["## " + title, "", "```json", json.dumps(d), "```", ""]
        r = d.get("roofline") or {}
        e = d.get("e2e") or {}
        c = d.get("cpu_baseline") or {}
        out += ["* value **%s %s** (%.3f ms per step), clocks %s MHz, reasons %s" % (d["value"], d["unit"], d["ms_per_step"],
                                                                                      d["clocks"]["sm_mhz"], d["clocks"]["reasons"])]
        if r.get("kernels_ms_per_step"):
            out += ["* per entry point (ms per step): " + ", ".join("`%s` %.3f" % (k.replace("dtcwt_b200_", ""), v)
                                                                     for k, v in r["kernels_ms_per_step"].items())]
        if r.get("frac") is not None:
            out += ["* roofline: `%s` %.1f GB/s = **%.3f** of the measured %.1f GB/s; whole step %s" % (
                r.get("kernel", "whole step"), r["achieved"], r["frac"], r["peak"], r.get("whole_step_frac"))]
        if e:
            cr = e.get("copy_roofline") or {}
            out += ["* e2e from pinned host memory: **%s %s** (%.2f ms per step)%s" % (
                e["value"], e["unit"], e["ms_per_step"], ", %.3f of the copy roofline (%.1f GB/s each way)" % (
                    cr["frac"], cr["bidir_gbs_per_gpu"]) if cr else "")]
        if c:
            out += ["* reference on the host cores (`%s`, %d cores): **%s %s** -- %s" % (c["kind"], c["cores"], c["value"], c["unit"], c["sample"])]
        if d.get("parity"):
            out += ["* parity: `%s`" % json.dumps(d["parity"])]
        out += [""]
    for key, title in (("bench_%s_defaults.json" % TAG, "Library-default wavelets (near_sym_a + qshift_a)"),
                       ("bench_%s_bp.json" % TAG, "Band-pass families (near_sym_b_bp + qshift_b_bp), two fused launches per level")):
        d = load(key)
        if d:
            out += ["## " + title, "", "%s %s, %.3f ms per step; per entry point: %s" % (
                d["value"], d["unit"], d["ms_per_step"], json.dumps(d["roofline"]["kernels_ms_per_step"])), ""]
    lat = os.path.join(G, "config2_latency_%s.txt" % TAG)
    if os.path.isfile(lat):
        out += ["## Config 2 latency", "", "```", open(lat).read().strip().splitlines()[-1], "```", ""]
    for tag2, title in (("", "2-D"), ("_3d", "3-D")):
        csvf = os.path.join(G, "launches_%s%s.csv" % (TAG, tag2))
        if os.path.isfile(csvf):
            out += ["## Launch list, %s (`ncu --metrics gpu__time_duration.sum`, 4 images / volumes, cold cache, serialised)" % title, "",
                    run([sys.executable, "tools/summarize_launches.py", csvf]).strip(), ""]
        rep = os.path.join(G, "prof_%s%s.ncu-rep" % (TAG, tag2))
        summ = os.path.join(G, "ncu_summary_%s%s.txt" % (TAG, tag2))
        if os.path.isfile(summ):          # summarised on the GPU box (the report itself may not have travelled: 64 MiB limit)
            out += ["## `ncu --set full`, one launch per kernel, %s" % title, "", "```", open(summ).read().strip(), "```", ""]
        elif os.path.isfile(rep):
            out += ["## `ncu --set full`, one launch per kernel, %s" % title, "", "```", run([sys.executable, "tools/ncu_summary.py", rep]).strip(), "```", ""]
    rep = os.path.join(G, "prof_%s.ncu-rep" % TAG)
    if os.path.isfile(rep):
        run([sys.executable, "tools/ncu_traffic.py", rep, "4", "4096", os.path.join(P, "ncu_traffic.json")])
        out += ["DRAM bytes per launch (`profiles/ncu_traffic.json`):", "", "```",
                run([sys.executable, "tools/ncu_traffic.py", rep, "4", "4096", os.path.join(P, "ncu_traffic.json")]).strip(), "```", ""]
    for name in ("memcheck", "racecheck"):
        f = os.path.join(G, "sanitizer_%s_%s.log" % (name, TAG))
        if os.path.isfile(f):
            tail = [l for l in open(f).read().splitlines() if l.strip()][-3:]
            out += ["## compute-sanitizer --tool %s (subset of the GPU tests)" % name, "", "```"] + tail + ["```", ""]
    mix = os.path.join(G, "ncu_sass_mix_%s.txt" % TAG)
    if os.path.isfile(mix):
        out += ["## Executed-instruction mix and stall reasons per kernel (`tools/ncu_sass_mix.py`, source page of the 2-D report)", "", "```",
                open(mix).read().strip(), "```", ""]
    with open(os.path.join(P, PREFIX + "_end_of_round.md"), "w") as f:
        f.write("\n".join(out) + "\n")
    print("wrote profiles/%s_end_of_round.md (%d lines)" % (PREFIX, len(out)))


if __name__ == "__main__":
    main()
