#!/usr/bin/env python
"""Bare pinned-memory H2D / D2H copy bandwidth: the roofline of bench.py's `e2e` number.

    python tools/exp/copy_bw.py [--gpus N] [--mib 1024] [--reps 10]

N processes (one per GPU, spawned here; no collective) each time cudaMemcpyAsync of one pinned buffer
H2D only, D2H only, and both directions at once on two streams, all ranks started together by a file
barrier.  Prints one JSON line: per-direction GB/s per GPU (min over ranks) and the aggregate.  The
e2e step of bench.py moves 1 GiB each way per 16-image chunk, so its copy roofline per GPU is
chunk_bytes / bidir GB/s."""
import argparse
import json
import os
import sys
import time


def worker(rank, world, mib, reps, tmp, q):
    import torch
    torch.cuda.set_device(rank)
    n = mib * 2 ** 20 // 4
    hin = torch.empty(n, dtype=torch.float32).pin_memory()
    hout = torch.empty(n, dtype=torch.float32).pin_memory()
    hin.fill_(1.0)
    din = torch.empty(n, dtype=torch.float32, device="cuda")
    dout = torch.ones(n, dtype=torch.float32, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def sync_all(tag):
        open(os.path.join(tmp, "%s.%d" % (tag, rank)), "w").close()
        while sum(os.path.exists(os.path.join(tmp, "%s.%d" % (tag, r))) for r in range(world)) < world:
            time.sleep(0.001)

    def timed(tag, h2d, d2h):
        for _ in range(2):
            if h2d:
                with torch.cuda.stream(s1):
                    din.copy_(hin, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    hout.copy_(dout, non_blocking=True)
        torch.cuda.synchronize()
        sync_all(tag)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    din.copy_(hin, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    hout.copy_(dout, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return reps * n * 4 / dt / 1e9          # GB/s per direction

    out = {"h2d": timed("a", True, False), "d2h": timed("b", False, True), "bidir_each": timed("c", True, True)}
    q.put((rank, out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--mib", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    import tempfile
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    tmp = tempfile.mkdtemp(prefix="copybw_")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, a.gpus, a.mib, a.reps, tmp, q)) for r in range(a.gpus)]
    for p in procs:
        p.start()
    res = dict(q.get() for _ in procs)
    for p in procs:
        p.join()
    line = {"n_gpus": a.gpus, "buffer_mib": a.mib, "reps": a.reps}
    for k in ("h2d", "d2h", "bidir_each"):
        vals = [res[r][k] for r in sorted(res)]
        line[k + "_gbs_per_gpu_min"] = round(min(vals), 2)
        line[k + "_gbs_aggregate"] = round(sum(vals), 2)
    try:
        line["host_cpus"] = len(os.sched_getaffinity(0))
    except AttributeError:
        pass
    print(json.dumps(line))


if __name__ == "__main__":
    main()
