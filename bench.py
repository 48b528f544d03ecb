#!/usr/bin/env python
"""Headline benchmark: Mpixels/s of 2-D DT-CWT forward+inverse, 4 levels, batched 4096x4096 fp32.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = forward + inverse of one chunk of IMAGES x 4096 x 4096 fp32 synthetic images per
GPU (BASELINE.json configs[2]: the 1024-image job is run as such chunks because 1024 images
plus their pyramids do not fit one GPU; every image is independent so each rank simply owns
its own chunks -> weak scaling, no data-path collective).  Wavelets near_sym_b + qshift_b.
A "pixel" is one input pixel taken through forward AND inverse.

Prints ONE JSON line (rank 0).  See the task contract for the keys; in short
  value      device-resident throughput (CUDA events, max over ranks)
  e2e        same metric through the public API from pinned HOST buffers (H2D + D2H timed)
  roofline   dominant kernel: algorithmic bytes / its CUDA-event time vs MEASURED_PEAKS.json
  cpu_baseline  the unmodified reference (dtcwt.numpy from oracle/_ref) on the host cores, same image size
`--impl reference` times that reference alone (it is pure numpy; it has no GPU path).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mpixels/s 2D DT-CWT fwd+inv 4-level"
UNIT = "Mpixels/s"
SIDE = 4096
NLEVELS = 4
BIORT, QSHIFT = "near_sym_b", "qshift_b"
ALGO_BYTES_PER_PIXEL = 40.0     # fwd: read 4 + write 16; inv: read 16 + write 4
# algorithmic bytes per level-1 pixel and STEP of each fused entry point (DESIGN.md "kernels"); the q-shift entry
# points are launched once per level 2..4 (8 + 2 + 0.5 B/pixel), so their per-launch figure is the per-step one
# divided by the launches per step
KERNEL_ALGO_BYTES = {
    "dtcwt_b200_fwd2d_level1_f32": 20.0,   # read X 4, write LoLo 4 + Yh 12
    "dtcwt_b200_inv2d_level1_f32": 20.0,   # read LoLo 4 + Yh 12, write Z 4
    "dtcwt_b200_fwd2d_levelq_f32": 10.5,   # levels 2..4: read 4 + write LoLo 1 + Yh 3 per input pixel of the level
    "dtcwt_b200_inv2d_levelq_f32": 10.5,
}


def measured_traffic(sym, pixels):
    """dram__bytes_read + dram__bytes_write of one launch from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json, bytes per level-1 pixel), scaled to this run's launch; None if absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return round(json.load(f)[sym]["dram_bytes_per_pixel"] * pixels / 1e9, 4)
    except Exception:
        return None


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images", type=int, default=16, help="images per GPU per step")
    ap.add_argument("--pool", type=int, default=3, help="distinct input chunks cycled through")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--side", type=int, default=SIDE, help=argparse.SUPPRESS)   # debugging only
    ap.add_argument("--biort", default=BIORT, help="level-1 family (default: the north-star pair near_sym_b/qshift_b)")
    ap.add_argument("--qshift", default=QSHIFT)
    ap.add_argument("--workload", default="2d", choices=["2d", "3d", "reg"],
                    help="2d (default) = the headline metric, BASELINE configs[2]; 3d = configs[3] (256^3 volumes, 3 levels, "
                         "discard_level_1, fused 3-D levels), Mvoxels/s; reg = configs[4] (estimatereg on 1080p frame pairs), pairs/s")
    ap.add_argument("--total-images", type=int, default=None,
                    help="2d only: STRONG scaling -- the job is this many images in all (BASELINE configs[2]: 1024), "
                         "total/world per rank in chunks of --images; overrides --steps")
    return ap.parse_args()


# =============================================================================== CPU legs (the reference itself)
# The reference's implementation of this path is single-threaded numpy (dtcwt.numpy).  oracle/build_ref.py installs
# it, unmodified, into oracle/_ref (git-ignored, travels to the GPU box); oracle/refshim.py imports it under numpy 2.
# When that install is missing the oracle port (oracle/dtcwt_oracle.py, ~2.7x slower) stands in and `kind` says "port".
def cpu_kind():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refshim
    return "reference" if refshim.available() else "port"


def _cpu_transform():
    """-> (Transform2d-like object of the CPU leg, kind)"""
    for v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[v] = "1"
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refshim
    if refshim.available():
        import logging
        logging.disable(logging.WARNING)
        d = refshim.load()
        return d.numpy.Transform2d(BIORT, QSHIFT), "reference"      # dtcwt/numpy/transform2d.py:27,40,190
    import dtcwt_oracle as O
    from dtcwt_b200 import coeffs
    return O.Transform2d(coeffs.biort(BIORT), coeffs.qshift(QSHIFT)), "port"


def image_for_seed(side, seed):
    """Deterministic U[0,1) float32 image, the same on the CPU legs and (uploaded) on the GPU."""
    import numpy as np
    return np.random.RandomState(seed).random_sample((side, side)).astype(np.float32)


def _cpu_worker(job):
    """One image forward+inverse on one core; with `dump` the whole pyramid and the reconstruction are saved there."""
    side, seed, dump, names = job
    global BIORT, QSHIFT
    BIORT, QSHIFT = names
    import numpy as np
    xf, kind = _cpu_transform()
    X = image_for_seed(side, seed)
    t0 = time.perf_counter()
    p = xf.forward(X, NLEVELS)
    Z = xf.inverse(p)
    dt = time.perf_counter() - t0
    if dump:
        np.save(os.path.join(dump, "Yl.npy"), np.asarray(p.lowpass, dtype=np.float32))
        for l, h in enumerate(p.highpasses):
            np.save(os.path.join(dump, "Yh%d.npy" % l), np.asarray(h, dtype=np.complex64))
        np.save(os.path.join(dump, "Z.npy"), np.asarray(Z, dtype=np.float32))
    return dt, kind


def cpu_throughput(side, nworkers, dump_seed=None, dump_dir=None):
    """P worker processes, one `side` x `side` image each -> (Mpix/s aggregate, seconds, kind).  Worker 0 transforms
    the image of `dump_seed` and leaves its full pyramid in `dump_dir` for the parity check."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    names = (BIORT, QSHIFT)
    jobs = [(side, 1000 + i, None, names) for i in range(nworkers)]
    if dump_seed is not None:
        jobs[0] = (side, dump_seed, dump_dir, names)
    with ctx.Pool(nworkers) as pool:
        pool.map(_cpu_worker, [(64, 1, None, names)] * nworkers)      # import + first call outside the timed part
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, jobs)
        wall = time.perf_counter() - t0
    return nworkers * side * side / wall / 1e6, wall, res[0][1]


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def _ref_loop(wid, side, names, delay, q, go, stop):
    """Reference-arm worker: forward+inverse of 4096^2 images back to back until told to stop; one token per
    half image (after the forward, after the inverse) so the parent can place step boundaries finely."""
    global BIORT, QSHIFT
    BIORT, QSHIFT = names
    xf, kind = _cpu_transform()
    xf.inverse(xf.forward(image_for_seed(64, 1), NLEVELS))
    X = image_for_seed(side, 1000 + wid)
    q.put(("ready", wid, kind))
    go.wait()
    time.sleep(delay)
    while not stop.is_set():
        p = xf.forward(X, NLEVELS)
        q.put(("half", wid, kind))
        Z = xf.inverse(p)
        q.put(("half", wid, kind))
        del p, Z


def run_reference(args):
    """`--impl reference`: the UNMODIFIED reference (dtcwt.numpy.Transform2d(...).forward/.inverse from oracle/_ref) on
    `side` x `side` fp32 images, one single-threaded worker process per host core, all cores busy for the whole run.

    The workers transform images back to back; a "step" is the completion of `n` images under that full-machine load
    (n <= cores, chosen so that K + W steps end within a few minutes: one 4096^2 image costs 15-30 core-seconds).
    Worker starts are staggered so completions are spread evenly and step boundaries fall on single completions."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    steps = args.steps if args.steps is not None else 3
    warm = args.warmup if args.warmup is not None else 1
    cores = min(host_cores(), 64)
    side = args.side
    est_image_s = 30.0 * (side / 4096.0) ** 2                      # under load, conservative
    n = int(cores * (240.0 / max(1, steps + warm)) / est_image_s)
    n = max(1, min(cores, n))
    ctx = mp.get_context("spawn")
    q, go, stop = ctx.Queue(), ctx.Event(), ctx.Event()
    names = (BIORT, QSHIFT)
    procs = [ctx.Process(target=_ref_loop, args=(w, side, names, est_image_s * w / cores, q, go, stop), daemon=True)
             for w in range(cores)]
    for p in procs:
        p.start()
    kind = "port"
    for _ in range(cores):
        kind = q.get()[2]
    go.set()
    halves = 0
    t0 = None
    # the first image of every worker is ramp-up (staggered starts): it is never timed
    skip = max(2 * warm * n, 2 * cores)
    while True:
        q.get()
        halves += 1
        if halves == skip:
            t0 = time.perf_counter()
        if halves == skip + 2 * steps * n:
            t1 = time.perf_counter()
            break
    stop.set()
    for p in procs:                      # mid-image workers are ended by handle (the processes started above, nothing else)
        p.terminate()
    for p in procs:
        p.join(timeout=10)
    wall = t1 - t0
    value = steps * n * side * side / wall / 1e6
    sample = ("unmodified reference dtcwt.numpy (oracle/_ref)" if kind == "reference" else "numpy oracle port (oracle/_ref missing)") + \
        ": %d single-threaded worker processes transform %dx%d fp32 images back to back (forward+inverse, %d levels); " \
        "a step = %d image(s) completed under that all-cores load; %d image(s) timed, ramp-up excluded" % (
            cores, side, side, NLEVELS, n, steps * n)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": round(1e3 * wall / steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "images_per_step": n, "height": side, "width": side},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, side=None, images=None):
    side = side or args.side
    return {"workload": "2-D DT-CWT forward+inverse, %d levels, %s+%s, chunk of %d x %dx%d fp32 images per GPU per step "
                        "(BASELINE configs[2], the 1024-image batch run as chunks)" % (
                            NLEVELS, BIORT, QSHIFT, images or args.images, side, side),
            "images_per_gpu_per_step": images or args.images, "height": side, "width": side, "nlevels": NLEVELS,
            "biort": BIORT, "qshift": QSHIFT,
            "cache": "inputs larger than L2: each step reads a %.0f MiB chunk, %d chunks cycled" % (
                (images or args.images) * side * side * 4 / 2 ** 20, args.pool)}


# =============================================================================== clocks sampler
class ClockSampler(object):
    """SM clock, power and throttle reasons DURING the timed region: NVML polled every few ms from a thread
    (nvidia-smi takes longer to start than a 10-step timed region lasts); nvidia-smi -lms as the fallback."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    BITS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml, self.stop_flag = index, [], None, None, False
        self.sm, self.power, self.reasons, self.max_mhz = [], [], set(), None
        self.ready = False

    def prepare(self):
        """NVML initialisation takes ~100 ms the first time: do it before the timed region, not inside it."""
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.index
            if visible:
                ids = [v.strip() for v in visible.split(",") if v.strip()]
                if self.index < len(ids) and ids[self.index].isdigit():
                    phys = int(ids[self.index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
        except Exception:
            self.nvml = None
        self.ready = True

    def start(self):
        if not self.ready:
            self.prepare()
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                self.power.append(n.nvmlDeviceGetPowerUsage(self.handle) / 1e3)
                try:
                    mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for name, bit in self.BITS:
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None,
                    "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(sm),
                    "power_w_max": max(self.power) if self.power else None, "source": "nvml, 4 ms polling inside the timed region"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None, "source": "nvidia-smi -lms 100"}


# =============================================================================== GPU legs
class LaunchLog(object):
    """Counts C-ABI launches and (optionally) brackets each with CUDA events on the launch stream."""

    def __init__(self, torch):
        self.torch, self.count, self.events, self.timing, self.only = torch, 0, [], False, None

    def __call__(self, symbol, thunk, launches=1):
        self.count += launches
        if not self.timing or (self.only is not None and symbol != self.only):
            thunk()
            return
        s = self.torch.cuda.current_stream()
        e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
        e0.record(s)
        thunk()
        e1.record(s)
        self.events.append((symbol, e0, e1))

    def per_kernel_ms(self):
        out = {}
        for sym, e0, e1 in self.events:
            tot, n = out.get(sym, (0.0, 0))
            out[sym] = (tot + e0.elapsed_time(e1), n + 1)
        return out


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s; MEASURED_PEAKS.json absent)"


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import dtcwt_b200
    from dtcwt_b200 import _lib, coeffs, parallel

    steps = args.steps if args.steps is not None else 30
    warm = args.warmup if args.warmup is not None else 10     # the CPU baseline leg leaves the GPU idle for ~40 s: ramp the clocks
    if warm < 3:
        warm = 3                                   # timing rule: at least 3 warm-up steps
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback")
    rank, world, local = parallel.init("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    assert _lib.lib().dtcwt_b200_is_device_build() == 1
    scaling = "weak"
    if args.total_images:
        # the job as SURVEY 8(d)/(e) specify it: a fixed batch (1024 images) split into contiguous slices, one per
        # rank (parallel.shard_range), each slice processed in chunks of --images: total work is fixed -> strong scaling
        lo, hi = parallel.shard_range(args.total_images, rank, world)
        steps = max(1, (parallel.shard_range(args.total_images, 0, world)[1]) // args.images)      # rank 0 owns the largest slice
        scaling = "strong"

    # the one collective of the job: rank 0's filter taps to everyone (about 1 KB)
    biort = parallel.broadcast_taps(coeffs.biort(BIORT), 0, dev)
    qshift = parallel.broadcast_taps(coeffs.qshift(QSHIFT), 0, dev)
    xf = dtcwt_b200.Transform2d(biort, qshift)

    side, nimg = args.side, args.images
    # image 0 of chunk 0 on rank 0 is a seeded host image so the CPU oracle can check it
    pool = []
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    for c in range(args.pool):
        pool.append(torch.rand((nimg, side, side), dtype=torch.float32, device=dev, generator=g))
    parity_seed = 777
    pool[0][0].copy_(torch.from_numpy(image_for_seed(side, parity_seed)))

    def step(i):
        p = xf.forward_channels(pool[i % len(pool)], "nhw", nlevels=NLEVELS)
        return p, xf.inverse_channels(p, "nhw")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- parity + CPU baseline (rank 0, N=1)
    # Image 0 of chunk 0 goes through the CPU leg (the unmodified reference when oracle/_ref exists) in one of the
    # timed worker processes; EVERY array of its pyramid and the reconstruction are compared in full.
    parity, cpu_baseline = None, None
    if rank == 0:
        p, Z = step(0)
        torch.cuda.synchronize()
        recon = float((Z - pool[0]).abs().max())              # whole chunk
        parity = {"roundtrip_max_abs_err": recon}
        if world == 1 and not args.no_cpu_baseline:
            import shutil
            import tempfile
            got = {"Yl": p.lowpass_t[0].cpu().numpy(), "Z": Z[0].cpu().numpy()}
            for l in range(NLEVELS):
                got["Yh%d" % l] = p.highpasses_t[l][0].cpu().numpy()
            del p, Z
            cores = min(host_cores(), 64)
            dump = tempfile.mkdtemp(prefix="dtcwt_bench_")
            try:
                mpix, wall, kind = cpu_throughput(side, cores, dump_seed=parity_seed, dump_dir=dump)
                worst, per = 0.0, {}
                for k in sorted(got):
                    ref = np.load(os.path.join(dump, k + ".npy"))
                    assert ref.shape == got[k].shape, (k, ref.shape, got[k].shape)
                    per[k] = float(np.abs(got[k] - ref).max() / np.abs(ref).max())
                    worst = max(worst, per[k])
                    del ref
            finally:
                shutil.rmtree(dump, ignore_errors=True)
            what = "unmodified reference dtcwt.numpy (oracle/_ref)" if kind == "reference" else "numpy oracle port"
            cpu_baseline = {"value": round(mpix, 3), "unit": UNIT, "cores": cores, "kind": kind,
                            "sample": "%d worker processes x one %dx%d fp32 image each, forward+inverse, %s (%.1f s wall)" % (
                                cores, side, side, what, wall)}
            parity.update({"vs_%s_max_rel_err" % kind: worst, "per_array_rel_err": {k: float("%.3g" % v) for k, v in per.items()},
                           "checked": "image 0 of chunk 0, all arrays, all levels in full (Yl, Yh[0..%d], reconstruction) vs %s" % (
                               NLEVELS - 1, what), "tolerance": 1e-5, "ok": bool(worst < 1e-5 and recon < 1e-4)})
            del got
        else:
            del p, Z

    # ---------------------------------------------------------------- device-resident timing
    # Warm-up: every launch is bracketed with CUDA events (per-kernel breakdown, and which entry point dominates);
    # in the timed region only the dominant one is bracketed, so that event overhead stays out of the headline number.
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.prepare()
    log = LaunchLog(torch)
    _lib.set_launch_hook(log)
    log.timing = True
    for i in range(warm):
        step(i)
    barrier()
    per_warm = log.per_kernel_ms()
    breakdown = {k: round(v[0] / warm, 4) for k, v in sorted(per_warm.items())}
    log.only = max(per_warm.items(), key=lambda kv: kv[1][0])[0] if per_warm else None
    log.events = []
    if rank == 0:
        sampler.start()
    log.count, log.timing = 0, True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(steps):
        step(warm + i)
    e1.record()
    barrier()
    log.timing = False
    launches = log.count
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None
    pix_per_step = world * nimg * side * side
    value = pix_per_step * steps / (ms_total / 1e3) / 1e6

    # ---------------------------------------------------------------- roofline of the dominant kernel
    peak, peak_src = measured_peak_gbs()
    per = log.per_kernel_ms()
    roofline = None
    if per:
        top = max(per.items(), key=lambda kv: kv[1][0])
        sym, (tot_ms, n) = top
        avg_ms = tot_ms / n
        share = tot_ms / ms_total
        traffic = None
        if sym in KERNEL_ALGO_BYTES:
            per_step = n / steps
            algo = KERNEL_ALGO_BYTES[sym] * nimg * side * side / per_step
            note = "%s: %.1f B/pixel algorithmic per step, %d launch(es) per step; traffic = GB per launch from the " \
                   "committed ncu capture (profiles/ncu_traffic.json)" % (sym, KERNEL_ALGO_BYTES[sym], per_step)
            if per_step == 1:
                traffic = measured_traffic(sym, nimg * side * side)
        else:
            # generic (unfused) path: no single kernel dominates by design; report the whole step
            algo, avg_ms, sym = ALGO_BYTES_PER_PIXEL * nimg * side * side, ms_total / steps, "whole step (unfused generic kernels)"
            note = "whole fwd+inv step: 40 B/pixel compulsory traffic"
            share = 1.0
        ach = algo / (avg_ms / 1e3) / 1e9
        roofline = {"bound": "hbm", "kernel": sym, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(ach / peak, 4), "traffic": traffic, "traffic_unit": "GB per launch",
                    "algorithmic_gb_per_launch": round(algo / 1e9, 4), "peak_source": peak_src, "note": note,
                    "avg_launch_ms": round(avg_ms, 4), "share_of_step": round(share, 4),
                    "whole_step_frac": round(ALGO_BYTES_PER_PIXEL * nimg * side * side * steps / (ms_total / 1e3) / 1e9 / peak, 4),   # per GPU
                    "kernels_ms_per_step": breakdown,
                    "kernels_ms_per_step_note": "all entry points bracketed during the warm-up steps; the timed region "
                                                "brackets only the dominant one"}
    _lib.set_launch_hook(None)

    # ---------------------------------------------------------------- end-to-end from pinned host memory
    e2e = None
    if not args.no_e2e:
        def fn(x):
            return xf.inverse_channels(xf.forward_channels(x, "nhw", nlevels=NLEVELS), "nhw")
        e2e = run_e2e_generic(torch, dist, fn, pool, steps, world, dev, barrier, (nimg, side, side), UNIT)

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": round(ms_total / steps, 4), "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": dict(workload_config(args), **({"total_images": args.total_images,
                                                                                      "images_per_rank": steps * nimg} if args.total_images else {})),
            "clocks": clocks,
            "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline, "parity": parity,
            "hbm_frac_of_measured_peak": round(ALGO_BYTES_PER_PIXEL * value * 1e6 / 1e9 / (peak * world), 4),
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  NCCL and other native libraries print banners to fd 1
    ("NCCL version ..."), so fd 1 is pointed at stderr for the whole run and the line goes to a saved copy."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


# =============================================================================== config 4: 3-D volumes
METRIC_3D = "Mvoxels/s 3D DT-CWT fwd+inv 3-level discard_level_1"
ALGO_BYTES_PER_VOXEL = 16.0      # fwd: read 4 + write Yh 3.94 + Yl 0.06; inverse the same
SIDE_3D = 256
# algorithmic bytes per voxel of the 256^3 input moved by each fused entry point per step: level 1 reads and
# writes a full-size volume; levels 2 + 3 read 4 (1 + 1/8) and write their LLL + 28 complex channels
KERNEL_ALGO_BYTES_3D = {
    "dtcwt_b200_fwd3d_level1_lo_f32": 8.0, "dtcwt_b200_inv3d_level1_lo_f32": 8.0,
    "dtcwt_b200_fwd3d_levelq_f32": 9.0, "dtcwt_b200_inv3d_levelq_f32": 9.0,
}


def _cpu_worker_3d(job):
    """One 256^3 volume forward+inverse (3 levels, discard_level_1) with the reference on one core."""
    side, seed, dump, names = job
    for v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[v] = "1"
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refshim
    if refshim.available():
        import logging
        logging.disable(logging.WARNING)
        xf, kind = refshim.load().numpy.Transform3d(*names), "reference"       # dtcwt/numpy/transform3d.py:22,37,133
    else:
        import dtcwt_oracle as O
        from dtcwt_b200 import coeffs
        xf, kind = O.Transform3d(coeffs.biort(names[0]), coeffs.qshift(names[1])), "port"
    X = np.random.RandomState(seed).random_sample((side, side, side)).astype(np.float32)
    t0 = time.perf_counter()
    p = xf.forward(X, 3, discard_level_1=True)
    Z = xf.inverse(p)
    dt = time.perf_counter() - t0
    if dump:
        np.save(os.path.join(dump, "Yl.npy"), np.asarray(p.lowpass, dtype=np.float32))
        for l in (1, 2):
            np.save(os.path.join(dump, "Yh%d.npy" % l), np.asarray(p.highpasses[l], dtype=np.complex64))
        # the reference's discard_level_1 inverse returns axes 0 and 2 exchanged (transform3d.py:452-454)
        np.save(os.path.join(dump, "Z.npy"), np.asarray(Z, dtype=np.float32).transpose(2, 1, 0) if kind == "reference"
                else np.asarray(Z, dtype=np.float32))
    return dt, kind


def run_reference_3d(args):
    """`--impl reference --workload 3d`: the unmodified reference Transform3d on 256^3 volumes, one single-threaded
    worker per core, a step = one volume per worker."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import multiprocessing as mp
    steps = args.steps if args.steps is not None else 2
    warm = args.warmup if args.warmup is not None else 1
    cores = min(host_cores(), 64)
    side = args.side if args.side != SIDE else SIDE_3D
    names = (BIORT, QSHIFT)
    ctx = mp.get_context("spawn")
    jobs = [(side, 2000 + i, None, names) for i in range(cores)]
    with ctx.Pool(cores) as pool:
        kind = pool.map(_cpu_worker_3d, [(32, 1, None, names)] * cores)[0][1]
        for _ in range(warm):
            pool.map(_cpu_worker_3d, jobs)
        t0 = time.perf_counter()
        for _ in range(steps):
            pool.map(_cpu_worker_3d, jobs)
        wall = time.perf_counter() - t0
    value = steps * cores * side ** 3 / wall / 1e6
    sample = "%s: %d worker processes x one %d^3 fp32 volume per step, forward+inverse, 3 levels, discard_level_1" % (
        "unmodified reference dtcwt.numpy.Transform3d (oracle/_ref)" if kind == "reference" else "numpy oracle port", cores, side)
    emit({"impl": "reference", "metric": METRIC_3D, "value": round(value, 3), "unit": "Mvoxels/s", "n_gpus": args.gpus,
          "steps": steps, "warmup": warm, "ms_per_step": round(1e3 * wall / steps, 3), "higher_is_better": True,
          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_3d(args, 8, side),
          "cpu_baseline": {"value": round(value, 3), "unit": "Mvoxels/s", "cores": cores, "kind": kind, "sample": sample},
          "e2e": {"value": round(value, 3), "unit": "Mvoxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
          "gpu_launches": 0})


def config_3d(args, nvol, side):
    return {"workload": "3-D DT-CWT forward+inverse, 3 levels, discard_level_1, %s+%s, chunk of %d x %d^3 fp32 volumes per GPU "
                        "per step (BASELINE configs[3])" % (BIORT, QSHIFT, nvol, side),
            "volumes_per_gpu_per_step": nvol, "side": side, "nlevels": 3, "discard_level_1": True, "ext_mode": 4,
            "cache": "inputs larger than L2: each step reads a %.0f MiB chunk, 3 chunks cycled" % (nvol * side ** 3 * 4 / 2 ** 20)}


def run_3d(args):
    """BASELINE configs[3]: 3-D forward+inverse on chunks of 256^3 fp32 volumes, 3 levels, discard_level_1,
    near_sym_b+qshift_b: Mvoxels/s device-resident (`value`), end to end from pinned host memory (`e2e`), the dominant
    fused entry point against the HBM roofline, the reference on the host cores beside it, full-array parity."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import dtcwt_b200
    from dtcwt_b200 import _lib, coeffs, parallel
    steps = args.steps if args.steps is not None else 10
    warm = max(3, args.warmup if args.warmup is not None else 5)
    rank, world, local = parallel.init("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    nvol = args.images if args.images != 16 else 8
    side = args.side if args.side != SIDE else SIDE_3D
    biort = parallel.broadcast_taps(coeffs.biort(BIORT), 0, dev)
    qshift = parallel.broadcast_taps(coeffs.qshift(QSHIFT), 0, dev)
    xf = dtcwt_b200.Transform3d(biort, qshift)
    g = torch.Generator(device=dev)
    g.manual_seed(4321 + rank)
    pool = [torch.rand((nvol, side, side, side), dtype=torch.float32, device=dev, generator=g) for _ in range(3)]
    parity_seed = 555
    pool[0][0].copy_(torch.from_numpy(np.random.RandomState(parity_seed).random_sample((side, side, side)).astype(np.float32)))

    def step(i):
        p = xf.forward_channels(pool[i % 3], nlevels=3, discard_level_1=True)
        return p, xf.inverse(p)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    parity, cpu_baseline = None, None
    if rank == 0:
        p, Z = step(0)
        torch.cuda.synchronize()
        if world == 1 and not args.no_cpu_baseline:
            import multiprocessing as mp
            import shutil
            import tempfile
            got = {"Yl": p.lowpass_t[0].cpu().numpy(), "Z": Z[0].cpu().numpy(),
                   "Yh1": p.highpasses_t[1][0].cpu().numpy(), "Yh2": p.highpasses_t[2][0].cpu().numpy()}
            del p, Z
            cores = min(host_cores(), 64)
            dump = tempfile.mkdtemp(prefix="dtcwt_bench3d_")
            names = (BIORT, QSHIFT)
            try:
                jobs = [(side, 2000 + i, None, names) for i in range(cores)]
                jobs[0] = (side, parity_seed, dump, names)
                with mp.get_context("spawn").Pool(cores) as wp:
                    wp.map(_cpu_worker_3d, [(32, 1, None, names)] * cores)
                    t0 = time.perf_counter()
                    res = wp.map(_cpu_worker_3d, jobs)
                    wall = time.perf_counter() - t0
                kind = res[0][1]
                per = {}
                for k in sorted(got):
                    ref = np.load(os.path.join(dump, k + ".npy"))
                    assert ref.shape == got[k].shape, (k, ref.shape, got[k].shape)
                    per[k] = float(np.abs(got[k] - ref).max() / np.abs(ref).max())
            finally:
                shutil.rmtree(dump, ignore_errors=True)
            what = "unmodified reference dtcwt.numpy.Transform3d (oracle/_ref)" if kind == "reference" else "numpy oracle port"
            cpu_baseline = {"value": round(cores * side ** 3 / wall / 1e6, 3), "unit": "Mvoxels/s", "cores": cores, "kind": kind,
                            "sample": "%d worker processes x one %d^3 fp32 volume each, forward+inverse, %s (%.1f s wall)" % (
                                cores, side, what, wall)}
            worst = max(per.values())
            parity = {"vs_%s_max_rel_err" % kind: worst, "per_array_rel_err": {k: float("%.3g" % v) for k, v in per.items()},
                      "checked": "volume 0 of chunk 0, all arrays in full (Yl, Yh[1], Yh[2], inverse) vs %s; the reference's "
                                 "discard_level_1 inverse is transposed back (its axes 0/2 come out exchanged)" % what,
                      "tolerance": 1e-5, "ok": bool(worst < 2e-5)}
        else:
            del p, Z

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.prepare()
    log = LaunchLog(torch)
    _lib.set_launch_hook(log)
    log.timing = True
    for i in range(warm):
        step(i)
    barrier()
    per_warm = log.per_kernel_ms()
    breakdown = {k: round(v[0] / warm, 4) for k, v in sorted(per_warm.items())}
    log.only = max(per_warm.items(), key=lambda kv: kv[1][0])[0] if per_warm else None
    log.events = []
    if rank == 0:
        sampler.start()
    log.count = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(steps):
        step(warm + i)
    e1.record()
    barrier()
    log.timing = False
    launches = log.count
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None
    vox_rank = nvol * side ** 3
    value = world * vox_rank * steps / (ms_total / 1e3) / 1e6
    peak, peak_src = measured_peak_gbs()
    roofline = None
    per = log.per_kernel_ms()
    if per:
        sym, (tot_ms, n) = max(per.items(), key=lambda kv: kv[1][0])
        per_step = n / steps
        algo = KERNEL_ALGO_BYTES_3D.get(sym, ALGO_BYTES_PER_VOXEL) * vox_rank      # per step
        ach = algo / (tot_ms / steps / 1e3) / 1e9
        roofline = {"bound": "hbm", "kernel": sym, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(ach / peak, 4), "traffic": None, "peak_source": peak_src,
                    "algorithmic_gb_per_step": round(algo / 1e9, 4), "launches_per_step": per_step,
                    "note": "%s: %.1f B/voxel algorithmic per step over its %d launch(es) per step (two kernels each: slices + depth)" % (
                        sym, KERNEL_ALGO_BYTES_3D.get(sym, ALGO_BYTES_PER_VOXEL), per_step),
                    "ms_per_step_of_kernel": round(tot_ms / steps, 4), "share_of_step": round(tot_ms / ms_total, 4),
                    "whole_step_frac": round(ALGO_BYTES_PER_VOXEL * vox_rank * steps / (ms_total / 1e3) / 1e9 / peak, 4),
                    "kernels_ms_per_step": breakdown}
    _lib.set_launch_hook(None)

    e2e = None
    if not args.no_e2e:
        def fn(x):
            return xf.inverse(xf.forward_channels(x, nlevels=3, discard_level_1=True))
        e2e = run_e2e_generic(torch, dist, fn, pool, steps, world, dev, barrier, (nvol, side, side, side), "Mvoxels/s")
    if rank == 0:
        emit({"metric": METRIC_3D, "value": round(value, 2), "unit": "Mvoxels/s", "n_gpus": world, "steps": steps, "warmup": warm,
              "ms_per_step": round(ms_total / steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
              "dtype": "f32", "data": "synthetic", "config": config_3d(args, nvol, side), "clocks": clocks, "e2e": e2e,
              "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline, "parity": parity})
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_e2e_generic(torch, dist, fn, pool, steps, world, dev, barrier, shape, unit, nbuf=3):
    """End to end through the public API from pinned HOST memory: every step uploads its chunk (H2D), runs
    res = fn(dev_in) and reads the result back (D2H) straight from the tensor the API returned; `nbuf` chunks are in
    flight so the copies of neighbouring steps overlap the kernels on side streams.  Timed with CUDA events, max over
    ranks.  The copy roofline of a step is bytes / the bidirectional pinned-copy bandwidth (tools/exp/copy_bw.py)."""
    numel = 1
    for v in shape:
        numel *= v
    host_in = [torch.empty(shape, dtype=torch.float32).pin_memory() for _ in range(nbuf)]
    host_out = [torch.empty(shape, dtype=torch.float32).pin_memory() for _ in range(nbuf)]
    for b in range(nbuf):
        host_in[b].copy_(pool[b % len(pool)].cpu())
    dev_in = [torch.empty(shape, dtype=torch.float32, device=dev) for _ in range(nbuf)]
    compute = torch.cuda.current_stream()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    ev_in = [torch.cuda.Event() for _ in range(nbuf)]
    ev_done = [torch.cuda.Event() for _ in range(nbuf)]
    ev_out = [torch.cuda.Event() for _ in range(nbuf)]
    keep = [None] * nbuf

    def run(n):
        for i in range(n):
            b = i % nbuf
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_done[b])            # the previous user of dev_in[b] has finished
                dev_in[b].copy_(host_in[b], non_blocking=True)
                ev_in[b].record(s_in)
            compute.wait_event(ev_in[b])
            res = fn(dev_in[b])
            ev_done[b].record(compute)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_out[b])            # host_out[b] has been drained by its previous copy (stream order)
                s_out.wait_event(ev_done[b])
                host_out[b].copy_(res, non_blocking=True)
                res.record_stream(s_out)
                ev_out[b].record(s_out)
            keep[b] = res
        s_out.synchronize()

    run(nbuf)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(steps)
    torch.cuda.synchronize()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms.item()) / steps
    val = world * numel * steps / (float(ms.item()) / 1e3) / 1e6
    out = {"value": round(val, 2), "unit": unit, "h2d_bytes_per_step": numel * 4, "d2h_bytes_per_step": numel * 4,
           "note": "pinned host chunk -> forward -> inverse -> pinned host through the public API; %d chunks in flight, copies on "
                   "side streams, D2H straight from the returned tensor" % nbuf,
           "ms_per_step": round(ms_step, 3)}
    bw = copy_roofline_gbs(world)
    if bw is not None:
        ideal_ms = numel * 4 / (bw * 1e9) * 1e3
        out["copy_roofline"] = {"bidir_gbs_per_gpu": bw, "ideal_ms_per_step": round(ideal_ms, 3), "frac": round(ideal_ms / ms_step, 4),
                                "source": "profiles/r2_copy_bw.json (tools/exp/copy_bw.py on this pool, %d GPU(s) copying at once)" % world}
    return out


def copy_roofline_gbs(world):
    """Measured pinned-memory copy bandwidth per GPU and direction with H2D and D2H running together on `world` GPUs of
    one box (committed measurement, profiles/r2_copy_bw.json); None when there is no entry for this GPU count."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_copy_bw.json")) as f:
            return float(json.load(f)[str(world)]["bidir_each_gbs_per_gpu_min"])
    except Exception:
        return None


# =============================================================================== config 5: registration of frame pairs
METRIC_REG = "frame pairs/s dtcwt.registration.estimatereg 1920x1080 5-level"


def _reg_pair(shape, seed):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import reg_frames
    f1, f2 = reg_frames(shape, seed)
    return f1.astype("float32"), f2.astype("float32")


def _cpu_worker_reg(job):
    """Two 5-level forward transforms + estimatereg of one frame pair with the reference on one core."""
    shape, seed, dump = job
    for v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[v] = "1"
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import logging
    import refshim
    logging.disable(logging.WARNING)
    d, reg = refshim.load(), refshim.load_registration()
    f1, f2 = _reg_pair(shape, seed)
    xf = d.numpy.Transform2d()                                     # library defaults, as examples/register_images.py
    t0 = time.perf_counter()
    avecs = reg.estimatereg(xf.forward(f1, nlevels=5), xf.forward(f2, nlevels=5))      # dtcwt/registration.py:304
    dt = time.perf_counter() - t0
    if dump:
        np.save(os.path.join(dump, "avecs.npy"), avecs)
    return dt


def run_reg(args):
    """BASELINE configs[4]: estimatereg on batched 1080 x 1920 fp32 frame pairs, 5 levels, default wavelets.  A step =
    forward transform of both frames of every pair of the chunk + estimatereg, all on the device."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import dtcwt_b200
    from dtcwt_b200 import _lib, parallel, registration as R
    steps = args.steps if args.steps is not None else 10
    warm = max(3, args.warmup if args.warmup is not None else 3)
    rank, world, local = parallel.init("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    npair = args.images if args.images != 16 else 32
    shape = (1080, 1920)
    xf = dtcwt_b200.Transform2d()
    base1, base2 = _reg_pair(shape, 99)
    g = torch.Generator(device=dev)
    g.manual_seed(99 + rank)
    # every pair of a chunk is the seeded pair plus its own low-amplitude noise: distinct data, same flow
    pool = []
    for c in range(3):
        n1 = 0.01 * torch.rand((npair,) + shape, dtype=torch.float32, device=dev, generator=g)
        pool.append((torch.from_numpy(base1).to(dev) + n1, torch.from_numpy(base2).to(dev) + n1))
    pool[0][0][0].copy_(torch.from_numpy(base1))
    pool[0][1][0].copy_(torch.from_numpy(base2))

    def step(i):
        a, b = pool[i % 3]
        return R.estimatereg(xf.forward_channels(a, "nhw", nlevels=5), xf.forward_channels(b, "nhw", nlevels=5))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    parity, cpu_baseline = None, None
    if rank == 0:
        got = step(0)[0].cpu().numpy()
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import refshim
        if world == 1 and not args.no_cpu_baseline and refshim.available():
            import multiprocessing as mp
            import shutil
            import tempfile
            cores = min(host_cores(), 64)
            dump = tempfile.mkdtemp(prefix="dtcwt_benchreg_")
            try:
                jobs = [(shape, 99, dump if i == 0 else None) for i in range(cores)]
                with mp.get_context("spawn").Pool(cores) as wp:
                    wp.map(_cpu_worker_reg, [((135, 240), 1, None)] * cores)
                    t0 = time.perf_counter()
                    for _ in range(3):
                        wp.map(_cpu_worker_reg, jobs)
                    wall = time.perf_counter() - t0
                want = np.load(os.path.join(dump, "avecs.npy"))
            finally:
                shutil.rmtree(dump, ignore_errors=True)
            err = np.abs(got - want)
            cpu_baseline = {"value": round(3 * cores / wall, 3), "unit": "frame pairs/s", "cores": cores, "kind": "reference",
                            "sample": "%d worker processes x 3 frame pairs each: two 5-level forward transforms + estimatereg with the "
                                      "unmodified reference (oracle/_ref) (%.1f s wall)" % (cores, wall)}
            parity = {"avecs_max_abs_err": float(err.max()), "avecs_median_abs_err": float(np.median(err)),
                      "avecs_max_abs": float(np.abs(want).max()), "checked": "affine-parameter grid of pair 0 of chunk 0 vs the reference's "
                      "estimatereg on the reference's own float32 pyramids", "tolerance": 1e-4, "ok": bool(err.max() < 1e-4)}
    log = LaunchLog(torch)
    _lib.set_launch_hook(log)
    for i in range(warm):
        step(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.prepare()
        sampler.start()
    log.count = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(steps):
        step(warm + i)
    e1.record()
    barrier()
    launches = log.count
    _lib.set_launch_hook(None)
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None
    value = world * npair * steps / (ms_total / 1e3)
    peak, peak_src = measured_peak_gbs()
    # compulsory traffic of the two forward transforms (read 4 + write 16 B/pixel each); estimatereg works on levels 3-5 only
    algo = 2 * 20.0 * npair * shape[0] * shape[1]
    if rank == 0:
        emit({"metric": METRIC_REG, "value": round(value, 2), "unit": "frame pairs/s", "n_gpus": world, "steps": steps, "warmup": warm,
              "ms_per_step": round(ms_total / steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
              "dtype": "f32 transforms, f64 registration arithmetic", "data": "synthetic",
              "config": {"workload": "estimatereg on chunks of %d frame pairs of 1080x1920 fp32, 5 levels, near_sym_a+qshift_a "
                                     "(BASELINE configs[4]); both forward transforms are inside the step" % npair,
                         "pairs_per_gpu_per_step": npair, "height": shape[0], "width": shape[1], "nlevels": 5},
              "clocks": clocks, "gpu_launches": launches,
              "roofline": {"bound": "hbm", "achieved": round(algo * steps / (ms_total / 1e3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                           "frac": round(algo * steps / (ms_total / 1e3) / 1e9 / peak, 4), "traffic": None, "peak_source": peak_src,
                           "note": "whole step against the compulsory traffic of its two 5-level forward transforms (40 B per frame-pair pixel)"},
              "cpu_baseline": cpu_baseline, "parity": parity})
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    global BIORT, QSHIFT
    args = parse()
    BIORT, QSHIFT = args.biort, args.qshift
    if args.impl == "reference":
        if args.workload == "3d":
            run_reference_3d(args)
        else:
            run_reference(args)
        return
    if not (args.gpus > 1 and "WORLD_SIZE" not in os.environ):
        claim_stdout()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: re-launch under torchrun the way the driver does
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.workload == "3d":
        run_3d(args)
        return
    if args.workload == "reg":
        run_reg(args)
        return
    run_ours(args)


if __name__ == "__main__":
    main()
