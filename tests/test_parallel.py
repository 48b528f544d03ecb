"""Multi-process host logic of the batch-sharded path (dtcwt_b200/parallel.py) at world size 2 over gloo.

The data path has no collective: each rank transforms its own contiguous slice of the batch.  What is
checked here is (a) the slices tile the batch exactly, (b) the one broadcast of the filter taps delivers
rank 0's values to everybody, (c) every rank's slice, pushed through the host layer (on the kernel-logic
emulator -- there is no GPU here), equals the oracle on the same images.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.normpath(os.path.join(os.path.dirname(__file__), ".."))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, emu_path, q):
    try:
        os.environ.update({"RANK": str(rank), "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank),
                           "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port)})
        for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
            if p not in sys.path:
                sys.path.insert(0, p)
        import torch
        import torch.distributed as dist
        import dtcwt_b200
        import dtcwt_oracle as O
        from dtcwt_b200 import _lib, coeffs, parallel
        import emu_seam
        emu_seam.install(emu_path)
        r, w, _ = parallel.init("gloo")
        assert (r, w) == (rank, world)
        # (b) rank 0 owns the real taps, the others start from garbage of the same lengths
        biort, qshift = coeffs.biort("near_sym_a"), coeffs.qshift("qshift_a")
        if rank != 0:
            biort = tuple(np.full_like(t, 7.0) for t in biort)
            qshift = tuple(np.full_like(t, -3.0) for t in qshift)
        biort = parallel.broadcast_taps(biort, 0, "cpu")
        qshift = parallel.broadcast_taps(qshift, 0, "cpu")
        for got, want in zip(biort + qshift, coeffs.biort("near_sym_a") + coeffs.qshift("qshift_a")):
            assert np.array_equal(got.reshape(-1), np.asarray(want).reshape(-1))
        # (a) + (c) the same seeded batch everywhere, each rank transforms its slice only
        batch = np.random.RandomState(5).rand(5, 48, 64).astype(np.float32)
        lo, hi = parallel.shard_range(batch.shape[0], rank, world)
        mine = parallel.shard(torch.from_numpy(batch), rank, world)
        assert mine.shape[0] == hi - lo
        xf = dtcwt_b200.Transform2d(biort, qshift)
        p = xf.forward_channels(mine, "nhw", nlevels=2)
        z = xf.inverse_channels(p, "nhw")
        to = O.Transform2d(coeffs.biort("near_sym_a"), coeffs.qshift("qshift_a"))
        worst = 0.0
        for i in range(hi - lo):
            po = to.forward(batch[lo + i], 2)
            worst = max(worst, float(np.abs(p.lowpass[i] - po.lowpass).max() / np.abs(po.lowpass).max()))
            for a, b in zip(p.highpasses, po.highpasses):
                worst = max(worst, float(np.abs(a[i] - b).max() / np.abs(b).max()))
            worst = max(worst, float(np.abs(z[i].numpy() - batch[lo + i]).max()))
        counts = torch.tensor([hi - lo], dtype=torch.int64)
        dist.all_reduce(counts)                      # test-only bookkeeping, not part of the data path
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, lo, hi, worst, int(counts.item())))
    except Exception as e:      # surface the failure in the parent
        import traceback
        q.put((rank, "error", traceback.format_exc(), repr(e), 0))


def test_shard_ranges_tile_the_batch():
    sys.path.insert(0, ROOT)
    from dtcwt_b200 import parallel
    for n in (0, 1, 5, 16, 1024):
        for world in (1, 2, 3, 8):
            edges = [parallel.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for (a, b), (c, d) in zip(edges, edges[1:]):
                assert b == c and 0 <= (b - a) - (d - c) <= 1


def test_world_size_2_gloo(emulator_path):
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, emulator_path, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = []
    try:
        for _ in procs:
            res.append(q.get(timeout=240))
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()
    for r in res:
        assert r[1] != "error", r[2]
    res.sort()
    assert [(r[1], r[2]) for r in res] == [(0, 3), (3, 5)]
    assert all(r[4] == 5 for r in res)
    assert max(r[3] for r in res) < 1e-5
