"""Batch sharding across the GPUs of one box: one process per GPU, no data-path collective.

Every image / volume of a batch is transformed independently (reference
``dtcwt/numpy/transform2d.py:40-188`` has no cross-sample term), so a batch is
split into contiguous slices, one per rank, and the outputs stay sharded.  The only
communication is ONE broadcast of the packed filter taps from rank 0 at start-up
(about 1 KB) so that all ranks provably filter with identical coefficients, plus
the barriers a benchmark needs for timing.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

__all__ = ["init", "shard_range", "shard", "broadcast_taps", "pack_taps", "unpack_taps"]


def init(backend=None):
    """Initialise torch.distributed from the torchrun environment; returns (rank, world, local_rank).
    A plain single-process run (no WORLD_SIZE) returns (0, 1, 0) without creating a group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(n_items, rank, world):
    """Contiguous slice [lo, hi) of ``n_items`` owned by ``rank``; sizes differ by at most one."""
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(batch, rank, world):
    """This rank's slice of a batch tensor / array along axis 0 (a view, no copy)."""
    lo, hi = shard_range(batch.shape[0], rank, world)
    return batch[lo:hi]


def pack_taps(taps):
    """tuple of tap vectors -> (flat float64 tensor, lengths) for one broadcast."""
    vecs = [np.asarray(t, dtype=np.float64).reshape(-1) for t in taps]
    lens = [len(v) for v in vecs]
    flat = np.concatenate(vecs) if vecs else np.zeros(0)
    return torch.from_numpy(flat.copy()), lens


def unpack_taps(flat, lens):
    flat = flat.detach().cpu().numpy()
    out, o = [], 0
    for n in lens:
        out.append(flat[o:o + n].reshape(-1, 1).copy())
        o += n
    return tuple(out)


def broadcast_taps(taps, src=0, device=None):
    """Broadcast a tuple of tap vectors from ``src``; every rank returns rank-``src``'s values.
    All ranks must pass tuples with the same lengths (they name the same wavelet family)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tuple(np.asarray(t, dtype=np.float64).reshape(-1, 1) for t in taps)
    flat, lens = pack_taps(taps)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else "cpu"
    flat = flat.to(device)
    dist.broadcast(flat, src=src)
    return unpack_taps(flat, lens)
