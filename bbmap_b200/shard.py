"""Sharding of read / task batches across the GPUs of one box.

The path needs no collective: the reference and (later) the index are replicated on every GPU and each rank gets a
contiguous slice of the batch in input order, so concatenating the per-rank results by rank reproduces the single-GPU
output order (the reference's `ordered=t` semantics, one list of reads per mapping thread:
current/align2/AbstractMapThread.java:390,574).  torch.distributed is used only for the barrier, the max-over-ranks of
device timings and, optionally, gathering results on rank 0.
"""
import numpy as np


def shard_bounds(n, rank, world):
    """Contiguous [lo, hi) of rank `rank`; sizes differ by at most one; concatenation over ranks = range(n)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_msa_tasks(reads, tasks, rank, world):
    """Slice of an MSA task batch with its own compact reads buffer (read_off rebased)."""
    lo, hi = shard_bounds(len(tasks), rank, world)
    t = tasks[lo:hi].copy()
    if len(t) == 0:
        return reads[:0].copy(), t
    # tasks produced by the generators are in buffer order; handle the general case anyway
    starts = t["read_off"].astype(np.int64)
    lens = t["read_len"].astype(np.int64)
    new_off = np.zeros(len(t), np.int64)
    np.cumsum(lens[:-1], out=new_off[1:])
    out = np.empty(int(lens.sum()), reads.dtype)
    if (np.diff(starts) == lens[:-1]).all():
        out[:] = reads[starts[0]: starts[0] + len(out)]
    else:
        for i in range(len(t)):
            out[new_off[i]: new_off[i] + lens[i]] = reads[starts[i]: starts[i] + lens[i]]
    t["read_off"] = new_off
    return out, t


def max_over_ranks(values, device=None):
    """MAX-all-reduce of a list of floats (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def sum_over_ranks(values, device=None):
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()]


def gather_records(local, dst=0):
    """Gather per-rank structured arrays on `dst` in rank order (None elsewhere)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size()
    bufs = [None] * world if dist.get_rank() == dst else None
    dist.gather_object(local.tobytes(), bufs, dst=dst)
    if dist.get_rank() != dst:
        return None
    return np.concatenate([np.frombuffer(b, dtype=local.dtype) for b in bufs])
