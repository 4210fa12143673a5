"""Column statistics of result tables, single- or multi-GPU.

The reference reduces every table over its sample axis with np.max / np.min / np.mean / np.std (utils.py:895-898).
Here each rank reduces its shard on the GPU (K5, csrc/k_stats.cu) and the only data-path collective of the whole
engine follows: tiny all-reduces of the per-column partials (MAX for {max, -min}; SUM for {sum, counts}; SUM for the
squared deviations about the global mean — np.std's own two-pass scheme, so the result matches numpy to rounding).

`merge_partials` is the pure host-side combination rule, shared by the NCCL path and the world_size-2 gloo tests.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np


def finish(mx, mn, total, n_fin, n_bad, sqdev):
    """Per-column dict in the reference's vocabulary from globally reduced partials (numpy arrays)."""
    with np.errstate(invalid="ignore", divide="ignore"):
        mean = total / n_fin
        std = np.sqrt(sqdev / n_fin)
    return {"max": mx, "min": mn, "mean": mean, "std": std, "count": n_fin, "n_nonfinite": n_bad}


def _allreduce(t, op, group):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=op, group=group)
    return t


def column_stats(engine, table, group=None):
    """table: [cols][S_local] device tensor. Returns dict of numpy arrays (max, min, mean, std, count,
    n_nonfinite), identical on every rank."""
    import torch
    import torch.distributed as dist
    raw = engine.column_stats_raw(table)                      # [cols][5]
    ext = torch.stack([raw[:, 0], -raw[:, 1]], dim=0).contiguous()
    add = torch.stack([raw[:, 2], raw[:, 3], raw[:, 4]], dim=0).contiguous()
    _allreduce(ext, dist.ReduceOp.MAX if dist.is_available() else None, group)
    _allreduce(add, dist.ReduceOp.SUM if dist.is_available() else None, group)
    mean = add[0] / add[1]
    sq = engine.column_sqdev_raw(table, mean)
    _allreduce(sq, dist.ReduceOp.SUM if dist.is_available() else None, group)
    return finish(ext[0].cpu().numpy(), -ext[1].cpu().numpy(), add[0].cpu().numpy(), add[1].cpu().numpy(),
                  add[2].cpu().numpy(), sq.cpu().numpy())


# ---------------------------------------------------------------------------------------------------------------
# host-side protocol mirror (used by the gloo CPU tests: same collectives, partials supplied by the caller)
def local_partials_numpy(table):
    """What K5 pass 1 produces for one shard, restated with numpy for the CPU protocol tests only."""
    t = np.asarray(table, dtype=np.float64)
    fin = np.isfinite(t)
    tz = np.where(fin, t, 0.0)
    mx = np.where(fin, t, -np.inf).max(axis=1) if t.shape[1] else np.full(t.shape[0], -np.inf)
    mn = np.where(fin, t, np.inf).min(axis=1) if t.shape[1] else np.full(t.shape[0], np.inf)
    return mx, mn, tz.sum(axis=1), fin.sum(axis=1).astype(np.float64), (~fin).sum(axis=1).astype(np.float64)


def distributed_stats_protocol(local_table, group=None, partials=local_partials_numpy):
    """The three-collective protocol on CPU tensors (gloo). `local_table`: [cols][S_local] numpy array."""
    import torch
    import torch.distributed as dist
    mx, mn, sm, nf, nb = partials(local_table)
    ext = torch.from_numpy(np.stack([mx, -mn]))
    add = torch.from_numpy(np.stack([sm, nf, nb]))
    dist.all_reduce(ext, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(add, op=dist.ReduceOp.SUM, group=group)
    mean = (add[0] / add[1]).numpy()
    t = np.asarray(local_table, dtype=np.float64)
    fin = np.isfinite(t)
    sq = torch.from_numpy(np.where(fin, (t - mean[:, None]) ** 2, 0.0).sum(axis=1))
    dist.all_reduce(sq, op=dist.ReduceOp.SUM, group=group)
    return finish(ext[0].numpy(), -ext[1].numpy(), add[0].numpy(), add[1].numpy(), add[2].numpy(), sq.numpy())


def shard_bounds(S: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of the sample axis for `rank`: sizes differ by at most one."""
    base, rem = divmod(S, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
