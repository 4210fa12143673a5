"""Column statistics of result tables, single- or multi-GPU.

The reference reduces every table over its sample axis with np.max / np.min / np.mean / np.std (utils.py:895-898).
Here each rank reduces its shard on the GPU in ONE pass (K5, csrc/k_stats.cu: max, min, counts, mean, M2 by
shifted-data accumulation with fixed-order folds) and the only data-path collective of the whole engine follows: one
all-gather of 6 doubles per column; every rank then merges the shards in rank order with Chan's pairwise update, so
all ranks hold bit-identical statistics and the result matches np.mean / np.std to rounding.

`merge_moments` is the pure host-side combination rule, shared by the NCCL path and the world_size-2 gloo tests.
`column_stats_two_pass` keeps np.std's literal two-pass scheme (three all-reduces) as a cross-check.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np


def finish(mx, mn, n_fin, n_bad, mean, m2):
    """Per-column dict in the reference's vocabulary from merged moments (numpy arrays)."""
    with np.errstate(invalid="ignore", divide="ignore"):
        std = np.sqrt(m2 / n_fin)
    return {"max": mx, "min": mn, "mean": mean, "std": std, "count": n_fin, "n_nonfinite": n_bad}


def merge_moments(parts):
    """Chan's pairwise update over shards, in shard (= rank) order. parts: [n_shards][cols][6] array with rows
    {max, min, n_finite, n_nonfinite, mean, M2}; empty shards (n_finite = 0) carry NaN mean/M2 and are skipped."""
    parts = np.asarray(parts, dtype=np.float64)
    cols = parts.shape[1]
    mx = np.full(cols, -np.inf); mn = np.full(cols, np.inf)
    n = np.zeros(cols); nb = np.zeros(cols); mean = np.zeros(cols); m2 = np.zeros(cols)
    for p in parts:
        bn = p[:, 2]
        has = bn > 0
        bmean = np.where(has, p[:, 4], 0.0)
        bm2 = np.where(has, p[:, 5], 0.0)
        tot = n + bn
        with np.errstate(invalid="ignore", divide="ignore"):
            w = np.where(tot > 0, bn / np.where(tot > 0, tot, 1.0), 0.0)
        delta = bmean - mean
        m2 = m2 + bm2 + delta * delta * n * w
        mean = mean + delta * w
        n = tot
        nb = nb + p[:, 3]
        mx = np.maximum(mx, p[:, 0]); mn = np.minimum(mn, p[:, 1])
    mean = np.where(n > 0, mean, np.nan)
    m2 = np.where(n > 0, m2, np.nan)
    return finish(mx, mn, n, nb, mean, m2)


def column_moments_device(engine, table, group=None):
    """Asynchronous half of `column_stats`: K5 on this rank's shard plus the all-gather, everything enqueued on the
    device; returns the [world][cols][6] DEVICE tensor of per-shard moments (no host synchronisation). Finish with
    `merge_moments(t.cpu().numpy())` whenever the numbers are needed on the host."""
    import torch
    import torch.distributed as dist
    mom = engine.column_moments_raw(table)                      # device [cols][6]
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        allm = torch.empty((world,) + tuple(mom.shape), dtype=mom.dtype, device=mom.device)
        dist.all_gather_into_tensor(allm, mom, group=group)
        return allm
    return mom[None]


def column_stats(engine, table, group=None):
    """table: [cols][S_local] device tensor. One pass over the shard (K5 `lqmpc_column_moments`), ONE collective
    (all-gather of 6 doubles per column), Chan merge in rank order. Returns dict of numpy arrays (max, min, mean, std,
    count, n_nonfinite), identical on every rank."""
    return merge_moments(column_moments_device(engine, table, group).cpu().numpy())


def column_stats_two_pass(engine, table, group=None):
    """np.std's own two-pass scheme (max/min/sum/counts, all-reduce, squared deviations about the global mean,
    all-reduce): three collectives, two passes. Kept as the cross-check of the one-pass path."""
    import torch
    import torch.distributed as dist

    def _allreduce(t, op):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(t, op=op, group=group)
        return t
    raw = engine.column_stats_raw(table)                      # [cols][5]
    ext = torch.stack([raw[:, 0], -raw[:, 1]], dim=0).contiguous()
    add = torch.stack([raw[:, 2], raw[:, 3], raw[:, 4]], dim=0).contiguous()
    _allreduce(ext, dist.ReduceOp.MAX)
    _allreduce(add, dist.ReduceOp.SUM)
    mean = add[0] / add[1]
    sq = engine.column_sqdev_raw(table, mean)
    _allreduce(sq, dist.ReduceOp.SUM)
    return finish(ext[0].cpu().numpy(), -ext[1].cpu().numpy(), add[1].cpu().numpy(), add[2].cpu().numpy(),
                  mean.cpu().numpy(), sq.cpu().numpy())


# ---------------------------------------------------------------------------------------------------------------
# host-side protocol mirror (used by the gloo CPU tests: same collective and merge, partials supplied by the caller)
def local_moments_numpy(table):
    """What K5 (`moments_partial/final_kernel`) produces for one shard, restated with numpy for the CPU protocol tests
    only: [cols][6] = max, min, n_finite, n_nonfinite, mean, M2 with the same shifted-data formula."""
    t = np.asarray(table, dtype=np.float64)
    cols = t.shape[0]
    out = np.zeros((cols, 6))
    for c in range(cols):
        col = t[c]
        fin = np.isfinite(col)
        v = col[fin]
        k = col[0] if col.size and np.isfinite(col[0]) else 0.0
        n = float(v.size)
        d = v - k
        s1, s2 = d.sum(), (d * d).sum()
        out[c] = [v.max() if v.size else -np.inf, v.min() if v.size else np.inf, n, float((~fin).sum()),
                  k + s1 / n if n else np.nan, max(s2 - s1 * (s1 / n), 0.0) if n else np.nan]
    return out


def distributed_stats_protocol(local_table, group=None, partials=local_moments_numpy):
    """The one-collective protocol on CPU tensors (gloo). `local_table`: [cols][S_local] numpy array."""
    import torch
    import torch.distributed as dist
    mom = torch.from_numpy(np.ascontiguousarray(partials(local_table)))
    world = dist.get_world_size(group)
    gathered = [torch.empty_like(mom) for _ in range(world)]
    dist.all_gather(gathered, mom, group=group)
    return merge_moments(np.stack([g.numpy() for g in gathered]))


def shard_bounds(S: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of the sample axis for `rank`: sizes differ by at most one."""
    base, rem = divmod(S, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
