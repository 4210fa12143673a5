"""CPU tier — the N>1 path with world_size-2 gloo process groups: the statistics protocol (the engine's only
collective), contiguous sharding, and shard-invariance of the seeded sample stream."""
import os
import socket

import numpy as np
import pytest

from lq_mpc_b200 import sampling as sp
from lq_mpc_b200 import stats as st


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, table, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = st.shard_bounds(table.shape[1], rank, world)
        out = st.distributed_stats_protocol(table[:, lo:hi])
        q.put((rank, {k: np.asarray(v) for k, v in out.items()}))
    finally:
        dist.destroy_process_group()


def _run(world, table):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, table, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_stats_equal_single_process_numpy(world):
    rng = np.random.default_rng(5)
    table = rng.normal(size=(7, 1001)) * np.logspace(-3, 3, 7)[:, None] + 10.0
    table[2, 17] = np.inf                      # unstable sample: excluded from the moments, counted separately
    table[4, 900] = np.nan
    res = _run(world, table)
    fin = np.isfinite(table)
    for r in range(world):
        o = res[r]
        for c in range(7):
            col = table[c][fin[c]]
            assert o["max"][c] == col.max() and o["min"][c] == col.min()          # exact
            assert abs(o["mean"][c] - col.mean()) <= 1e-12 * abs(col.mean())
            assert abs(o["std"][c] - col.std()) <= 1e-12 * col.std()
            assert o["count"][c] == col.size and o["n_nonfinite"][c] == (~fin[c]).sum()
    for k in res[0]:                           # identical on every rank
        for r in range(1, world):
            assert np.array_equal(res[0][k], res[r][k], equal_nan=True)


def test_merge_moments_matches_numpy_and_kernel_formula():
    """Chan merge over ragged shards (incl. an empty one and an all-non-finite one) == numpy on the whole column;
    `local_moments_numpy` restates the kernel's shifted-data formula."""
    rng = np.random.default_rng(2)
    table = 5.0 + rng.normal(size=(4, 5000)) * np.array([1e-6, 1.0, 1e3, 1e-2])[:, None]
    table[1, 100:160] = np.inf
    cuts = [0, 0, 60, 100, 160, 1700, 5000]                        # shard [100,160) of column 1 is all-inf
    parts = [st.local_moments_numpy(table[:, a:b]) for a, b in zip(cuts, cuts[1:])]
    out = st.merge_moments(np.stack(parts))
    for c in range(4):
        col = table[c][np.isfinite(table[c])]
        assert out["max"][c] == col.max() and out["min"][c] == col.min() and out["count"][c] == col.size
        assert abs(out["mean"][c] - col.mean()) <= 1e-14 * abs(col.mean())
        assert abs(out["std"][c] - col.std()) <= 1e-9 * col.std()
    assert out["n_nonfinite"][1] == 60
    none = st.merge_moments(np.stack([st.local_moments_numpy(np.zeros((2, 0)))]))
    assert np.all(none["count"] == 0) and np.all(np.isnan(none["mean"])) and np.all(np.isnan(none["std"]))


def test_shard_bounds_partition():
    for S in (0, 1, 7, 100, 12_500_001):
        for world in (1, 2, 3, 8):
            edges = [st.shard_bounds(S, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == S
            for (a, b), (c, d) in zip(edges, edges[1:]):
                assert b == c and 0 <= (b - a) - (d - c) <= 1


def test_sample_stream_is_shard_invariant():
    """Rank r of N asks for global samples [first, first+S): it must see exactly the single-GPU stream's slice."""
    n, m, S = 4, 2, 200_000
    full = sp.synth_samples_soa(n, m, S, seed=1, first=0)
    for world in (2, 8):
        for rank in (0, world - 1):
            lo, hi = st.shard_bounds(S, rank, world)
            part = sp.synth_samples_soa(n, m, hi - lo, seed=1, first=lo)
            for a, b in zip(full, part):
                assert np.array_equal(a[:, lo:hi], b)
    odd = sp.synth_samples_soa(n, m, 70_001, seed=1, first=65_530)        # straddles a Philox block boundary
    for a, b in zip(full, odd):
        assert np.array_equal(a[:, 65_530:65_530 + 70_001], b)


def test_engine_sampler_matches_oracle_sampler():
    """bench.py's GPU arm (lq_mpc_b200.sampling) and CPU arm (oracle.np_batched) must draw the same workload."""
    from oracle import np_batched as nb
    A1, B1, Q1, R1 = sp.synth_problem(4, 2, seed=0)
    A2, B2, Q2, R2 = nb.synth_problem(4, 2, seed=0)
    assert np.array_equal(A1, A2) and np.array_equal(B1, B2)
    dA, dB, x0 = sp.synth_samples_soa(4, 2, 5000, seed=1, first=123)
    rA, rB, rx = nb.to_soa(*nb.synth_samples(4, 2, 5000, seed=1, first=123))
    assert np.array_equal(dA, rA) and np.array_equal(dB, rB) and np.array_equal(x0, rx)
