"""Worker of tests/test_multigpu.py (one process per GPU under torch.distributed.run, NCCL): SURVEY 4 item 4 —
sharded statistics == single-GPU statistics == numpy, max/min bit-equal.

Every rank evaluates K1 on its contiguous shard of ONE seeded sample set (the Philox stream is indexed by the global
sample number, so the shards together are exactly the single-GPU batch), reduces it with K5 and takes part in the one
all-gather; rank 0 then evaluates the WHOLE set alone, reduces it with K5 and with numpy, and compares."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from lq_mpc_b200 import sampling as sp, stats
    from lq_mpc_b200.engine import Engine
    eng = Engine(local)
    S = int(os.environ.get("LQMPC_TEST_S", 1_000_003))              # ragged: shards differ by one sample
    A, B, Q, R = sp.synth_problem(4, 2, seed=0)
    eng.set_problem(A, B, Q, R, Q, None, None, 30)
    lo, hi = stats.shard_bounds(S, rank, world)
    dA, dB, x0 = sp.synth_samples_soa(4, 2, hi - lo, seed=3, first=lo, e=0.05)
    r = eng.eval_batch(dA, dB, x0, 9, 10, want=("J", "rho", "ratio"))
    merged = stats.column_stats(eng, r["table"])                    # K5 + the one all-gather + Chan merge
    two = stats.column_stats_two_pass(eng, r["table"])              # literal np.std scheme, three all-reduces
    # every rank must hold bit-identical merged statistics
    mine = torch.tensor(np.stack([merged[k] for k in ("max", "min", "mean", "std", "count")]), device="cuda")
    allm = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allm, mine)
    for other in allm:
        assert torch.equal(other, allm[0]), "ranks disagree on the merged statistics"
    ok = True
    if rank == 0:
        fA, fB, fx = sp.synth_samples_soa(4, 2, S, seed=3, first=0, e=0.05)
        full = eng.eval_batch(fA, fB, fx, 9, 10, want=("J", "rho", "ratio"))
        single_raw = eng.column_moments_raw(full["table"]).cpu().numpy()
        single = stats.merge_moments(single_raw[None])
        tab = full["table"].cpu().numpy()
        fin = np.where(np.isfinite(tab), tab, np.nan)
        for k, f in (("max", np.nanmax), ("min", np.nanmin)):
            assert np.array_equal(merged[k], single[k]), k            # bit-equal: sharded == single GPU
            assert np.array_equal(merged[k], f(fin, axis=1)), k       # == numpy
            assert np.array_equal(two[k], single[k]), k
        assert np.array_equal(merged["count"], single["count"])
        assert np.array_equal(merged["count"], np.isfinite(tab).sum(axis=1))
        for k, f in (("mean", np.nanmean), ("std", np.nanstd)):
            ref = f(fin, axis=1)
            tol = 1e-12 if k == "mean" else 1e-9
            assert np.max(np.abs(merged[k] - ref) / np.abs(ref)) < tol, k
            assert np.max(np.abs(single[k] - ref) / np.abs(ref)) < tol, k
            assert np.max(np.abs(two[k] - ref) / np.abs(ref)) < tol, k
        print("MULTIGPU_STATS_OK world=%d S=%d unstable=%d" % (world, S, int((~np.isfinite(tab[0])).sum())), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
