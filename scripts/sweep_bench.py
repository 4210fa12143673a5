#!/usr/bin/env python
"""cfg-sweep (BASELINE.json configs[2]): the reference's 2-state example, model-error grids `_f` and `_2` extended to
`--per-level` seeded perturbations per error level, 10 levels x horizons 1..50, on one GPU (or sharded
under torchrun). Prints one JSON line per grid pair: evals/s of the full K2a + K2b + K3 + K5 sweep.
1 eval = one (perturbation, level, horizon) triple = 8 open-loop QPs + 30 closed-loop QPs + 1 DARE + bound."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lq_mpc_b200 import sampling as sp
from lq_mpc_b200.engine import Engine
from lq_mpc_b200.sweep import error_horizon_sweep

ap = argparse.ArgumentParser()
ap.add_argument("--per-level", type=int, default=100_000)
ap.add_argument("--nmax", type=int, default=50)
ap.add_argument("--norms", default="f,2")
ap.add_argument("--out", default="gpurun_out/sweep_bench.json")
ap.add_argument("--host-sampler", action="store_true", help="numpy Philox sampler on the host instead of K6")
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
A = np.array([[1, 0.7], [0.12, 0.4]]); B = np.array([[1], [1.2]])
Q = 2 * np.eye(2); R = np.eye(1); F_u = np.array([[10.0], [-10.0]])
eng = Engine(local)
eng.set_problem(A, B, Q, R, Q, [-0.1], [0.1], 30)
error_vec = np.linspace(1e-3, 1e-2, 10)
lines = []
for nt in a.norms.split(","):
    from lq_mpc_b200.stats import shard_bounds
    horizons = list(range(1, a.nmax + 1))
    t = time.time()
    if a.host_sampler:
        eA, eB = sp.seeded_error_grids(2, 1, error_vec, a.per_level // 5, nt, seed=20240522)
        wA, wB = eA[:, :, :2000], eB[:, :, :2000]
        shard = (rank, world) if world > 1 else None
    else:                       # K6: this rank's shard of the grids is generated in HBM (same samples for any sharding)
        lo, hi = shard_bounds(a.per_level, rank, world)
        eA, eB = sp.device_error_grids(eng, 2, 1, error_vec, a.per_level // 5, nt, seed=20240522, j_first=lo,
                                       N_sys=hi - lo)
        torch.cuda.synchronize()
        wA, wB = sp.device_error_grids(eng, 2, 1, error_vec, 400, nt)
        shard = None
    t_gen = time.time() - t
    error_horizon_sweep(eng, wA, wB, error_vec, [1, a.nmax], F_u, Q)                                # warm-up
    r = error_horizon_sweep(eng, eA, eB, error_vec, horizons, F_u, Q, shard=shard)
    evals = a.per_level * len(error_vec) * len(horizons)
    line = {"workload": "cfg-sweep norm=%s: %d perturbations/level x 10 levels x N=1..%d" % (nt, a.per_level, a.nmax),
            "sampler": "host numpy Philox" if a.host_sampler else "K6 device Philox4x32-10",
            "evals": evals, "seconds": r["seconds"], "evals_per_s": evals / r["seconds"], "n_gpus": world,
            "sampler_seconds": t_gen, "V_expert": r["V_expert"],
            "worst_true_ratio": float(np.nanmax(r["ratio_true_max"])),
            "true_ratio_max_by_level_N7": r["ratio_true_max"][:, 6].tolist() if a.nmax >= 7 else None,
            "n_invalid_total": float(r["n_invalid"].sum()), "launches": eng.launch_count}
    lines.append(line)
    if rank == 0:
        print(json.dumps(line))
if rank == 0:
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(lines, open(a.out, "w"), indent=1)
