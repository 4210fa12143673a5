"""Development probe: per-phase device time of the sweep loop (synchronising after each phase)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lq_mpc_b200 import sampling as sp, stats as _stats
from lq_mpc_b200.engine import Engine
from lq_mpc_b200.sweep import _strided_columns
A = np.array([[1, 0.7], [0.12, 0.4]]); B = np.array([[1], [1.2]])
eng = Engine(0); eng.set_problem(A, B, 2 * np.eye(2), np.eye(1), 2 * np.eye(2), [-0.1], [0.1], 30)
lev = np.linspace(1e-3, 1e-2, 10)
dA, dB = sp.device_error_grids(eng, 2, 1, lev, 20000, "f")
S = dA.shape[1]
e_per = eng._dev(np.tile(lev, S // 10)); x = np.array([0.15913403, 0.15913403])
ring = eng._dev(np.stack([x, -x, 0.5 * x, x * 0.1]))
T = {"mv": 0.0, "sim": 0.0, "bounds": 0.0, "stats": 0.0}
def timed(key, fn):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize(); T[key] += time.perf_counter() - t; return r
for rep in range(2):
    for k in T: T[k] = 0.0
    for N in range(1, 51):
        mv = timed("mv", lambda: eng.mpc_solve_batch(dA, dB, N, pts=ring, want=("M_V",))["M_V"])
        sim = timed("sim", lambda: eng.simulate_batch(dA, dB, N, 30, x0_shared=x, want=("J_T", "flags")))
        b = timed("bounds", lambda: eng.bounds_batch(dA, dB, N, e_per, e_per, mv, x, (0.1, 1, 0.6), 0.2))
        def st():
            cols = torch.cat([_strided_columns(v, 10) for v in (sim["J_T"], b["bound"], b["alpha"], b["beta"], b["xi"], b["eta"], mv)], dim=0)
            return _stats.column_stats(eng, cols)
        timed("stats", st)
    print(rep, {k: round(v, 3) for k, v in T.items()}, "total", round(sum(T.values()), 3))
