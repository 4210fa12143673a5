"""Development probe (GPU box): the three sweep kernels (K2a ring solves, K2b closed loops, K3 bounds) at ONE horizon
on `--samples` seeded (perturbation, level) pairs of the 2-state example, timed with CUDA events. Used for the ncu
captures of those kernels (profiles/r02_k2a_*, r02_k2b_*, r02_k3_*). Not part of the product."""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lq_mpc_b200 import sampling as sp
from lq_mpc_b200.engine import Engine
from lq_mpc_b200.utils import circle_generator

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=50)
ap.add_argument("--samples", type=int, default=1_000_000)
ap.add_argument("--reps", type=int, default=1)
a = ap.parse_args()
A = np.array([[1, 0.7], [0.12, 0.4]]); B = np.array([[1], [1.2]])
eng = Engine(0)
eng.set_problem(A, B, 2 * np.eye(2), np.eye(1), 2 * np.eye(2), [-0.1], [0.1], 30)
lev = np.linspace(1e-3, 1e-2, 10)
per = a.samples // 10
dA, dB = sp.device_error_grids(eng, 2, 1, lev, per // 5, "f")
S = dA.shape[1]
e_per = eng._dev(np.tile(lev, S // 10))
K = eng.dlqr_batch(S=1)["K"].cpu().numpy()[:, 0].reshape(1, 2)
eps = 1.0 / (float(K[0] @ K[0]) / 2.0 / 0.01)
ring = circle_generator(8, 1.5, eps, 2 * np.eye(2))
x = ring[:, 1].copy()
ring_d = eng._dev(ring.T.copy())


def once(dA, dB, e_per):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    mv = eng.mpc_solve_batch(dA, dB, a.N, pts=ring_d, want=("M_V", "flags"))
    ev[1].record()
    sim = eng.simulate_batch(dA, dB, a.N, 30, x0_shared=x, want=("J_T", "flags"))
    ev[2].record()
    b = eng.bounds_batch(dA, dB, a.N, e_per, e_per, mv["M_V"], x, (0.1, 1, 0.6), 0.2)
    ev[3].record()
    torch.cuda.synchronize()
    return [ev[i].elapsed_time(ev[i + 1]) for i in range(3)], b


once(dA[:, :4096].contiguous(), dB[:, :4096].contiguous(), e_per[:4096].contiguous())     # warm-up (3 launches)
best = None
for _ in range(a.reps):
    t, b = once(dA, dB, e_per)
    best = t if best is None else [min(u, v) for u, v in zip(best, t)]
void = int(((b["flags"] & (32 | 512)) != 0).sum())
print(json.dumps({"N": a.N, "samples": S, "ms_ring_solves": best[0], "ms_simulate": best[1], "ms_bounds": best[2],
                  "void_bounds": void}))
