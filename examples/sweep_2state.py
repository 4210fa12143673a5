"""Error-level x horizon sweep of the reference's 2-state example on the engine (BASELINE cfg 2, reduced size):
worst-case true performance ratio and the share of (level, horizon) cells whose bound is valid."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lq_mpc_b200 import sampling as sp
from lq_mpc_b200.engine import Engine
from lq_mpc_b200.sweep import error_horizon_sweep

A = np.array([[1, 0.7], [0.12, 0.4]]); B = np.array([[1], [1.2]])
Q, R, F_u = 2 * np.eye(2), np.eye(1), np.array([[10.0], [-10.0]])
eng = Engine(0).set_problem(A, B, Q, R, Q, [-0.1], [0.1], 30)
levels = np.linspace(1e-3, 1e-2, 10)
eA, eB = sp.seeded_error_grids(2, 1, levels, 2000, "f")           # 10 000 perturbations per level
r = error_horizon_sweep(eng, eA, eB, levels, range(1, 21), F_u, Q)
print("evals %d in %.3f s" % (r["evals"], r["seconds"]))
print("worst true-cost ratio J/V_expert per horizon:", np.round(r["ratio_true_max"].max(axis=0), 5))
print("share of samples with a valid bound per (level, horizon):")
print(np.round(1 - r["n_invalid"] / eA.shape[2], 2))
