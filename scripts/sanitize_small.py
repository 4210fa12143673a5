"""Smallest end-to-end pass over every kernel family for compute-sanitizer (memcheck / racecheck) on the GPU box:
  gpurun -- 'compute-sanitizer --tool memcheck python scripts/sanitize_small.py'"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
from lq_mpc_b200.engine import Engine
from oracle import np_batched as nb

g.smoke()                                                   # K1, K2a, K2b, K3, K5 on tiny batches vs oracle/golden
eng = Engine(0)
for (n, m, N) in ((16, 4, 6), (32, 8, 5)):
    A, B, Q, R = nb.synth_problem(n, m, seed=0)
    eng.set_problem_tiled(A, B, Q, R, Q, 8)
    dA, dB, x0 = nb.synth_samples(n, m, 37, seed=1, e=1e-3)
    ref = nb.eval_batch(A, B, Q, R, Q, nb.expert_matrix(A, B, Q, R, Q, 8), dA, dB, x0, N - 1, N)
    got = eng.eval_batch_tiled(dA, dB, x0, N - 1, N, want=("J", "rho", "ratio", "flags", "V_N"))
    err = float(np.max(np.abs(got["J"].cpu().numpy() - ref["J"]) / ref["J"]))
    assert err < 1e-9, err
    one = eng.eval_batch_tiled(dA, dB, x0, N, N)
    assert float(np.max(np.abs(one["rho"].cpu().numpy()[0] - ref["rho"][1]) / ref["rho"][1])) < 1e-9
print("sanitize_small ok")
# round 2 additions: seeded entry point, matrix-free K3 at a long horizon, the run-time-dimension route (K1/K2/K3/dlqr)
eng.set_problem(*nb.synth_problem(4, 2, seed=0), None, None, None, 30)
eng.eval_seeded(1, 5, 777, 0.01, 0.01, 9, 10, want=("J", "rho", "ratio", "flags", "moments"))
A2 = np.array([[1, 0.7], [0.12, 0.4]]); B2 = np.array([[1], [1.2]])
eng.set_problem(A2, B2, 2 * np.eye(2), np.eye(1), 2 * np.eye(2), [-0.1], [0.1], 30)
rng = np.random.default_rng(0)
dA = rng.uniform(-5e-3, 5e-3, size=(4, 333)); dB = rng.uniform(-5e-3, 5e-3, size=(2, 333))
eng.bounds_batch(dA, dB, 50, 5e-3, 5e-3, 0.2, np.array([0.15, 0.1]), (0.1, 1, 0.6), 0.2)
eng.mpc_solve_batch(dA, dB, 50, pts=rng.normal(size=(8, 2)) * 0.3)
eng.simulate_batch(dA, dB, 50, 30, x0_shared=np.array([0.16, 0.16]))
for (n, m) in ((5, 2), (12, 4), (32, 8)):
    A, B, Q, R = nb.synth_problem(n, m, seed=0)
    eng.set_problem(A, B, Q, R, Q, -0.3 * np.ones(m), 0.3 * np.ones(m), 10)
    dA, dB, x0 = nb.synth_samples(n, m, 21, seed=1, e=2e-3)
    sA, sB, sx = nb.to_soa(dA, dB, x0)
    eng.eval_batch(sA, sB, sx, 3, 4, T=5, want=("J", "rho", "ratio", "flags", "V_N", "J_T", "K0"))
    eng.mpc_solve_batch(sA, sB, 4, x0=sx * 3)
    eng.simulate_batch(sA, sB, 4, 3, x0=sx * 3, want=("J_T", "X", "U", "flags", "n_active"))
    eng.bounds_batch(sA, sB, 4, 1e-3, 1e-3, 0.3, sx, (0.1, 1, 0.6), 0.2, want_K=True, want_P=True)
    eng.dlqr_batch(sA, sB)
eng.sync()
print("sanitize_small round-2 additions ok")
