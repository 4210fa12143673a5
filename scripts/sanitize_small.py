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
