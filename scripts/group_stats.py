"""Development probe (GPU box; build with LQMPC_NVCC_EXTRA=-DLQ_GROUP_STATS): squarings the lane-group K1 needs until
a sample's spectral radius is accepted, and until its warp leaves the loop."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lq_mpc_b200.engine import Engine
from lq_mpc_b200 import sampling as sp
n, m, S = int(sys.argv[1]), int(sys.argv[2]), 200000
for seed in (0, 1, 2, 3):
    eng = Engine(0)
    A, B, Q, R = sp.synth_problem(n, m, seed=seed)
    eng.set_problem(A, B, Q, R, Q, None, None, 30)
    g = torch.Generator(device="cuda").manual_seed(1)
    dA = (torch.rand((n * n, S), device="cuda", dtype=torch.float64, generator=g) - 0.5) * 0.02
    dB = (torch.rand((n * m, S), device="cuda", dtype=torch.float64, generator=g) - 0.5) * 0.02
    x0 = torch.randn((n, S), device="cuda", dtype=torch.float64, generator=g) * 0.3
    out = eng.eval_batch(dA, dB, x0, 10, 10)
    fl = out["flags"].cpu().numpy().ravel()
    acc, end = (fl >> 16) & 0xff, (fl >> 24) & 0xff
    print(seed, "rho", float(out["rho"].mean()), "accepted at: median %d p90 %d max %d never %.4f | warp exit: mean %.1f" % (
        np.median(acc[acc > 0]) if (acc > 0).any() else -1, np.percentile(acc[acc > 0], 90) if (acc > 0).any() else -1,
        acc.max(), float((acc == 0).mean()), end.mean() + 1))
