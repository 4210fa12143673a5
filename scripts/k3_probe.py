"""Development probe (GPU box): K3 / K3g timing at one horizon. Not part of the product."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lq_mpc_b200 import sampling as sp
from lq_mpc_b200.engine import Engine
N = int(sys.argv[1]) if len(sys.argv) > 1 else 36
S = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
A = np.array([[1, 0.7], [0.12, 0.4]]); B = np.array([[1], [1.2]])
eng = Engine(0)
eng.set_problem(A, B, 2 * np.eye(2), np.eye(1), 2 * np.eye(2), [-0.1], [0.1], 30)
g = torch.Generator(device="cuda").manual_seed(0)
dA = (torch.rand((4, S), device="cuda", dtype=torch.float64, generator=g) - 0.5) * 0.01
dB = (torch.rand((2, S), device="cuda", dtype=torch.float64, generator=g) - 0.5) * 0.01
x = np.array([0.159, 0.159])
for rep in range(3):
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    b = eng.bounds_batch(dA, dB, N, 0.005, 0.005, 0.25, x, (0.1, 1, 0.6), 0.2)
    e1.record(); torch.cuda.synchronize()
    print("N", N, "S", S, "ms", e0.elapsed_time(e1), "min_H mean", float(b["min_H"].mean()))
