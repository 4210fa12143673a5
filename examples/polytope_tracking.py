"""The two parts of LQ_MPC_Controller's interface the reference's own scripts never exercise, on the engine:
a general input polytope F_u u <= 1 (rows coupling the inputs, utils_class.py:81) and non-zero state / input references
(utils_class.py:62-81). Same class names and call signatures as the reference (GPU box only)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lq_mpc_b200.utils_class import LQ_MPC_Controller, LQ_MPC_Simulator

A = np.array([[1.0, 0.7, 0.0], [0.12, 0.4, 0.3], [0.0, -0.2, 0.9]])
B = np.array([[1.0, 0.0], [1.2, 0.5], [0.0, 1.0]])
Q, R, N, T = 2 * np.eye(3), np.eye(2), 8, 25
# a pentagon around the origin: every row mixes both inputs
ang = np.linspace(0, 2 * np.pi, 5, endpoint=False) + 0.3
F_u = np.stack([np.cos(ang), np.sin(ang)], axis=1) / 0.15
x0 = np.array([0.6, -0.4, 0.5])
x_ref = np.tile(np.array([[0.05], [0.0], [-0.05]]), (1, N))
u_ref = np.zeros((2, N))

sol = LQ_MPC_Controller(N, A, B, Q, R, Q, F_u).solve(x0, x_ref, u_ref)
print("u_0 =", sol["u_0"], " V_N =", sol["V_N"], " max F_u u_0 =", float(np.max(F_u @ sol["u_0"])))
sim = LQ_MPC_Simulator(T, N, A, B, Q, R, Q, F_u).simulate(x0, 1.01 * A, B, x_ref, u_ref)   # plant differs from the model
print("closed loop: J_T = %.6f, steps on the polytope boundary: %d of %d, final state %s"
      % (sim["J_T"], int(np.sum(np.max(F_u @ sim["U"], axis=0) > 1 - 1e-9)), T, np.round(sim["X"][:, -1], 4)))
