"""Shim (oracle only) for python-control's `dlqr` (call sites utils_class.py:761,840,923;
working_example_single.py:39).  python-control (un-pinned dependency, absent here) solves the DARE with
scipy.linalg.solve_discrete_are (or slycot) and returns K = (R + B'PB)^-1 B'PA, the convention u = -Kx."""
import numpy as np
import scipy.linalg as sla


def dlqr(A, B, Q, R):
    A = np.atleast_2d(np.asarray(A, dtype=float))
    B = np.atleast_2d(np.asarray(B, dtype=float))
    Q = np.atleast_2d(np.asarray(Q, dtype=float))
    R = np.atleast_2d(np.asarray(R, dtype=float))
    P = sla.solve_discrete_are(A, B, Q, R)
    K = np.linalg.solve(R + B.T @ P @ B, B.T @ P @ A)
    E = np.linalg.eigvals(A - B @ K)
    return K, P, E
