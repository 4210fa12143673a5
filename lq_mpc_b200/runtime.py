"""Process-wide engine cache for the drop-in layer (one Engine per CUDA device)."""
from __future__ import annotations

import os

import numpy as np

from .engine import Engine

_engines = {}


def default_device() -> int:
    if "LQMPC_DEVICE" in os.environ:
        return int(os.environ["LQMPC_DEVICE"])
    return int(os.environ.get("LOCAL_RANK", "0"))


def get_engine(device=None) -> Engine:
    d = default_device() if device is None else int(device)
    if d not in _engines:
        _engines[d] = Engine(d)
    return _engines[d]


def box_from_F(F_u):
    """{u : F_u u <= 1} -> (lo, hi). The engine's exact QP handles input boxes, which is the only constraint shape
    the reference builds (working_example_multiple.py:25: F_u = [10 I; -10 I]); a general polytope raises."""
    F_u = np.atleast_2d(np.asarray(F_u, dtype=np.float64))
    m = F_u.shape[1]
    lo, hi = np.full(m, -np.inf), np.full(m, np.inf)
    for row in F_u:
        nz = np.flatnonzero(row)
        if len(nz) != 1:
            raise NotImplementedError("lq_mpc_b200 supports box-shaped F_u (one non-zero per row) only")
        j = nz[0]
        if row[j] > 0:
            hi[j] = min(hi[j], 1.0 / row[j])
        else:
            lo[j] = max(lo[j], 1.0 / row[j])
    return lo, hi


def problem_for(A, B, Q, R, P=None, F_u=None, N_opc=30, device=None) -> Engine:
    """Engine with (A, B, Q, R, P, box(F_u)) installed as the TRUE/nominal problem."""
    eng = get_engine(device)
    lo = hi = None
    if F_u is not None:
        lo, hi = box_from_F(F_u)
    eng.set_problem(np.asarray(A, dtype=np.float64), np.asarray(B, dtype=np.float64),
                    np.asarray(Q, dtype=np.float64), np.asarray(R, dtype=np.float64),
                    None if P is None else np.asarray(P, dtype=np.float64), lo, hi, N_opc)
    return eng
