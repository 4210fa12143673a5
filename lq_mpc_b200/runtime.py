"""Process-wide engine cache for the drop-in layer (one Engine per CUDA device)."""
from __future__ import annotations

import os

import numpy as np

from .engine import Engine, EngineError

_engines = {}


def default_device() -> int:
    if "LQMPC_DEVICE" in os.environ:
        return int(os.environ["LQMPC_DEVICE"])
    return int(os.environ.get("LOCAL_RANK", "0"))


def get_engine(device=None, scratch: bool = False) -> Engine:
    """The process-wide engine of a device, or (scratch=True) a SECOND context on the same device that the stateless
    helper functions of `utils` / `control` install their throw-away problems in (my_eigen, local_radius, geo_M,
    ex_stability_lq, dlqr, ...): a helper call never disturbs the problem, references or polytope a caller installed
    on the main engine."""
    d = default_device() if device is None else int(device)
    key = (d, bool(scratch))
    if key not in _engines:
        _engines[key] = Engine(d)
    return _engines[key]


def box_from_F(F_u):
    """{u : F_u u <= 1} -> (lo, hi). The engine's exact QP handles input boxes, which is the only constraint shape
    the reference builds (working_example_multiple.py:25: F_u = [10 I; -10 I]); a general polytope raises."""
    F_u = np.atleast_2d(np.asarray(F_u, dtype=np.float64))
    m = F_u.shape[1]
    lo, hi = np.full(m, -np.inf), np.full(m, np.inf)
    for row in F_u:
        nz = np.flatnonzero(row)
        if len(nz) != 1:
            raise NotImplementedError("not a box: this row of F_u couples several inputs (general polytopes go "
                                      "through Engine.set_input_polytope / runtime.problem_for)")
        j = nz[0]
        if row[j] > 0:
            hi[j] = min(hi[j], 1.0 / row[j])
        else:
            lo[j] = max(lo[j], 1.0 / row[j])
    return lo, hi


def bind_to_gpu_numa(device_index: int):
    """Pin the calling process to the CPU cores NVML reports as local to GPU `device_index` (its NUMA node), so that
    host buffers allocated afterwards (first touch) sit behind the same root complex as the GPU: the host-buffer entry
    points (`lqmpc_eval_batch_host`, `..._tiled_host`) are PCIe-bound, and with one process per GPU the copies of all
    ranks otherwise meet on one socket's memory controllers. Returns the core list, or None when NVML is unavailable
    (the process is then left as it was)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = device_index
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if device_index < len(ids) and ids[device_index].isdigit():
                idx = int(ids[device_index])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return None


def is_box(F_u) -> bool:
    """Every row of F_u has exactly one non-zero (the only shape the reference's own scripts build)."""
    F_u = np.atleast_2d(np.asarray(F_u, dtype=np.float64))
    return all(len(np.flatnonzero(row)) == 1 for row in F_u)


def polytope_vertices(F_u):
    """Vertices of {u : F_u u <= 1}: feasible intersections of m rows (m <= 4, p <= 12: at most 495 candidates).
    Host-side and once per problem, like the Gurobi models it stands in for (utils.py:592-650)."""
    import itertools
    F = np.atleast_2d(np.asarray(F_u, dtype=np.float64))
    p, m = F.shape
    out = []
    for rows in itertools.combinations(range(p), m):
        Fs = F[list(rows)]
        if abs(np.linalg.det(Fs)) < 1e-12 * np.prod(np.linalg.norm(Fs, axis=1)):
            continue
        v = np.linalg.solve(Fs, np.ones(m))
        if np.all(F @ v <= 1 + 1e-9):
            out.append(v)
    if not out:
        raise ValueError("input set is unbounded or empty")
    # a bounded polytope: every direction is blocked by some row (checked on the vertex set's recession: any row
    # combination with no vertex would have raised above; an unbounded set with vertices is caught here)
    V = np.array(out)
    for d in np.vstack((np.eye(m), -np.eye(m))):
        if not np.any(F @ d > 1e-12):
            raise ValueError("input set is unbounded")
    return V


def problem_for(A, B, Q, R, P=None, F_u=None, N_opc=30, device=None, scratch: bool = False) -> Engine:
    """Engine with (A, B, Q, R, P, F_u) installed as the TRUE/nominal problem: a box-shaped F_u goes into the problem
    itself, a general polytope is installed behind it (lqmpc_set_input_polytope). scratch=True: the helper context
    (see get_engine)."""
    eng = get_engine(device, scratch)
    lo = hi = None
    general = F_u is not None and not is_box(F_u)
    if F_u is not None and not general:
        lo, hi = box_from_F(F_u)
    eng.set_problem(np.asarray(A, dtype=np.float64), np.asarray(B, dtype=np.float64),
                    np.asarray(Q, dtype=np.float64), np.asarray(R, dtype=np.float64),
                    None if P is None else np.asarray(P, dtype=np.float64), lo, hi, N_opc)
    if general:
        if np.atleast_2d(np.asarray(F_u)).shape[1] != eng.m:
            raise EngineError("F_u has %d columns, the plant has %d inputs"
                              % (np.atleast_2d(np.asarray(F_u)).shape[1], eng.m))
        V = polytope_vertices(F_u)
        bar_u = float(np.max(np.sum(V * V, axis=1)))
        D = V[:, None, :] - V[None, :, :]
        eng.set_input_polytope(F_u, bar_u=bar_u, bar_d_u=float(np.max(np.sum(D * D, axis=2))))
    return eng
