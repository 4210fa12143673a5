"""TEST INFRASTRUCTURE ONLY — the untouched reference, imported behind shims.

This module imports `/root/reference/utils.py` and `/root/reference/utils_class.py` UNCHANGED (no source is
copied into this repo) by putting `oracle/shims/` in front of them on `sys.path`.  The shims stand in for the
four third-party packages the reference imports at module top and that are not installed in this image
(matplotlib, gurobipy, control, cvxpy — see each shim's docstring for what it restates).

It only works inside the build container (where `/root/reference` is mounted); the GPU box has no
`/root/reference`.  It is therefore used for exactly two things:
  * validating the independent restatement in `oracle/np_oracle.py`, and
  * generating the committed golden fixtures under `tests/golden/` (see `oracle/make_golden.py`).
Nothing in `lq_mpc_b200/` imports it.
"""
from __future__ import annotations

import contextlib
import itertools
import os
import sys
import warnings

import numpy as np

REFERENCE_DIR = os.environ.get("LQMPC_REFERENCE_DIR", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "utils_class.py"))


def _box_from_F(F_u):
    """Parse {u : F_u u <= 1} as a box lo <= u <= hi (rows with a single non-zero; utils_class.py:81)."""
    F_u = np.atleast_2d(np.asarray(F_u, dtype=float))
    m = F_u.shape[1]
    lo = np.full(m, -np.inf)
    hi = np.full(m, np.inf)
    for row in F_u:
        nz = np.flatnonzero(row)
        if len(nz) != 1:
            raise NotImplementedError("only box-shaped F_u is supported by the oracle")
        j = nz[0]
        b = 1.0 / row[j]
        if row[j] > 0:
            hi[j] = min(hi[j], b)
        else:
            lo[j] = max(lo[j], b)
    return lo, hi


def _is_box(F_u):
    return all(len(np.flatnonzero(r)) == 1 for r in np.atleast_2d(np.asarray(F_u, dtype=float)))


def bar_u_vertex(F_u):
    """max ||u||^2 over the box / polytope (replaces the Gurobi model of utils.py:592-619)."""
    if not _is_box(F_u):
        from .qp_dense import polytope_vertices
        return float(max(np.dot(v, v) for v in polytope_vertices(F_u)))
    lo, hi = _box_from_F(F_u)
    return float(max(np.dot(v, v) for v in itertools.product(*zip(lo, hi))))


def bar_d_u_vertex(F_u):
    """max ||u1-u2||^2 over box x box (replaces the Gurobi model of utils.py:622-650)."""
    if not _is_box(F_u):
        from .qp_dense import polytope_vertices
        verts = list(polytope_vertices(F_u))
    else:
        lo, hi = _box_from_F(F_u)
        verts = [np.array(v) for v in itertools.product(*zip(lo, hi))]
    return float(max(np.dot(v - w, v - w) for v in verts for w in verts))


_loaded = None


def load():
    """Import the reference modules (once) and patch the two Gurobi entry points. Returns (utils, utils_class)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference not mounted at %s" % REFERENCE_DIR)
    sys.dont_write_bytecode = True  # the reference directory is read-only
    saved = list(sys.path)
    saved_mods = {k: sys.modules.get(k) for k in ("utils", "utils_class", "control", "cvxpy", "gurobipy",
                                                  "matplotlib", "matplotlib.pyplot", "mpl_toolkits",
                                                  "mpl_toolkits.axes_grid1",
                                                  "mpl_toolkits.axes_grid1.inset_locator")}
    for k in saved_mods:
        sys.modules.pop(k, None)
    sys.path[:0] = [_SHIMS, REFERENCE_DIR]
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", SyntaxWarning)
            import utils as ref_utils  # noqa
            import utils_class as ref_utils_class  # noqa
    finally:
        sys.path[:] = saved
        # keep the reference modules private to this oracle: the product package ships modules of the same name
        for k, v in saved_mods.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v
    for mod in (ref_utils, ref_utils_class):
        mod.bar_u_solve = bar_u_vertex
        mod.bar_d_u_solve = bar_d_u_vertex
    _loaded = (ref_utils, ref_utils_class)
    return _loaded


@contextlib.contextmanager
def reference_cwd():
    """The reference reads `error_*.npy` relative to cwd (utils_class.py:749-750) and writes its result file there
    (utils_class.py:958); run it from the reference directory with np.savez/np.save stubbed (read-only mount)."""
    old = os.getcwd()
    savez, save = np.savez, np.save
    os.chdir(REFERENCE_DIR)
    np.savez = lambda *a, **k: None
    np.save = lambda *a, **k: None
    try:
        yield
    finally:
        np.savez, np.save = savez, save
        os.chdir(old)


def example_multiple_config():
    """Constants of working_example_multiple.py:13-76 (restated, not imported: the script plots at import)."""
    A = np.array([[1, 0.7], [0.12, 0.4]])
    B = np.array([[1], [1.2]])
    n_x, n_u = 2, 1
    Q = 2 * np.eye(n_x)
    R = np.eye(n_u)
    F_u = np.vstack((10 * np.eye(n_u), -10 * np.eye(n_u)))
    info_opc = {'A': A, 'B': B, 'Q': Q, 'R': R, 'F_u': F_u}
    info_N = {'N_min': 6, 'N_max': 10, 'N_nominal': 7, 'N_opc': 30, 'N_mpc': 30}
    info_ref = {'x_ref': np.zeros([n_x, 7]), 'u_ref': np.zeros([n_u, 7]),
                'x_ref_long': np.zeros([n_x, 30]), 'u_ref_long': np.zeros([n_u, 30])}
    info_e = {'e_min': 10 ** (-3), 'e_max': 10 ** (-2), 'e_nominal': 5 * (10 ** (-3))}
    return dict(info_opc=info_opc, info_N=info_N, info_ref=info_ref, info_e=info_e, N_matrix=20,
                N_points=8, ext_radius_max=1.5, p=np.array([0.1, 1, 0.6]))


def run_data_generation(norm_type="f"):
    """LQ_RDP_Behavior_Multiple(...).data_generation(...) exactly as the commented block of
    working_example_multiple.py:98-101 would run it."""
    _, uc = load()
    cfg = example_multiple_config()
    with reference_cwd():
        beh = uc.LQ_RDP_Behavior_Multiple(cfg['info_opc'], cfg['info_N'], cfg['info_e'], cfg['N_matrix'], norm_type)
        out = beh.data_generation(cfg['N_points'], cfg['ext_radius_max'], cfg['info_ref'], cfg['p'])
    return out
