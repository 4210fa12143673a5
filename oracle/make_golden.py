"""TEST INFRASTRUCTURE ONLY — generates the committed fixtures under tests/golden/.

Run in the build container (needs /root/reference mounted):   python -m oracle.make_golden
Everything written here comes from the *untouched* reference modules (imported behind oracle/shims) or is a
byte-for-byte array copy of the reference's own shipped data files (its only pinned results).

  tests/golden/multiple_sys.npz       the 13 arrays of data_lq_mpc_multipleSys.npz (authors' run: real cvxpy /
                                      control / Gurobi) + the four error_{A,B}_{f,2}.npy input grids
  tests/golden/ref_known_answers.json scalar/array answers of working_example_single.py, mpc_test.py and
                                      behavior_test.py scenarios + randomised calls of the reference API
  tests/golden/ref_norm2_subset.npz   reference data_generation on the first 5 systems of the `_2` grids
"""
from __future__ import annotations

import json
import os

import numpy as np

from . import ref_oracle as ro

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _l(x):
    return np.asarray(x, dtype=float).tolist()


def copy_shipped():
    out = {}
    g = np.load(os.path.join(ro.REFERENCE_DIR, "data_lq_mpc_multipleSys.npz"))
    for k in g.files:
        out[k] = g[k]
    for f in ("error_A_f", "error_B_f", "error_A_2", "error_B_2"):
        out[f] = np.load(os.path.join(ro.REFERENCE_DIR, f + ".npy"))
    np.savez_compressed(os.path.join(GOLD, "multiple_sys.npz"), **out)


def single_example(u, uc, ct):
    """working_example_single.py:20-108, evaluated through the reference functions."""
    A = np.array([[1, 0.7], [0.12, 0.4]])
    B = np.array([[1], [1.2]])
    Q, R = 2 * np.eye(2), np.eye(1)
    F_u = np.array([[10], [-10]])
    K_lqr, P_lqr, _ = ct.dlqr(A, B, Q, R)
    eps = u.local_radius(F_u, -K_lqr, Q)
    x0_vec = u.circle_generator(8, 1.5, eps, Q)
    N = 6
    x_ref, u_ref = np.zeros((2, N)), np.zeros((1, N))
    beh = uc.LQ_RDP_Behavior(A, B, Q, R, F_u, -K_lqr, 6, 10, -6, -2)
    M_V = beh.OL_energy_bound(N, 8, 1.5, x_ref, u_ref)
    ex = u.ex_stability_lq(A, B, Q, R, -K_lqr)
    bar = u.ex_stability_bounds(ex['gamma'], eps, M_V)
    err = u.fc_omega_eta(N, A, B, Q, R, -K_lqr, bar['L_V'], bar['N_0'])
    calc = uc.LQ_RDP_Calculator(A, B, Q, R, F_u)
    dec = calc.energy_decreasing(N, 0.01, 0.01, -K_lqr, M_V)
    p = np.array([0.1, 1, 0.6])
    bnd = calc.energy_bound(N, 0.01, 0.01, x0_vec[:, 0], p)
    E = u.fc_ec_E(N, 0.01, 0.01, A, B, Q, R, x0_vec[:, 0], u.bar_u_solve(F_u), u.bar_d_u_solve(F_u))
    th = u.fc_ec_theta(N, 0.01, 0.01, A, B, 2.0)
    # per-point open-loop solves of the ring (utils_class.py:459-461)
    mpc = uc.LQ_MPC_Controller(N, A, B, Q, R, Q, F_u)
    ring = [mpc.solve(x0_vec[:, i], x_ref, u_ref) for i in range(8)]
    return {'K_lqr': _l(K_lqr), 'P_lqr': _l(P_lqr), 'epsilon_lqr': float(eps), 'x0_vec': _l(x0_vec),
            'M_V': float(M_V), 'ex': {k: float(v) for k, v in ex.items()},
            'bar': {'L_V': float(bar['L_V']), 'N_0': int(bar['N_0'])},
            'omega_eta': {k: float(v) for k, v in err.items()},
            'decrease': {k: float(v) for k, v in dec.items()}, 'bound': {k: float(v) for k, v in bnd.items()},
            'E': {k: float(v) for k, v in E.items()}, 'theta': {k: float(v) for k, v in th.items()},
            'ring_V': [float(r['V_N']) for r in ring], 'ring_u0': [_l(r['u_0']) for r in ring],
            'bar_u': float(u.bar_u_solve(F_u)), 'bar_d_u': float(u.bar_d_u_solve(F_u))}


def mpc_test_example(u, uc):
    """mpc_test.py:13-56."""
    A = np.array([[1, 0.7], [0.12, 0.4]])
    B = np.array([[1], [1.2]])
    Q, R = 2 * np.eye(2), np.eye(1)
    x0 = np.array([0.1125, 0.19])
    F_u = np.vstack((10 * np.eye(1), -10 * np.eye(1)))
    N_open = 20
    info = uc.LQ_MPC_Controller(N_open, A, B, Q, R, Q, F_u).solve(x0, np.zeros((2, N_open)), np.zeros((1, N_open)))
    A_true = np.array([[1.01, 0.7], [0.12, 0.41]])
    B_true = np.array([[1], [1.21]])
    sim = uc.LQ_MPC_Simulator(20, 6, A, B, Q, R, Q, F_u).simulate(x0, A_true, B_true, np.zeros((2, N_open)),
                                                                 np.zeros((1, N_open)))
    return {'V_N': float(info['V_N']), 'u_0': _l(info['u_0']), 'J_T': float(sim['J_T']), 'X': _l(sim['X']),
            'U': _l(sim['U'])}


def behavior_example(u, uc, ct):
    """behavior_test.py:15-103 up to its plotting tail (including the reference's indexing quirks)."""
    A = np.array([[1, 0.7], [0.12, 0.4]])
    B = np.array([[1], [1.2]])
    Q, R = 2 * np.eye(2), np.eye(1)
    F_u = np.array([[10], [-10]])
    K_lqr, _, _ = ct.dlqr(A, B, Q, R)
    eps = u.local_radius(F_u, -K_lqr, Q)
    N = 6
    x_ref, u_ref = np.zeros((2, N)), np.zeros((1, N))
    err_nominal = {'e_A': 0.01, 'e_B': 0.01}
    p = np.array([0.1, 1, 0.6])
    beh = uc.LQ_RDP_Behavior(A, B, Q, R, F_u, -K_lqr, 6, 10, -6, -2)
    M_V = beh.OL_energy_bound(N, 8, 1.5, x_ref, u_ref)
    x0_vec = u.circle_generator(8, 1.5, eps, Q)
    data_xi = beh.data_generation_xi(-K_lqr, M_V, N, err_nominal)
    data_alpha, data_beta = beh.data_generation_alpha_beta(x0_vec[:, 1], p, N, err_nominal)
    A_true = np.array([[1.01, 0.7], [0.12, 0.41]])
    B_true = np.array([[1], [1.21]])
    info_ref = {'x_ref': x_ref, 'u_ref': u_ref, 'x_ref_long': np.zeros((2, 20)), 'u_ref_long': np.zeros((1, 20))}
    quad = {'x': np.array([0.12, 0.16]), 'y': np.array([0.12, 0.16])}
    mesh = beh.data_generation_mesh(N, {'T_mpc': 20, 'N_opc': 20}, {'A_true': A_true, 'B_true': B_true},
                                    err_nominal, info_ref, M_V, p, quad)
    ratio_vec = np.arange(1.1, 1.6, 0.1)[:2]
    plane = beh.data_generation_plane(N, {'T_mpc': 20, 'N_opc': 20}, {'A_true': A_true, 'B_true': B_true},
                                      err_nominal, info_ref, M_V, p, 8, ratio_vec)
    return {'M_V': float(M_V), 'xi': {k: _l(v) for k, v in data_xi.items()},
            'alpha': {k: _l(v) for k, v in data_alpha.items()}, 'beta': {k: _l(v) for k, v in data_beta.items()},
            'mesh': {k: _l(v) for k, v in mesh.items()}, 'plane': {k: _l(v) for k, v in plane.items()},
            'plane_ratio': _l(ratio_vec)}


def random_api_cases(u, uc, ct, seed=7):
    """Randomised calls of the reference classes (n in {2,3}, m in {1,2}); m=2 only where the reference runs."""
    rng = np.random.default_rng(seed)
    cases = []
    for c in range(24):
        n = 2 if c % 2 == 0 else 3
        m = 1 if c % 3 else 2
        N = int(rng.integers(1, 13))
        T = int(rng.integers(3, 16))
        A = rng.normal(size=(n, n))
        A *= rng.uniform(0.6, 1.15) / np.max(np.abs(np.linalg.eigvals(A)))
        B = rng.normal(size=(n, m))
        q, r = rng.uniform(0.5, 3.0), rng.uniform(0.3, 2.0)
        Q, R = q * np.eye(n), r * np.eye(m)
        ub = rng.uniform(0.05, 0.4)
        F_u = np.vstack((np.eye(m) / ub, -np.eye(m) / ub))
        x0 = rng.normal(size=n) * rng.uniform(0.05, 0.6)
        dA = rng.uniform(-0.02, 0.02, size=(n, n))
        dB = rng.uniform(-0.02, 0.02, size=(n, m))
        xr, ur = np.zeros((n, N)), np.zeros((m, N))
        sol = uc.LQ_MPC_Controller(N, A + dA, B + dB, Q, R, Q, F_u).solve(x0, xr, ur)
        sim = uc.LQ_MPC_Simulator(T, N, A + dA, B + dB, Q, R, Q, F_u).simulate(x0, A, B, xr, ur)
        rec = {'n': n, 'm': m, 'N': N, 'T': T, 'A': _l(A), 'B': _l(B), 'dA': _l(dA), 'dB': _l(dB), 'q': q, 'r': r,
               'ub': ub, 'x0': _l(x0), 'u_0': _l(sol['u_0']), 'V_N': float(sol['V_N']), 'J_T': float(sim['J_T']),
               'X': _l(sim['X']), 'U': _l(sim['U'])}
        if m == 1 and N >= 2:
            K, _, _ = ct.dlqr(A + dA, B + dB, Q, R)
            calc = uc.LQ_RDP_Calculator(A + dA, B + dB, Q, R, F_u)
            e = float(rng.uniform(1e-3, 2e-2))
            M_V = float(sol['V_N'] * rng.uniform(1.0, 3.0))
            try:
                dec = calc.energy_decreasing(N, e, e, -K, M_V)
                bnd = calc.energy_bound(N, e, e, x0, np.array([0.1, 1, 0.6]))
                rec.update({'e': e, 'M_V': M_V, 'K_dlqr': _l(K), 'xi': float(dec['xi']), 'eta': float(dec['eta']),
                            'alpha': float(bnd['alpha']), 'beta': float(bnd['beta'])})
            except ValueError as ex:      # math domain error (utils.py:506-507) when rho_K >= 1 etc.
                rec.update({'e': e, 'M_V': M_V, 'K_dlqr': _l(K), 'raises': str(ex)})
        cases.append(rec)
    return cases


def norm2_subset(uc):
    """Reference data_generation on the `_2` grids, N_sys = 5*1 systems per column (utils_class.py:726)."""
    cfg = ro.example_multiple_config()
    with ro.reference_cwd():
        beh = uc.LQ_RDP_Behavior_Multiple(cfg['info_opc'], cfg['info_N'], cfg['info_e'], 1, '2')
        out = beh.data_generation(cfg['N_points'], cfg['ext_radius_max'], cfg['info_ref'], cfg['p'])
    np.savez_compressed(os.path.join(GOLD, "ref_norm2_subset.npz"), **out)


def main():
    os.makedirs(GOLD, exist_ok=True)
    u, uc = ro.load()
    import importlib.util
    spec = importlib.util.spec_from_file_location("_shim_control", os.path.join(HERE, "shims", "control.py"))
    ct = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ct)
    copy_shipped()
    ans = {'single': single_example(u, uc, ct), 'mpc_test': mpc_test_example(u, uc),
           'behavior': behavior_example(u, uc, ct), 'random_api': random_api_cases(u, uc, ct)}
    with open(os.path.join(GOLD, "ref_known_answers.json"), "w") as f:
        json.dump(ans, f, indent=0)
    norm2_subset(uc)
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    import sys as _sys
    if not any(a in _sys.argv for a in ("--tracking", "--polytope", "--round2")):
        main()


def tracking_cases(uc, seed=11):
    """Non-zero references (utils_class.py:62-81: (x_{i+1} - x_ref[:, i])' Q~ (.) + (u_i - u_ref[:, i])' R (.)) through
    the untouched reference classes: open-loop solves and closed loops, saturated and not."""
    rng = np.random.default_rng(seed)
    cases = []
    for c in range(16):
        n = 2 if c % 2 == 0 else 3
        m = 1 if c % 3 else 2
        N = int(rng.integers(2, 11))
        T = int(rng.integers(3, 10))
        A = rng.normal(size=(n, n))
        A *= rng.uniform(0.6, 1.1) / np.max(np.abs(np.linalg.eigvals(A)))
        B = rng.normal(size=(n, m))
        q, r = rng.uniform(0.5, 3.0), rng.uniform(0.3, 2.0)
        Q, R = q * np.eye(n), r * np.eye(m)
        ub = rng.uniform(0.05, 0.5)
        F_u = np.vstack((np.eye(m) / ub, -np.eye(m) / ub))
        x0 = rng.normal(size=n) * rng.uniform(0.05, 0.6)
        dA = rng.uniform(-0.02, 0.02, size=(n, n))
        dB = rng.uniform(-0.02, 0.02, size=(n, m))
        xr = rng.normal(size=(n, N + 2)) * rng.uniform(0.0, 0.4)        # wider than N: only the first N columns count
        ur = rng.normal(size=(m, N + 2)) * rng.uniform(0.0, 0.3)
        sol = uc.LQ_MPC_Controller(N, A + dA, B + dB, Q, R, Q, F_u).solve(x0, xr, ur)
        sim = uc.LQ_MPC_Simulator(T, N, A + dA, B + dB, Q, R, Q, F_u).simulate(x0, A, B, xr, ur)
        cases.append({'n': n, 'm': m, 'N': N, 'T': T, 'A': _l(A), 'B': _l(B), 'dA': _l(dA), 'dB': _l(dB), 'q': q,
                      'r': r, 'ub': ub, 'x0': _l(x0), 'x_ref': _l(xr), 'u_ref': _l(ur), 'u_0': _l(sol['u_0']),
                      'V_N': float(sol['V_N']), 'J_T': float(sim['J_T']), 'X': _l(sim['X']), 'U': _l(sim['U'])})
    return cases


def polytope_cases(u, uc, seed=23):
    """General input polytopes F_u (rows coupling several inputs, utils_class.py:81) through the untouched reference
    classes: open-loop solves, closed loops, bar_u / bar_d_u (vertex maxima) and energy_bound; m = 2 and 3.
    (energy_decreasing cannot be pinned here: the reference's `A + B * K`, utils.py:356, raises for m > 1.)"""
    rng = np.random.default_rng(seed)
    cases = []
    shapes = [(2, 2, 3), (2, 2, 5), (3, 2, 4), (3, 2, 6), (2, 2, 8), (3, 3, 4), (3, 3, 7), (2, 2, 4)]
    for c in range(16):
        n, m, p = shapes[c % len(shapes)]
        N = int(rng.integers(2, min(10, 128 // p) + 1))
        T = int(rng.integers(3, 9))
        A = rng.normal(size=(n, n))
        A *= rng.uniform(0.6, 1.15) / np.max(np.abs(np.linalg.eigvals(A)))
        B = rng.normal(size=(n, m))
        q, r = rng.uniform(0.5, 3.0), rng.uniform(0.3, 2.0)
        Q, R = q * np.eye(n), r * np.eye(m)
        while True:                                   # bounded polytope around the origin: unit normals / support
            D = rng.normal(size=(p, m))
            D /= np.linalg.norm(D, axis=1, keepdims=True)
            probe = rng.normal(size=(4000, m))
            probe /= np.linalg.norm(probe, axis=1, keepdims=True)
            if np.min(np.max(probe @ D.T, axis=1)) > 0.1:
                break
        F_u = D / rng.uniform(0.1, 0.45, size=(p, 1))
        x0 = rng.normal(size=n) * rng.uniform(0.3, 1.5)
        dA = rng.uniform(-0.02, 0.02, size=(n, n))
        dB = rng.uniform(-0.02, 0.02, size=(n, m))
        zx, zu = np.zeros((n, N)), np.zeros((m, N))
        sol = uc.LQ_MPC_Controller(N, A + dA, B + dB, Q, R, Q, F_u).solve(x0, zx, zu)
        sim = uc.LQ_MPC_Simulator(T, N, A + dA, B + dB, Q, R, Q, F_u).simulate(x0, A, B, zx, zu)
        e = float(rng.uniform(1e-3, 1e-2))
        bnd = uc.LQ_RDP_Calculator(A + dA, B + dB, Q, R, F_u).energy_bound(N, e, e, x0, np.array([0.1, 1, 0.6]))
        cases.append({'n': n, 'm': m, 'p': p, 'N': N, 'T': T, 'A': _l(A), 'B': _l(B), 'dA': _l(dA), 'dB': _l(dB),
                      'q': q, 'r': r, 'F_u': _l(F_u), 'x0': _l(x0), 'u_0': _l(sol['u_0']), 'V_N': float(sol['V_N']),
                      'J_T': float(sim['J_T']), 'X': _l(sim['X']), 'U': _l(sim['U']), 'e': e,
                      'bar_u': float(u.bar_u_solve(F_u)), 'bar_d_u': float(u.bar_d_u_solve(F_u)),
                      'alpha': float(bnd['alpha']), 'beta': float(bnd['beta'])})
    return cases


def main_polytope():
    """Adds tests/golden/ref_polytope_cases.json without touching the other fixtures."""
    u, uc = ro.load()
    with open(os.path.join(GOLD, "ref_polytope_cases.json"), "w") as f:
        json.dump(polytope_cases(u, uc), f, indent=0)
    print("polytope fixtures written")


def main_tracking():
    """Adds tests/golden/ref_tracking_cases.json without touching the other fixtures."""
    u, uc = ro.load()
    with open(os.path.join(GOLD, "ref_tracking_cases.json"), "w") as f:
        json.dump(tracking_cases(uc), f, indent=0)
    print("tracking fixtures written")


if __name__ == "__main__":
    import sys as _sys
    if "--tracking" in _sys.argv:
        main_tracking()
    if "--polytope" in _sys.argv:
        main_polytope()


def extension_cases(u, uc, ct, seed=31):
    """energy_decreasing_extension / fc_omega_eta_extension (utils_class.py:375-406, utils.py:412-466) through the
    untouched reference, m = 1 (its `A + B * K` is an outer product only then): the shipped 2-state example with the
    LQR gain and detuned second gains, random well-damped plants, and cases whose math.log argument is <= 0."""
    rng = np.random.default_rng(seed)
    cases = []
    A0 = np.array([[1, 0.7], [0.12, 0.4]])
    B0 = np.array([[1], [1.2]])
    for c in range(14):
        if c < 4:
            n, A, B, q, r, ub = 2, A0, B0, 2.0, 1.0, 0.1
        else:
            n = 2 if c % 2 == 0 else 3
            A = rng.normal(size=(n, n))
            A *= rng.uniform(0.3, 0.95) / np.max(np.abs(np.linalg.eigvals(A)))
            B = rng.normal(size=(n, 1))
            q, r, ub = rng.uniform(0.5, 3.0), rng.uniform(0.3, 2.0), rng.uniform(0.05, 0.4)
        Q, R = q * np.eye(n), r * np.eye(1)
        F_u = np.array([[1 / ub], [-1 / ub]])
        K_lqr, _, _ = ct.dlqr(A, B, Q, R)
        K = -K_lqr
        # second gain: the LQR gain of a detuned weight pair (c == 0: the same gain)
        r2 = r * (1.0 if c == 0 else rng.uniform(0.2, 5.0))
        hatK = -ct.dlqr(A, B, Q, r2 * np.eye(1))[0]
        N = int(rng.integers(3, 12))
        e = float(rng.uniform(1e-3, 2e-2))
        M_V = float(rng.uniform(0.02, 1.5))
        if c in (3, 13):                      # a gain that leaves rho(A + BK) + 0.4 >= 1: log of a negative number
            K = np.zeros((1, n)) if c == 3 else 0.05 * K
        rec = {'n': n, 'A': _l(A), 'B': _l(B), 'q': q, 'r': r, 'ub': ub, 'K': _l(K), 'hatK': _l(hatK), 'N': N,
               'e': e, 'M_V': M_V}
        calc = uc.LQ_RDP_Calculator(A, B, Q, R, F_u)
        try:
            eps = u.local_radius(F_u, K, Q)
            st = u.ex_stability_lq(A, B, Q, R, K)
            bd = u.ex_stability_bounds(st['gamma'], eps, M_V)
            oe = u.fc_omega_eta_extension(N, A, B, Q, R, K, hatK, bd['L_V'], bd['N_0'])
            dec = calc.energy_decreasing_extension(N, e, e, K, hatK, M_V)
            rec.update({'L_V': float(bd['L_V']), 'N_0': int(bd['N_0']),
                        'omega_eta': {k: float(v) for k, v in oe.items()},
                        'xi': float(dec['xi']), 'eta': float(dec['eta'])})
        except ValueError as ex:
            rec['raises'] = str(ex)
        cases.append(rec)
    return cases


def cfg4_cases(uc, seed=41):
    """BASELINE configs[3] shape (n = 4, m = 2, N = 10, Q = I, R = I, the seed-0 synthetic plant of
    lq_mpc_b200/sampling.py) through the UNTOUCHED LQ_MPC_Controller / LQ_MPC_Simulator (utils_class.py:48-91, 245-285)
    with a loose input box (never active, so the cvxpy shim's QP is the unconstrained one): the first-step gain K0
    (u_0 for the four unit states), J_T for T = 400 (= J_inf to rounding: rho^800 ~ 0) and the last state."""
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    from lq_mpc_b200 import sampling as sp
    A, B, Q, R = sp.synth_problem(4, 2, seed=0)
    rng = np.random.default_rng(seed)
    n, m, N, T = 4, 2, 10, 400
    F_u = np.vstack((np.eye(m) / 1e6, -np.eye(m) / 1e6))
    zx, zu = np.zeros((n, N)), np.zeros((m, N))
    cases = []
    for c in range(12):
        e = 0.01 if c < 8 else 0.05
        dA = rng.uniform(-e, e, size=(n, n))
        dB = rng.uniform(-e, e, size=(n, m))
        x0 = rng.normal(size=n)
        ctl = uc.LQ_MPC_Controller(N, A + dA, B + dB, Q, R, Q, F_u)
        K0 = np.column_stack([ctl.solve(np.eye(n)[:, i], zx, zu)['u_0'] for i in range(n)])
        sol = ctl.solve(x0, zx, zu)
        sim = uc.LQ_MPC_Simulator(T, N, A + dA, B + dB, Q, R, Q, F_u).simulate(x0, A, B, zx, zu)
        cases.append({'dA': _l(dA), 'dB': _l(dB), 'x0': _l(x0), 'K0': _l(K0), 'V_N': float(sol['V_N']),
                      'u_0': _l(sol['u_0']), 'J_T': float(sim['J_T']), 'x_T': _l(sim['X'][:, -1]),
                      'U_head': _l(sim['U'][:, :5])})
    return {'n': n, 'm': m, 'N': N, 'T': T, 'A': _l(A), 'B': _l(B), 'cases': cases}


def main_round2():
    """Adds tests/golden/ref_extension_cases.json and ref_cfg4_cases.json without touching the other fixtures."""
    u, uc = ro.load()
    import importlib.util
    spec = importlib.util.spec_from_file_location("_shim_control", os.path.join(HERE, "shims", "control.py"))
    ct = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ct)
    with open(os.path.join(GOLD, "ref_extension_cases.json"), "w") as f:
        json.dump(extension_cases(u, uc, ct), f, indent=0)
    with open(os.path.join(GOLD, "ref_cfg4_cases.json"), "w") as f:
        json.dump(cfg4_cases(uc), f, indent=0)
    print("extension + cfg4 fixtures written")


if __name__ == "__main__":
    import sys as _sys
    if "--round2" in _sys.argv:
        main_round2()
