"""Drop-in for the reference module `utils_class` (same class names, constructor/method signatures and returned
dict keys; citations are into /root/reference/utils_class.py). Every numeric result comes from the CUDA engine:
single calls are S = 1 batches, the sweep drivers pack all (system, level) pairs of a table into one launch.

Not carried over: the three Plotter_* classes (presentation, out of scope) — importing them raises a clear error.
Both box-shaped and general polytopic F_u are supported (the latter up to 12 rows and N * rows <= 256, beyond which the
solve is flagged, not approximated); non-zero references are supported by the controller / simulator / M_V paths (K2);
the bound formulas have no reference terms upstream either.
"""
from __future__ import annotations

import math

import numpy as np

from . import runtime as _rt
from .sampling import grids_to_soa
from .utils import (circle_generator, error_matrix_generator, local_radius)  # noqa: F401  (re-exported names)


def _cpu(t):
    return t.detach().cpu().numpy()


def _raise_if_domain_error(flags):
    """The reference raises ValueError('math domain error') from math.log / math.sqrt (utils.py:506-507,514)."""
    from .engine import FLAG_DOMAIN_ERROR
    if int(flags) & FLAG_DOMAIN_ERROR:
        raise ValueError("math domain error")


def _raise_if_solver_failed(flags, what):
    """A solve the engine could not complete exactly (working set beyond 256 entries, R + B'PB not positive definite,
    DARE / Lyapunov / eigenvalue iteration not converged) must not flow silently into tables: the reference's third
    party solvers raise in these situations (cvxpy SolverError, scipy LinAlgError)."""
    from .engine import (EngineError, FLAG_CHOL_FAIL, FLAG_DARE_NOCONV, FLAG_EIG_NOCONV, FLAG_LYAP_NOCONV,
                         FLAG_QP_MAXITER)
    bad = int(np.bitwise_or.reduce(np.asarray(flags, dtype=np.int64).reshape(-1), initial=0)) & (
        FLAG_QP_MAXITER | FLAG_CHOL_FAIL | FLAG_DARE_NOCONV | FLAG_LYAP_NOCONV | FLAG_EIG_NOCONV)
    if bad:
        raise EngineError("%s: the engine flagged an incomplete solve (flags 0x%x: QP_MAXITER=4, DARE_NOCONV=8, "
                          "LYAP_NOCONV=64, EIG_NOCONV=128, CHOL_FAIL=256)" % (what, bad))


class LQ_MPC_Controller:
    """utils_class.py:18-91 — open-loop input-constrained LQ MPC; solved exactly on the GPU (K2)."""

    def __init__(self, N, A, B, Q, R, P, F_u):
        self.N = int(N)
        self.A, self.B, self.Q, self.R, self.P, self.F_u = A, B, Q, R, P, F_u

    def solve(self, x0, x_ref, u_ref):
        eng = _rt.problem_for(self.A, self.B, self.Q, self.R, self.P, self.F_u)
        x0 = np.asarray(x0, dtype=np.float64).reshape(1, -1)
        with eng.references(x_ref, u_ref):        # utils_class.py:62-81; all-zero references = regulation
            out = eng.mpc_solve_batch(None, None, self.N, pts=x0, S=1, want=("V", "u0", "flags"))
        _raise_if_solver_failed(_cpu(out["flags"]), "LQ_MPC_Controller.solve")
        return {'u_0': _cpu(out["u0"])[0, :, 0].copy(), 'V_N': float(_cpu(out["V"])[0, 0])}


class LQ_MPC_Simulator:
    """utils_class.py:213-285 — closed-loop simulation: the controller plans with (A, B), the plant is
    (A_true, B_true). X and U are instance buffers returned by reference, as in the original (242-243, 285)."""

    def __init__(self, T, N, A, B, Q, R, P, F_u):
        self.T, self.N = int(T), int(N)
        self.A, self.B, self.Q, self.R, self.P, self.F_u = A, B, Q, R, P, F_u
        self.U = np.zeros((np.asarray(B).shape[1], self.T))
        self.X = np.zeros((np.asarray(A).shape[1], self.T + 1))

    def simulate(self, x0, A_true, B_true, x_ref, u_ref):
        A_true = np.asarray(A_true, dtype=np.float64)
        n = A_true.shape[0]
        B_true = np.asarray(B_true, dtype=np.float64).reshape(n, -1)
        eng = _rt.problem_for(A_true, B_true, self.Q, self.R, self.P, self.F_u)
        dA = (np.asarray(self.A, dtype=np.float64) - A_true).reshape(-1, 1)
        dB = (np.asarray(self.B, dtype=np.float64).reshape(n, -1) - B_true).reshape(-1, 1)
        with eng.references(x_ref, u_ref):        # the same reference window at every step (utils_class.py:269)
            out = eng.simulate_batch(dA, dB, self.N, self.T, x0_shared=np.asarray(x0, dtype=np.float64).reshape(n),
                                     want=("J_T", "X", "U", "flags"))
        _raise_if_solver_failed(_cpu(out["flags"]), "LQ_MPC_Simulator.simulate")
        self.X[:, :] = _cpu(out["X"])[:, :, 0].T
        self.U[:, :] = _cpu(out["U"])[:, :, 0].T
        return {'X': self.X, 'U': self.U, 'J_T': float(_cpu(out["J_T"])[0])}


class LQ_RDP_Calculator:
    """utils_class.py:288-406 — coefficients of the performance bound for ONE estimated model (K3, S = 1)."""

    def __init__(self, A, B, Q, R, F_u):
        self.A, self.B, self.Q, self.R, self.F_u = A, B, Q, R, F_u

    def _engine(self):
        return _rt.problem_for(self.A, self.B, self.Q, self.R, None, self.F_u)

    def energy_bound(self, N, e_A, e_B, x, p):
        n = np.asarray(self.A).shape[0]
        m = np.asarray(self.B).reshape(n, -1).shape[1]
        out = self._engine().bounds_batch(None, None, int(N), float(e_A), float(e_B), 0.0,
                                          np.asarray(x, dtype=np.float64).reshape(n), p, 0.0,
                                          K=np.zeros((m, n)), S=1)
        return {'alpha': float(_cpu(out['alpha'])[0]), 'beta': float(_cpu(out['beta'])[0])}

    def energy_decreasing(self, N, e_A, e_B, K, M_V):
        n = np.asarray(self.A).shape[0]
        out = self._engine().bounds_batch(None, None, int(N), float(e_A), float(e_B), float(M_V), np.zeros(n),
                                          (1.0, 1.0, 1.0), 0.0, K=np.atleast_2d(np.asarray(K, dtype=np.float64)), S=1)
        _raise_if_domain_error(_cpu(out['flags'])[0])
        return {'xi': float(_cpu(out['xi'])[0]), 'eta': float(_cpu(out['eta'])[0])}

    def energy_decreasing_extension(self, N, e_A, e_B, K, hatK, M_V):
        """utils_class.py:375-406 (never called by the reference's scripts): engine pieces + scalar composition."""
        from .utils import ex_stability_bounds, ex_stability_lq, fc_ec_h, fc_omega_eta_extension
        eps = local_radius(self.F_u, K, self.Q)
        st = ex_stability_lq(self.A, self.B, self.Q, self.R, K)
        bd = ex_stability_bounds(st['gamma'], eps, M_V)
        oe = fc_omega_eta_extension(N, self.A, self.B, self.Q, self.R, K, hatK, bd['L_V'], bd['N_0'])
        h = fc_ec_h(e_A, e_B, self.Q, self.R)
        return {'xi': h * oe['omega_N1'] + 2 * math.sqrt(h) * oe['omega_N0d5'], 'eta': oe['eta']}


class LQ_RDP_Behavior:
    """utils_class.py:409-682 — single-system curves and surfaces (batched over error levels / horizons / states)."""

    def __init__(self, A, B, Q, R, F_u, K, N_min, N_max, e_pow_min, e_pow_max):
        self.A, self.B, self.Q, self.R, self.F_u, self.K = A, B, Q, R, F_u, K
        self.epsilon = local_radius(F_u, K, Q)
        self.calculator = LQ_RDP_Calculator(A, B, Q, R, F_u)
        self.horizon = np.arange(N_min, N_max + 1)
        self.error_vec = np.array([10 ** i for i in range(e_pow_min, e_pow_max + 1)], dtype=np.float64)

    def _engine(self, A=None, B=None):
        return _rt.problem_for(self.A if A is None else A, self.B if B is None else B, self.Q, self.R, self.Q,
                               self.F_u)

    def OL_energy_bound(self, N, N_points, ext_radius_max, x_ref, u_ref):
        """utils_class.py:439-466."""
        x0_vec = circle_generator(N_points, ext_radius_max, self.epsilon, self.Q)
        eng = self._engine()
        with eng.references(x_ref, u_ref):
            out = eng.mpc_solve_batch(None, None, int(N), pts=x0_vec.T, S=1, want=("M_V", "flags"))
        _raise_if_solver_failed(_cpu(out["flags"]), "OL_energy_bound")
        return float(_cpu(out["M_V"])[0])

    def _bounds_over(self, N, e_vec, K, M_V, x, p):
        """One K3 launch with one 'sample' per error level (dA = dB = 0: the calculator's own model)."""
        n = np.asarray(self.A).shape[0]
        m = np.asarray(self.B).reshape(n, -1).shape[1]
        S = len(e_vec)
        z = np.zeros
        return self._engine().bounds_batch(z((n * n, S)), z((n * m, S)), int(N), np.asarray(e_vec, dtype=np.float64),
                                           np.asarray(e_vec, dtype=np.float64), float(M_V), x, p, 0.0, K=K, S=S)

    def data_generation_xi(self, K, M_V, N_nominal, err_nominal):
        """utils_class.py:468-490."""
        n = np.asarray(self.A).shape[0]
        K = np.atleast_2d(np.asarray(K, dtype=np.float64))
        o = self._bounds_over(N_nominal, self.error_vec, K, M_V, np.zeros(n), (1.0, 1.0, 1.0))
        for f in _cpu(o['flags']):
            _raise_if_domain_error(f)
        xi_error = _cpu(o['xi']).copy()
        xi_horizon = np.zeros(len(self.horizon))
        for i, N in enumerate(self.horizon):
            xi_horizon[i] = self.calculator.energy_decreasing(N, err_nominal['e_A'], err_nominal['e_B'], K, M_V)['xi']
        return {'error': xi_error, 'horizon': xi_horizon}

    def data_generation_alpha_beta(self, x, p, N_nominal, err_nominal):
        """utils_class.py:492-521. The reference's second loop runs `len(self.error_vec)` times over the HORIZON
        array (line 515): kept literally (trailing zeros if there are fewer error levels, IndexError if more)."""
        n = np.asarray(self.A).shape[0]
        m = np.asarray(self.B).reshape(n, -1).shape[1]
        o = self._bounds_over(N_nominal, self.error_vec, np.zeros((m, n)), 0.0, np.asarray(x, dtype=np.float64), p)
        alpha_error, beta_error = _cpu(o['alpha']).copy(), _cpu(o['beta']).copy()
        alpha_horizon = np.zeros(len(self.horizon))
        beta_horizon = np.zeros(len(self.horizon))
        for i in range(len(self.error_vec)):
            r = self.calculator.energy_bound(self.horizon[i], err_nominal['e_A'], err_nominal['e_B'], x, p)
            alpha_horizon[i], beta_horizon[i] = r['alpha'], r['beta']
        return {'error': alpha_error, 'horizon': alpha_horizon}, {'error': beta_error, 'horizon': beta_horizon}

    def _surface(self, N, sim_info, sys_true, err_nominal, info_ref, M_V, p, points):
        """Shared body of data_generation_plane / _mesh for a (n, S) array of initial states: three launches."""
        e_A, e_B = err_nominal['e_A'], err_nominal['e_B']
        A_true = np.asarray(sys_true['A_true'], dtype=np.float64)
        n = A_true.shape[0]
        B_true = np.asarray(sys_true['B_true'], dtype=np.float64).reshape(n, -1)
        S = points.shape[1]
        dec = self.calculator.energy_decreasing(N, e_A, e_B, self.K, M_V)
        factor = 1 / (1 - dec['xi'] - dec['eta'])
        pts = np.ascontiguousarray(points, dtype=np.float64)
        # alpha, beta at every state (the calculator's model)
        m = B_true.shape[1]
        eng = self._engine()
        b = eng.bounds_batch(np.zeros((n * n, S)), np.zeros((n * m, S)), int(N), float(e_A), float(e_B), 0.0, pts, p,
                             0.0, K=np.zeros((m, n)), S=S)
        alpha, beta = _cpu(b['alpha']), _cpu(b['beta'])
        # open-loop expert cost on the TRUE system and closed-loop cost of the estimated-model controller
        eng_t = self._engine(A_true, B_true)
        with eng_t.references(info_ref['x_ref_long'], info_ref['u_ref_long']):            # :591, :672
            V = _cpu(eng_t.mpc_solve_batch(None, None, int(sim_info['N_opc']), x0=pts, S=S, want=("V",))["V"])[0]
        dA = np.repeat((np.asarray(self.A, dtype=np.float64) - A_true).reshape(-1, 1), S, axis=1)
        dB = np.repeat((np.asarray(self.B, dtype=np.float64).reshape(n, -1) - B_true).reshape(-1, 1), S, axis=1)
        with eng_t.references(info_ref['x_ref'], info_ref['u_ref']):                      # :598, :679
            J = _cpu(eng_t.simulate_batch(dA, dB, int(N), int(sim_info['T_mpc']), x0=pts, want=("J_T",))["J_T"])
        return J, factor * (alpha * V + beta), V

    def data_generation_plane(self, N, sim_info, sys_true, err_nominal, info_ref, M_V, p, N_points,
                              ratio_ext_radius):
        """utils_class.py:523-609. Lines 604-605 of the reference append the bound / expert rows to J_MPC_true
        instead of to their own arrays; that behaviour is kept literally."""
        X_1, X_2, J_true, J_bound, V_OPC = [], [], [], [], []
        for j in range(len(ratio_ext_radius)):
            pts = circle_generator(N_points, ratio_ext_radius[j], self.epsilon, self.Q)
            Jt, Jb, V = self._surface(N, sim_info, sys_true, err_nominal, info_ref, M_V, p, pts)
            X_1 = np.append(X_1, pts[0, :])
            X_2 = np.append(X_2, pts[1, :])
            J_true = np.append(J_true, Jt)
            J_bound = np.append(J_true, Jb)
            V_OPC = np.append(J_true, V)
        return {'X': np.vstack((X_1, X_2)), 'J_MPC_true': J_true, 'J_MPC_bound': J_bound, 'V_OPC': V_OPC}

    def data_generation_mesh(self, N, sim_info, sys_true, err_nominal, info_ref, M_V, p, quadrant_range):
        """utils_class.py:611-682 — every mesh node in one batch."""
        X = np.hstack((-quadrant_range['x'], quadrant_range['x']))
        Y = np.hstack((-quadrant_range['y'], quadrant_range['y']))
        coord_X, coord_Y = np.meshgrid(X, Y)
        nx, ny = 2 * len(quadrant_range['x']), 2 * len(quadrant_range['y'])
        pts = np.array([[X[i], Y[j]] for i in range(nx) for j in range(ny)]).T
        Jt, Jb, V = self._surface(N, sim_info, sys_true, err_nominal, info_ref, M_V, p, pts)
        return {'X': coord_X, 'Y': coord_Y, 'J_MPC_true': Jt.reshape(nx, ny), 'J_MPC_bound': Jb.reshape(nx, ny),
                'V_OPC': V.reshape(nx, ny)}


class LQ_RDP_Behavior_Multiple:
    """utils_class.py:685-959 — the batch sweep over sampled model errors (error-level table and horizon table)."""

    def __init__(self, info_opc: dict, info_N: dict, info_e_pow: dict, N_sys: int, norm_type: str,
                 errM_import=True):
        self.A_true, self.B_true = info_opc['A'], info_opc['B']
        self.Q, self.R, self.F_u = info_opc['Q'], info_opc['R'], info_opc['F_u']
        self.N_sys = 5 * N_sys                                   # utils_class.py:726
        self.N_min, self.N_max = info_N['N_min'], info_N['N_max']
        self.N_nominal, self.N_opc, self.N_mpc = info_N['N_nominal'], info_N['N_opc'], info_N['N_mpc']
        self.e_min, self.e_max, self.e_nominal = info_e_pow['e_min'], info_e_pow['e_max'], info_e_pow['e_nominal']
        self.horizon = np.arange(info_N['N_min'], info_N['N_max'] + 1)
        self.error_vec = np.linspace(self.e_min, self.e_max, 10)  # utils_class.py:745
        if errM_import:
            self.error_A = np.load('error_A' + '_' + norm_type + '.npy')    # cwd-relative, as in the reference
            self.error_B = np.load('error_B' + '_' + norm_type + '.npy')
        else:
            d = error_matrix_generator(self.A_true, self.B_true, self.error_vec, N_sys, norm_type)
            self.error_A, self.error_B = d['error_A'], d['error_B']
        self.mpc_open = LQ_MPC_Controller(self.N_opc, self.A_true, self.B_true, self.Q, self.R, self.Q, self.F_u)
        self.engine = _rt.problem_for(self.A_true, self.B_true, self.Q, self.R, self.Q, self.F_u, N_opc=self.N_opc)
        K_lqr = _cpu(self.engine.dlqr_batch(S=1)["K"])[:, 0].reshape(np.asarray(self.B_true).shape[1], -1)
        self.epsilon_lqr = local_radius(self.F_u, -K_lqr, self.Q)             # utils_class.py:761-764

    def _column_block(self, dA, dB, N, e, x0_vec, x_start, V_expert, p, strict_reference=True, refs=(None, None)):
        """All five quantities for S estimated models sharing one horizon N: three launches
        (ring solves -> M_V, closed-loop simulate -> J_T, bounds -> alpha, beta, xi, eta, bound)."""
        eng = self.engine
        with eng.references(*refs):
            ring = eng.mpc_solve_batch(dA, dB, int(N), pts=x0_vec.T, want=("M_V", "flags"))  # utils_class.py:813-824
            mv = ring["M_V"]
            J = eng.simulate_batch(dA, dB, int(N), int(self.N_mpc), x0_shared=x_start, want=("J_T", "flags"))  # 828-833
        b = eng.bounds_batch(dA, dB, int(N), e, e, mv, x_start, p, V_expert, K=None,                       # 840-859
                             strict_reference=strict_reference)
        _raise_if_solver_failed(_cpu(ring["flags"]), "data_generation (M_V ring solves)")
        _raise_if_solver_failed(_cpu(J["flags"]), "data_generation (closed-loop simulation)")
        _raise_if_solver_failed(_cpu(b["flags"]), "data_generation (bounds)")
        return {'alpha': b['alpha'], 'beta': b['beta'], 'xi': b['xi'], 'bound': b['bound'], 'J': J['J_T'],
                'flags': b['flags'] | J['flags'], 'M_V': mv, 'eta': b['eta']}

    def data_generation(self, N_points: int, ext_radius_max: float, info_ref: dict, p: np.ndarray) -> dict:
        """utils_class.py:766-959. Returns the reference's 13 keys and writes data_lq_mpc_multipleSys.npz to cwd."""
        self.engine = _rt.problem_for(self.A_true, self.B_true, self.Q, self.R, self.Q, self.F_u, N_opc=self.N_opc)
        x0_vec = circle_generator(N_points, ext_radius_max, self.epsilon_lqr, self.Q)
        x_start = x0_vec[:, 1].copy()                                                     # :783
        V_expert = self.mpc_open.solve(x_start, info_ref['x_ref_long'], info_ref['u_ref_long'])['V_N']   # :786
        self.engine = _rt.problem_for(self.A_true, self.B_true, self.Q, self.R, self.Q, self.F_u, N_opc=self.N_opc)
        n_err, N_sys = len(self.error_vec), self.N_sys
        eA, eB = self.error_A[:, :, :N_sys, :], self.error_B[:, :, :N_sys, :]
        # ---- error sweep: every (system j, level i) pair in one batch, s = j*n_err + i
        dA, dB = grids_to_soa(np.ascontiguousarray(eA), np.ascontiguousarray(eB))
        e_per = np.tile(self.error_vec, N_sys)
        r = self._column_block(dA, dB, self.N_nominal, e_per, x0_vec, x_start, V_expert, p,
                               refs=(info_ref['x_ref'], info_ref['u_ref']))                # :820, :832
        tab_e = {k: _cpu(r[k]).reshape(N_sys, n_err) for k in ('alpha', 'beta', 'xi', 'bound', 'J')}
        # ---- horizon sweep: level index 4 is hard-coded in the reference (:880), scalar error = e_nominal (:872);
        #      its references are zeros built per horizon (:887-888), whatever info_ref holds
        dA4, dB4 = grids_to_soa(np.ascontiguousarray(eA), np.ascontiguousarray(eB), level=4)
        tab_h = {k: np.zeros([N_sys, len(self.horizon)]) for k in ('alpha', 'beta', 'xi', 'bound', 'J')}
        for i, N in enumerate(self.horizon):
            rh = self._column_block(dA4, dB4, int(N), float(self.e_nominal), x0_vec, x_start, V_expert, p)
            for k in tab_h:
                tab_h[k][:, i] = _cpu(rh[k])
        out_dict = {'error': self.error_vec, 'horizon': self.horizon, 'V_expert': V_expert,
                    'alpha_table_error': tab_e['alpha'], 'beta_table_error': tab_e['beta'],
                    'xi_table_error': tab_e['xi'], 'bound_table_error': tab_e['bound'],
                    'true_cost_error': tab_e['J'],
                    'alpha_table_horizon': tab_h['alpha'], 'beta_table_horizon': tab_h['beta'],
                    'xi_table_horizon': tab_h['xi'], 'bound_table_horizon': tab_h['bound'],
                    'true_cost_horizon': tab_h['J']}
        np.savez('data_lq_mpc_multipleSys.npz', **out_dict)                               # :958
        return out_dict


def __getattr__(name):
    if name in ("Plotter_MPC", "Plotter_PF_LQMPC", "Plotter_PF_LQMPC_Multiple"):
        raise AttributeError(name + " is presentation code (matplotlib) and is out of scope of lq_mpc_b200; "
                             "feed the returned tables to the reference's own plotter")
    raise AttributeError(name)
