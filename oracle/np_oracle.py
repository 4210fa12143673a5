"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy/scipy) of the lq_mpc hot path.

This is the *oracle* the CUDA engine is checked against.  It is an independent restatement of the reference
algorithms (file:line citations into /root/reference), written per sample with numpy/scipy/LAPACK, i.e. with
the same third-party arithmetic the reference itself rests on.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import it; nothing under `lq_mpc_b200/` does.

PARITY PINNING: pinned.  `tests/test_oracle_golden.py` checks this file against
  * the reference's shipped result file `data_lq_mpc_multipleSys.npz` (13 arrays; committed copy in
    tests/golden/), produced by the authors with real cvxpy / python-control / Gurobi, and
  * known answers generated in the build container by importing the untouched reference modules
    (oracle/ref_oracle.py, oracle/make_golden.py -> tests/golden/ref_known_answers.json).
For m > 1 the reference's own `ex_stability_lq` cannot run (`A + B * K` is an element-wise product,
utils.py:356, which raises for n=4,m=2); there this file uses `B @ K`, identical wherever the reference runs.

Third-party arithmetic restated here (all un-pinned in the reference; no lockfile exists):
  cvxpy   -> dense condensed box-QP solved exactly (Cholesky + BVLS active set)      utils_class.py:46-91
  control -> dlqr via scipy.linalg.solve_discrete_are                                utils_class.py:761,840,923
  gurobi  -> vertex enumeration of the input box                                     utils.py:592-650
"""
from __future__ import annotations

import itertools
import math

import numpy as np
import scipy.linalg as sla
from scipy.optimize import lsq_linear

LAMBDA_K = 1.21          # utils.py:364 (hard-coded)
RHO_SHIFT = 0.4          # utils.py:358 (hard-coded)


# ----------------------------------------------------------------------------------------------- input set
def box_from_F(F_u):
    """{u: F_u u <= 1} -> (lo, hi); only single-non-zero rows (the only kind built by the reference:
    working_example_multiple.py:25, working_example_single.py:27)."""
    F_u = np.atleast_2d(np.asarray(F_u, dtype=float))
    m = F_u.shape[1]
    lo, hi = np.full(m, -np.inf), np.full(m, np.inf)
    for row in F_u:
        nz = np.flatnonzero(row)
        if len(nz) != 1:
            raise NotImplementedError("box-shaped F_u only")
        j = nz[0]
        if row[j] > 0:
            hi[j] = min(hi[j], 1.0 / row[j])
        else:
            lo[j] = max(lo[j], 1.0 / row[j])
    return lo, hi


def bar_u(lo, hi):
    """max ||u||^2 on the box (utils.py:592-619)."""
    return float(np.sum(np.maximum(lo * lo, hi * hi)))


def bar_d_u(lo, hi):
    """max ||u1-u2||^2 on box x box (utils.py:622-650)."""
    return float(np.sum((hi - lo) ** 2))


# ----------------------------------------------------------------------------------------------- a1: MPC solve
def riccati(A, B, Q, R, P_term, N):
    """Finite-horizon Riccati recursion for the cost of utils_class.py:62-75 with zero references:
    V_N(x0) = sum_{k<N}(x_k'Q x_k + u_k'R u_k) + x_N'P x_N (V_N adds x0'Qx0 at utils_class.py:91).
    Returns gains K[0..N-1] (u_k = K_k x_k) and cost-to-go matrices P[0..N] (P[N] = P_term)."""
    P = [None] * (N + 1)
    K = [None] * N
    P[N] = np.array(P_term, dtype=float)
    for k in range(N - 1, -1, -1):
        G = R + B.T @ P[k + 1] @ B
        H = B.T @ P[k + 1] @ A
        K[k] = -np.linalg.solve(G, H)
        P[k] = Q + A.T @ P[k + 1] @ A + H.T @ K[k]
        P[k] = 0.5 * (P[k] + P[k].T)
    return K, P


def sl_syn_Phi(N, A):
    """[I; A; ...; A^N]  (utils.py:126-142)."""
    n = A.shape[0]
    out = np.empty(((N + 1) * n, n))
    M = np.eye(n)
    for k in range(N + 1):
        out[k * n:(k + 1) * n] = M
        M = A @ M
    return out


def sl_syn_Gamma(N, A, B):
    """Block lower-triangular Toeplitz map from the stacked inputs to the stacked states x_0..x_N
    (utils.py:145-174); block (i, j) = A^(i-1-j) B for i > j."""
    n, m = B.shape
    out = np.zeros(((N + 1) * n, N * m))
    AB = B.copy()
    for d in range(N):                  # d = i-1-j
        for j in range(N - d):
            i = j + d + 1
            out[i * n:(i + 1) * n, j * m:(j + 1) * m] = AB
        AB = A @ AB
    return out


def condensed_qp(N, A, B, Q, R, P_term, x0, x_ref=None, u_ref=None):
    """H, g, c0 with objective z'Hz + 2g'z + c0 for z = vec(u_0..u_{N-1}) (time-major)
    (utils_class.py:59-75: stage cost Q on x_1..x_{N-1}, P on x_N, R on every u_k; x_ref[:, i] is the reference of
    x_{i+1}, u_ref[:, i] of u_i — only the first N columns are read)."""
    n, m = B.shape
    Gam = sl_syn_Gamma(N, A, B)[n:]            # rows for x_1..x_N
    Phi = sl_syn_Phi(N, A)[n:]
    W = sla.block_diag(*([Q] * (N - 1) + [P_term]))
    Rb = sla.block_diag(*([R] * N))
    H = Rb + Gam.T @ W @ Gam
    f = Phi @ x0
    if x_ref is not None:
        f = f - np.asarray(x_ref, dtype=float)[:, :N].T.reshape(-1)
    g = Gam.T @ W @ f
    c0 = f @ W @ f
    if u_ref is not None:
        w = np.asarray(u_ref, dtype=float)[:, :N].T.reshape(-1)
        g = g - Rb @ w
        c0 = c0 + w @ Rb @ w
    return 0.5 * (H + H.T), g, float(c0)


def box_qp(H, g, lo, hi):
    """Exact minimiser of z'Hz + 2g'z on a box: Cholesky -> bounded least squares (BVLS active set)."""
    L = np.linalg.cholesky(H)
    rhs = -sla.solve_triangular(L, g, lower=True)
    nz = len(g)
    if np.all(np.isinf(lo)) and np.all(np.isinf(hi)):
        return sla.solve_triangular(L.T, rhs, lower=False)
    res = lsq_linear(L.T, rhs, bounds=(lo, hi), method="bvls", tol=1e-15, max_iter=10 * nz + 50)
    return res.x


def mpc_solve(N, A, B, Q, R, P_term, lo, hi, x0, exact_fast=True, _ric=None, x_ref=None, u_ref=None, F_u=None):
    """LQ_MPC_Controller.solve (utils_class.py:48-91). Returns (u_0, V_N, active). Non-zero references always take
    the dense condensed QP (the law is affine then). F_u (p, m): general input polytope F_u u_k <= 1
    (utils_class.py:81) instead of the box lo / hi — dense active-set QP of oracle/qp_dense.py.

    exact_fast: if the unconstrained (Riccati) open-loop plan is feasible it IS the QP minimiser (KKT with zero
    multipliers), so the dense QP is only assembled when some planned input leaves the box."""
    n, m = B.shape
    x0 = np.asarray(x0, dtype=float)
    lo = np.full(m, -np.inf) if lo is None else np.asarray(lo, dtype=float)
    hi = np.full(m, np.inf) if hi is None else np.asarray(hi, dtype=float)
    tracking = (x_ref is not None and np.any(np.asarray(x_ref) != 0)) or \
        (u_ref is not None and np.any(np.asarray(u_ref) != 0))
    if exact_fast and not tracking and F_u is None:
        K, P = riccati(A, B, Q, R, P_term, N) if _ric is None else _ric
        x = x0
        feas = True
        for k in range(N):
            u = K[k] @ x
            if np.any(u < lo) or np.any(u > hi):
                feas = False
                break
            x = A @ x + B @ u
        if feas:
            return K[0] @ x0, float(x0 @ P[0] @ x0), False
    H, g, c0 = condensed_qp(N, A, B, Q, R, P_term, x0, x_ref, u_ref)
    if F_u is not None:
        from .qp_dense import ineq_qp
        F_u = np.atleast_2d(np.asarray(F_u, dtype=float))
        z, Wact = ineq_qp(H, g, np.kron(np.eye(N), F_u), np.ones(N * F_u.shape[0]))
        V = float(z @ H @ z + 2 * g @ z + c0 + x0 @ Q @ x0)
        return z[:m].copy(), V, len(Wact) > 0
    z = box_qp(H, g, np.tile(lo, N), np.tile(hi, N))
    V = float(z @ H @ z + 2 * g @ z + c0 + x0 @ Q @ x0)
    active = bool(np.any(z <= np.tile(lo, N)) or np.any(z >= np.tile(hi, N))) if tracking else True
    return z[:m].copy(), V, active


# ----------------------------------------------------------------------------------------------- a2: simulator
def simulate(T, N, A, B, Q, R, P_term, lo, hi, x0, A_true, B_true, exact_fast=True, x_ref=None, u_ref=None,
             F_u=None):
    """LQ_MPC_Simulator.simulate (utils_class.py:245-285): controller plans with (A,B), plant is (A_true,B_true).
    J_T = x0'Qx0 + sum_t (x_{t+1}'Q x_{t+1} + u_t'R u_t)."""
    n, m = B.shape
    X = np.zeros((n, T + 1))
    U = np.zeros((m, T))
    X[:, 0] = x0
    cost = float(X[:, 0] @ Q @ X[:, 0])
    n_active = 0
    ric = riccati(A, B, Q, R, P_term, N) if exact_fast else None
    if F_u is not None:
        exact_fast = False
    for t in range(T):
        u, _, act = mpc_solve(N, A, B, Q, R, P_term, lo, hi, X[:, t], exact_fast, ric, x_ref, u_ref, F_u)
        n_active += int(act)
        U[:, t] = u
        xn = A_true @ X[:, t] + B_true @ u
        X[:, t + 1] = xn
        cost += float(xn @ Q @ xn + u @ R @ u)
    return {'X': X, 'U': U, 'J_T': cost, 'n_active': n_active}


def spectral_radius(M):
    return float(np.max(np.abs(np.linalg.eigvals(M))))


def closed_loop_inf_cost(A_true, B_true, K, Q, R, x0):
    """J_inf = sum_t x_t'(Q + K'RK) x_t for x+ = (A_true + B_true K) x: the T->inf limit of the J_T of
    utils_class.py:261,282-283 under a fixed linear law. Returns (J_inf, rho); J_inf = +inf when rho >= 1."""
    Acl = A_true + B_true @ K
    rho = spectral_radius(Acl)
    if not rho < 1.0:
        return math.inf, rho
    W = Q + K.T @ R @ K
    S = sla.solve_discrete_lyapunov(Acl.T, W)
    return float(x0 @ S @ x0), rho


# ----------------------------------------------------------------------------------------------- a3: dlqr
def dlqr(A, B, Q, R):
    """control.dlqr convention (u = -Kx): K = (R+B'PB)^-1 B'PA with P the stabilising DARE solution."""
    P = sla.solve_discrete_are(A, B, Q, R)
    K = np.linalg.solve(R + B.T @ P @ B, B.T @ P @ A)
    return K, P


# ----------------------------------------------------------------------------------------------- a4/a5: scalars
def my_eigen(M):
    """utils.py:52-68 (general eigvals, then max/min)."""
    ev = np.linalg.eigvals(np.atleast_2d(M))
    if np.all(np.isreal(ev)):
        ev = ev.real
    mx, mn = np.max(ev), np.min(ev)
    return {'max': mx, 'min': mn, 'ratio': mx / mn}


def g_x(p, i, e_A, f_A):
    """utils.py:78-95."""
    return ((e_A + f_A) ** i - f_A ** i) ** p


def g_u(p, i, e_A, f_A, e_B, f_B):
    """utils.py:98-117."""
    return ((e_B + f_B) * g_x(1, i, e_A, f_A) + e_B * f_A ** i) ** p


def bar_g_x(N, e_A, f_A):
    """utils.py:186-201."""
    return sum(g_x(1, i + 1, e_A, f_A) for i in range(N))


def bar_g_u(N, e_A, f_A, e_B, f_B):
    """utils.py:204-223 (double partial sum)."""
    s_in = s_out = 0.0
    for i in range(N):
        s_in += g_u(1, i, e_A, f_A, e_B, f_B)
        s_out += s_in
    return s_out


def norm2(M):
    return float(np.linalg.norm(np.atleast_2d(M), ord=2))


def hat_H(N, A, B, Q, R, strict_reference=True):
    """utils.py:316-319. strict_reference keeps the literal kron(Q, I_{N+1}) / kron(R, I_N) ordering, which only
    coincides with the time-major layout of Gamma when Q and R are multiples of the identity."""
    Gam = sl_syn_Gamma(N, A, B)
    if strict_reference:
        bQ, bR = np.kron(Q, np.eye(N + 1)), np.kron(R, np.eye(N))
    else:
        bQ, bR = np.kron(np.eye(N + 1), Q), np.kron(np.eye(N), R)
    return bR + Gam.T @ bQ @ Gam


def fc_ec_theta(N, e_A, e_B, A, B, maxQ):
    """utils.py:226-264."""
    f_A, f_B = norm2(A), norm2(B)
    nG, nP = norm2(sl_syn_Gamma(N, A, B)), norm2(sl_syn_Phi(N, A))
    bx, bu = bar_g_x(N, e_A, f_A), bar_g_u(N, e_A, f_A, e_B, f_B)
    return {'theta_u': maxQ * (2 * nG * bu + bu ** 2), 'theta_x_u': maxQ * (nG * bx + nP * bu + bx * bu)}


def fc_ec_E(N, e_A, e_B, A, B, Q, R, x, bu_max, bdu_max, strict_reference=True):
    """utils.py:267-334."""
    iQ, iR = my_eigen(Q), my_eigen(R)
    f_A, f_B = norm2(A), norm2(B)
    nx = float(np.linalg.norm(x))
    s_in = s_out = 0.0
    for i in range(N + 1):
        s_out += (s_in + g_x(2, i, e_A, f_A)) * (nx ** 2 + i * bu_max)
        s_in += g_u(2, i, e_A, f_A, e_B, f_B)
    E_psi = iQ['max'] * s_out
    th = fc_ec_theta(N, e_A, e_B, A, B, iQ['max'])
    bar_theta = math.sqrt(N * bu_max) * th['theta_u'] + nx * th['theta_x_u']
    H = hat_H(N, A, B, Q, R, strict_reference)
    min_H = np.min(np.linalg.eigvals(H))
    min_H = float(np.real(min_H))
    E_u = iR['max'] * min(math.sqrt(N * bdu_max), bar_theta / min_H) ** 2
    nG = norm2(sl_syn_Gamma(N, A, B))
    E_psi_u = iQ['max'] / iR['max'] * (nG + bar_g_u(N, e_A, f_A, e_B, f_B)) ** 2 * E_u
    return {'E_psi': float(E_psi), 'E_u': float(E_u), 'E_psi_u': float(E_psi_u), 'min_H': min_H,
            'norm_Gamma': nG, 'theta_u': th['theta_u'], 'theta_x_u': th['theta_x_u']}


def bar_u_poly(F_u):
    """utils.py:592-650 for a general polytope: the maxima of ||u||^2 and ||u1 - u2||^2 sit at vertices."""
    from .qp_dense import polytope_vertices
    V = polytope_vertices(F_u)
    D = V[:, None, :] - V[None, :, :]
    return float(np.max(np.sum(V * V, axis=1))), float(np.max(np.sum(D * D, axis=2)))


def energy_bound(A, B, Q, R, lo, hi, N, e_A, e_B, x, p, strict_reference=True, F_u=None):
    """LQ_RDP_Calculator.energy_bound (utils_class.py:308-342). F_u: general polytope instead of the box lo / hi."""
    bu, bdu = (bar_u(lo, hi), bar_d_u(lo, hi)) if F_u is None else bar_u_poly(F_u)
    E = fc_ec_E(N, e_A, e_B, A, B, Q, R, x, bu, bdu, strict_reference)
    p = np.asarray(p, dtype=float)
    q = 1.0 / p
    sp, su, spu = math.sqrt(E['E_psi']), math.sqrt(E['E_u']), math.sqrt(E['E_psi_u'])
    alpha = max(p[0] * sp + p[2] * spu + p[0] * sp * p[2] * spu, p[1] * su)
    beta = (1 + p[0] * sp) * (q[2] * spu + E['E_psi_u']) + q[1] * su + E['E_u'] + q[0] * sp + E['E_psi']
    return {'alpha': float(alpha), 'beta': float(beta), **E}


def local_radius(lo, hi, K, Q):
    """utils.py:548-564 for the box rows F_u = [diag(1/hi); diag(1/lo)]."""
    K = np.atleast_2d(K)
    Qi = np.linalg.inv(Q)
    a = []
    for j in range(K.shape[0]):
        kq = float(K[j] @ Qi @ K[j])
        for b in (lo[j], hi[j]):
            if np.isfinite(b):
                a.append(kq / (b * b))
    with np.errstate(divide='ignore'):
        return float(np.float64(1.0) / np.float64(max(a)))      # numpy semantics: 1/0 = inf (utils.py:564)


def ex_stability_lq(A, B, Q, R, K):
    """utils.py:343-380, with B @ K for the closed loop (see module docstring)."""
    K = np.atleast_2d(K)
    nK = norm2(K)
    rho_K = (spectral_radius(A + B @ K) + RHO_SHIFT) ** 2
    iQ, iR = my_eigen(Q), my_eigen(R)
    C = (1 + iR['max'] * nK ** 2 / iQ['min']) * max(1.0, iQ['ratio'] * LAMBDA_K)
    gamma = C / (1 - rho_K)
    return {'C_K': float(C), 'lambda_K': LAMBDA_K, 'rho_K': float(rho_K), 'gamma': float(gamma),
            'rho_gamma': float((gamma - 1) / gamma)}


def ex_stability_bounds(gamma, eps_K, M_V):
    """utils.py:567-584."""
    return {'L_V': max(gamma, M_V / eps_K), 'N_0': math.ceil(max(0.0, M_V / eps_K - gamma))}


def geo_M(M, k):
    """utils.py:393-409."""
    f = norm2(M)
    return float(k) if f == 1 else (1 - f ** (2 * k)) / (1 - f ** 2)


def fc_omega_eta(N, A, B, Q, R, K, L_V, N_0):
    """utils.py:469-523."""
    f_A = norm2(A)
    iQ = my_eigen(Q)
    G_A = geo_M(A, N - 1)
    st = ex_stability_lq(A, B, Q, R, K)
    term = 1 + f_A ** 2 * iQ['ratio']
    N_min = math.ceil(N_0 - math.log(f_A ** 2 * iQ['ratio'] * st['gamma']) / math.log(st['rho_gamma']))
    w1 = iQ['max'] * (term * f_A ** (2 * N - 2) + G_A)
    decay = iQ['max'] * f_A ** (2 * N - 2) * st['gamma'] * st['rho_gamma'] ** (N - N_0)
    w05 = math.sqrt(iQ['max'] * (L_V - 1) * G_A) + 0.5 * term * math.sqrt(decay)
    eta = (term - 1) * st['gamma'] * st['rho_gamma'] ** (N - N_0)
    err_th = ((math.sqrt(w05 ** 2 + w1 * (1 - eta)) - w05) / w1) ** 2
    return {'omega_N1': float(w1), 'omega_N0d5': float(w05), 'eta': float(eta), 'err_th': float(err_th),
            'N_min': N_min}


def fc_ec_h(e_A, e_B, Q, R):
    """utils.py:526-538."""
    return e_A ** 2 / my_eigen(Q)['min'] + e_B ** 2 / my_eigen(R)['min']


def energy_decreasing(A, B, Q, R, lo, hi, N, e_A, e_B, K, M_V, F_u=None):
    """LQ_RDP_Calculator.energy_decreasing (utils_class.py:344-373). K in the u = +Kx convention. F_u: general
    polytope rows for local_radius (utils.py:548-564) instead of the box lo / hi."""
    if F_u is None:
        eps = local_radius(lo, hi, K, Q)
    else:
        Mx = np.atleast_2d(F_u) @ np.atleast_2d(K)
        Qi = np.linalg.inv(Q)
        eps = 1.0 / max(float(r @ Qi @ r) for r in Mx)
    st = ex_stability_lq(A, B, Q, R, K)
    bd = ex_stability_bounds(st['gamma'], eps, M_V)
    oe = fc_omega_eta(N, A, B, Q, R, K, bd['L_V'], bd['N_0'])
    h = fc_ec_h(e_A, e_B, Q, R)
    xi = h * oe['omega_N1'] + 2 * math.sqrt(h) * oe['omega_N0d5']
    return {'xi': float(xi), 'eta': oe['eta'], 'epsilon_K': eps, 'h': float(h), **st, **bd,
            'omega_N1': oe['omega_N1'], 'omega_N0d5': oe['omega_N0d5'], 'err_th': oe['err_th'],
            'N_min': oe['N_min']}


def fc_omega_eta_extension(N, A, B, Q, R, K, hatK, L_V, N_0):
    """utils.py:412-466 — the variant with a second gain hatK: the terminal-cost propagation term becomes
    C_K(hatK) + ||A + B K||_2^2 cond(Q) and N_min is NOT rounded up (no math.ceil at utils.py:446)."""
    K, hatK = np.atleast_2d(K), np.atleast_2d(hatK)
    f_A = norm2(A)
    f_cl = norm2(A + B @ K)
    iQ = my_eigen(Q)
    G_A = geo_M(A, N - 1)
    st = ex_stability_lq(A, B, Q, R, K)
    st_dev = ex_stability_lq(A, B, Q, R, hatK)
    term = st_dev['C_K'] + f_cl ** 2 * iQ['ratio']
    N_min = N_0 - math.log((term - 1) * st['gamma']) / math.log(st['rho_gamma'])
    w1 = iQ['max'] * (term * f_A ** (2 * N - 2) + G_A)
    decay = iQ['max'] * f_A ** (2 * N - 2) * st['gamma'] * st['rho_gamma'] ** (N - N_0)
    w05 = math.sqrt(iQ['max'] * (L_V - 1) * G_A) + 0.5 * term * math.sqrt(decay)
    eta = (term - 1) * st['gamma'] * st['rho_gamma'] ** (N - N_0)
    err_th = ((math.sqrt(w05 ** 2 + w1 * (1 - eta)) - w05) / w1) ** 2
    return {'omega_N1': float(w1), 'omega_N0d5': float(w05), 'eta': float(eta), 'err_th': float(err_th),
            'N_min': float(N_min)}


def energy_decreasing_extension(A, B, Q, R, lo, hi, N, e_A, e_B, K, hatK, M_V):
    """LQ_RDP_Calculator.energy_decreasing_extension (utils_class.py:375-406). K, hatK in the u = +Kx convention."""
    eps = local_radius(lo, hi, K, Q)
    st = ex_stability_lq(A, B, Q, R, K)
    bd = ex_stability_bounds(st['gamma'], eps, M_V)
    oe = fc_omega_eta_extension(N, A, B, Q, R, K, hatK, bd['L_V'], bd['N_0'])
    h = fc_ec_h(e_A, e_B, Q, R)
    xi = h * oe['omega_N1'] + 2 * math.sqrt(h) * oe['omega_N0d5']
    return {'xi': float(xi), 'eta': oe['eta'], **oe}


# ----------------------------------------------------------------------------------------------- a8: x0 ring
def circle_generator(N_points, ratio, base, Q):
    """utils.py:683-704. scipy's cho_factor leaves the strict lower triangle of its return value untouched (it
    holds Q's entries) and the reference inverts that whole array; this is reproduced literally. For diagonal Q
    (every shipped scenario) it is the plain upper Cholesky factor."""
    Q = np.asarray(Q, dtype=float)
    U = np.linalg.cholesky(Q).T
    root = np.triu(U) + np.tril(Q, -1)
    th = np.linspace(0, 2 * (1 - 1 / N_points) * math.pi, N_points)
    r = ratio * math.sqrt(base)
    pts = np.vstack([r * np.cos(th), r * np.sin(th)])
    return np.linalg.inv(root) @ pts


# ----------------------------------------------------------------------------------------------- a10: statistics
def column_stats(table):
    """The four per-column statistics the plotters draw (utils.py:895-898). np.std is the population std."""
    t = np.asarray(table, dtype=float)
    return {'max': t.max(axis=0), 'min': t.min(axis=0), 'mean': t.mean(axis=0), 'std': t.std(axis=0)}


# ----------------------------------------------------------------------------------------------- a6/a7: sweep
def M_V_of(N, A, B, Q, R, lo, hi, x0_vec):
    """OL_energy_bound (utils_class.py:439-466; inline copies 813-824, 896-907)."""
    ric = riccati(A, B, Q, R, Q, N)
    return max(mpc_solve(N, A, B, Q, R, Q, lo, hi, x0_vec[:, k], True, ric)[1] for k in range(x0_vec.shape[1]))


def eval_one(A, B, A_true, B_true, Q, R, lo, hi, N, T, e, x0_vec, x_start, V_expert, p,
             strict_reference=True):
    """Body of the two hot loops of data_generation (utils_class.py:806-859 / 889-942) for one (sample, column)."""
    M_V = M_V_of(N, A, B, Q, R, lo, hi, x0_vec)
    J = simulate(T, N, A, B, Q, R, Q, lo, hi, x_start, A_true, B_true)['J_T']
    K, _ = dlqr(A, B, Q, R)
    dec = energy_decreasing(A, B, Q, R, lo, hi, N, e, e, -K, M_V)
    bnd = energy_bound(A, B, Q, R, lo, hi, N, e, e, x_start, p, strict_reference)
    bound = (bnd['alpha'] * V_expert + bnd['beta']) / (1 - dec['xi'] - dec['eta'])
    return {'M_V': M_V, 'J': J, 'xi': dec['xi'], 'eta': dec['eta'], 'alpha': bnd['alpha'], 'beta': bnd['beta'],
            'bound': bound}


def data_generation(A_true, B_true, Q, R, F_u, error_A, error_B, info_N, info_e, N_points, ext_radius_max, p,
                    strict_reference=True, max_sys=None):
    """LQ_RDP_Behavior_Multiple.__init__ + data_generation (utils_class.py:691-959), zero references.
    error_A: (n,n,N_sys,n_err), error_B: (n,m,N_sys,n_err). max_sys truncates the sample axis (tests)."""
    lo, hi = box_from_F(F_u)
    N_sys = error_A.shape[2] if max_sys is None else min(max_sys, error_A.shape[2])
    error_vec = np.linspace(info_e['e_min'], info_e['e_max'], 10)             # utils_class.py:745
    horizon = np.arange(info_N['N_min'], info_N['N_max'] + 1)                  # utils_class.py:741
    K_lqr, _ = dlqr(A_true, B_true, Q, R)                                      # 761
    eps_lqr = local_radius(lo, hi, -K_lqr, Q)                                  # 764
    x0_vec = circle_generator(N_points, ext_radius_max, eps_lqr, Q)            # 782
    x_start = x0_vec[:, 1]                                                     # 783
    V_expert = mpc_solve(info_N['N_opc'], A_true, B_true, Q, R, Q, lo, hi, x_start)[1]   # 786
    names = ('alpha', 'beta', 'xi', 'bound', 'J')
    tab_e = {k: np.zeros((N_sys, len(error_vec))) for k in names}
    tab_h = {k: np.zeros((N_sys, len(horizon))) for k in names}
    for i, e in enumerate(error_vec):                                           # 802
        for j in range(N_sys):                                                  # 806
            r = eval_one(A_true + error_A[:, :, j, i], B_true + error_B[:, :, j, i], A_true, B_true, Q, R, lo, hi,
                         info_N['N_nominal'], info_N['N_mpc'], e, x0_vec, x_start, V_expert, p, strict_reference)
            for k in names:
                tab_e[k][j, i] = r[k]
    e_h = info_e['e_nominal']                                                   # 872
    idx = 4                                                                     # 880 (hard-coded level index)
    for i, N in enumerate(horizon):                                             # 886
        for j in range(N_sys):
            r = eval_one(A_true + error_A[:, :, j, idx], B_true + error_B[:, :, j, idx], A_true, B_true, Q, R, lo,
                         hi, int(N), info_N['N_mpc'], e_h, x0_vec, x_start, V_expert, p, strict_reference)
            for k in names:
                tab_h[k][j, i] = r[k]
    return {'error': error_vec, 'horizon': horizon, 'V_expert': V_expert,
            'alpha_table_error': tab_e['alpha'], 'beta_table_error': tab_e['beta'], 'xi_table_error': tab_e['xi'],
            'bound_table_error': tab_e['bound'], 'true_cost_error': tab_e['J'],
            'alpha_table_horizon': tab_h['alpha'], 'beta_table_horizon': tab_h['beta'],
            'xi_table_horizon': tab_h['xi'], 'bound_table_horizon': tab_h['bound'],
            'true_cost_horizon': tab_h['J'], 'x0_vec': x0_vec, 'epsilon_lqr': eps_lqr}
