"""Drop-in for the reference module `utils` (same function names, argument meaning and returned dict keys;
citations are into /root/reference/utils.py).

Everything that needs linear algebra on the model — matrix 2-norms, spectral radii, extreme eigenvalues, the DARE,
the (N m) x (N m) Gram/Hessian spectra — is computed by the CUDA engine (K3 with S = 1); what remains on the host is
the reference's scalar bookkeeping (powers, sums, ceil/log of Python floats), index arithmetic, sampling and matrix
*builders* the kernels never need. There is no CPU fallback for the engine-backed functions.
"""
from __future__ import annotations

import bisect
import math
import random

import numpy as np

from . import runtime as _rt

# --------------------------------------------------------------------------------------------- engine-backed helpers


def _detail(A, B, Q, R, K=None, N=1, e_A=0.0, e_B=0.0, M_V=0.0, x=None, p=(1.0, 1.0, 1.0), F_u=None,
            bar_u=-1.0, bar_d_u=-1.0):
    A = np.atleast_2d(np.asarray(A, dtype=np.float64))
    n = A.shape[0]
    B = np.asarray(B, dtype=np.float64).reshape(n, -1)
    eng = _rt.problem_for(A, B, Q, R, None, F_u, scratch=True)
    x = np.zeros(n) if x is None else np.asarray(x, dtype=np.float64).reshape(n)
    out = eng.bounds_batch(None, None, int(N), float(e_A), float(e_B), float(M_V), x, p, 0.0, K=K, S=1,
                           bar_u=bar_u, bar_d_u=bar_d_u)
    res = {k: float(v.cpu().numpy()[0]) for k, v in out.items() if k != "flags"}
    res["flags"] = int(out["flags"].cpu().numpy()[0])
    return res


def my_eigen(M):
    """utils.py:52-68 — max / min eigenvalue and their ratio. Engine-backed for symmetric M (the reference only ever
    passes the weights Q and R)."""
    M = np.atleast_2d(np.asarray(M, dtype=np.float64))
    if not np.array_equal(M, M.T):
        raise NotImplementedError("my_eigen: only symmetric matrices (Q, R) are supported by the engine")
    n = M.shape[0]
    eng = _rt.problem_for(np.eye(n), np.eye(n, 1), M, np.eye(1), None, None, scratch=True)
    pr = eng.prepared()
    return {'max': pr['maxQ'], 'min': pr['minQ'], 'ratio': pr['maxQ'] / pr['minQ']}


def fc_ec_g_x(n, i, e_A, f_A):
    """utils.py:78-95."""
    return ((e_A + f_A) ** i - f_A ** i) ** n


def fc_ec_g_u(n, i, e_A, f_A, e_B, f_B):
    """utils.py:98-117."""
    return ((e_B + f_B) * fc_ec_g_x(1, i, e_A, f_A) + e_B * (f_A ** i)) ** n


def sl_syn_Phi(N, A):
    """utils.py:126-142: [I; A; ...; A^N] (host builder; the kernels work on Gram matrices instead)."""
    A = np.asarray(A, dtype=np.float64)
    blocks = [np.eye(A.shape[0])]
    for _ in range(N):
        blocks.append(A @ blocks[-1])
    return np.vstack(blocks)


def sl_syn_Gamma(N, A, B):
    """utils.py:145-174 (host builder)."""
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    n, m = B.shape
    out = np.zeros(((N + 1) * n, N * m))
    blk = B.copy()
    for d in range(N):
        for j in range(N - d):
            out[(j + d + 1) * n:(j + d + 2) * n, j * m:(j + 1) * m] = blk
        blk = A @ blk
    return out


def fc_ec_bar_g_x(N, e_A, f_A):
    """utils.py:186-201."""
    return sum(fc_ec_g_x(1, i + 1, e_A, f_A) for i in range(N))


def fc_ec_bar_g_u(N, e_A, f_A, e_B, f_B):
    """utils.py:204-223."""
    s_in = s_out = 0
    for i in range(N):
        s_in += fc_ec_g_u(1, i, e_A, f_A, e_B, f_B)
        s_out += s_in
    return s_out


def fc_ec_theta(N, e_A, e_B, A, B, maxQ):
    """utils.py:226-264 — ||Gamma||_2, ||Phi||_2, ||A||_2, ||B||_2 from the engine."""
    n = np.atleast_2d(A).shape[0]
    m = np.asarray(B).reshape(n, -1).shape[1]
    d = _detail(A, B, np.eye(n), np.eye(m), K=np.zeros((m, n)), N=N, e_A=e_A, e_B=e_B)
    bx = fc_ec_bar_g_x(N, e_A, d['norm_A'])
    bu = fc_ec_bar_g_u(N, e_A, d['norm_A'], e_B, d['norm_B'])
    return {'theta_u': maxQ * (2 * d['norm_Gamma'] * bu + bu ** 2),
            'theta_x_u': maxQ * (d['norm_Gamma'] * bx + d['norm_Phi'] * bu + bx * bu)}


def fc_ec_E(N, e_A, e_B, A, B, Q, R, x, bar_u, bar_d_u):
    """utils.py:267-334 — evaluated entirely by K3."""
    n = np.atleast_2d(A).shape[0]
    m = np.asarray(B).reshape(n, -1).shape[1]
    d = _detail(A, B, Q, R, K=np.zeros((m, n)), N=N, e_A=e_A, e_B=e_B, x=x, bar_u=float(bar_u),
                bar_d_u=float(bar_d_u))
    return {'E_psi': d['E_psi'], 'E_u': d['E_u'], 'E_psi_u': d['E_psi_u']}


def ex_stability_lq(A, B, Q, R, K):
    """utils.py:343-380 — ||K||_2 and rho(A + B K) from the engine (see DESIGN.md on `A + B * K`)."""
    d = _detail(A, B, Q, R, K=np.atleast_2d(K))
    return {'C_K': d['C_K'], 'lambda_K': 1.21, 'rho_K': d['rho_K'], 'gamma': d['gamma'],
            'rho_gamma': d['rho_gamma']}


def geo_M(M, n):
    """utils.py:393-409."""
    M = np.atleast_2d(np.asarray(M, dtype=np.float64))
    k = M.shape[0]
    f = _detail(M, np.eye(k, 1), np.eye(k), np.eye(1), K=np.zeros((1, k)))['norm_A']
    return n if f == 1 else (1 - f ** (2 * n)) / (1 - f ** 2)


def _omega_eta_core(N, normA, iQ, st, my_term, L_V, N_0, G_A, N_min):
    w1 = iQ['max'] * (my_term * (normA ** (2 * N - 2)) + G_A)
    decay = iQ['max'] * (normA ** (2 * N - 2)) * st['gamma'] * (st['rho_gamma'] ** (N - N_0))
    w05 = math.sqrt(iQ['max'] * (L_V - 1) * G_A) + 0.5 * my_term * math.sqrt(decay)
    eta = (my_term - 1) * st['gamma'] * (st['rho_gamma'] ** (N - N_0))
    err_th = ((math.sqrt(w05 ** 2 + w1 * (1 - eta)) - w05) / w1) ** 2
    return {'omega_N1': w1, 'omega_N0d5': w05, 'eta': eta, 'err_th': err_th, 'N_min': N_min}


def fc_omega_eta(N, A, B, Q, R, K, L_V, N_0):
    """utils.py:469-523 — engine supplies ||A||_2, gamma, rho_gamma, eigen extremes; scalar composition as in
    the reference (so a non-positive log argument raises ValueError exactly as there)."""
    d = _detail(A, B, Q, R, K=np.atleast_2d(K))
    iQ = my_eigen(Q)
    normA = d['norm_A']
    st = {'gamma': d['gamma'], 'rho_gamma': d['rho_gamma']}
    G_A = (N - 1) if normA == 1 else (1 - normA ** (2 * (N - 1))) / (1 - normA ** 2)
    my_term = 1 + (normA ** 2) * iQ['ratio']
    N_min = math.ceil((N_0 - math.log((normA ** 2) * iQ['ratio'] * st['gamma']) / math.log(st['rho_gamma'])))
    return _omega_eta_core(N, normA, iQ, st, my_term, L_V, N_0, G_A, N_min)


def fc_omega_eta_extension(N, A, B, Q, R, K, hatK, L_V, N_0):
    """utils.py:412-466 (never called by the reference's scripts; host composition of engine-computed pieces)."""
    A = np.atleast_2d(np.asarray(A, dtype=np.float64))
    n = A.shape[0]
    B = np.asarray(B, dtype=np.float64).reshape(n, -1)
    K = np.atleast_2d(K)
    d = _detail(A, B, Q, R, K=K)
    dh = _detail(A, B, Q, R, K=np.atleast_2d(hatK))
    normA_cl = _detail(A + B @ K, B, Q, R, K=np.zeros_like(K))['norm_A']
    iQ = my_eigen(Q)
    normA = d['norm_A']
    st = {'gamma': d['gamma'], 'rho_gamma': d['rho_gamma']}
    G_A = (N - 1) if normA == 1 else (1 - normA ** (2 * (N - 1))) / (1 - normA ** 2)
    my_term = dh['C_K'] + (normA_cl ** 2) * iQ['ratio']
    N_min = N_0 - math.log((my_term - 1) * st['gamma']) / math.log(st['rho_gamma'])
    return _omega_eta_core(N, normA, iQ, st, my_term, L_V, N_0, G_A, N_min)


def fc_ec_h(e_A, e_B, Q, R):
    """utils.py:526-538."""
    return (e_A ** 2) / my_eigen(Q)['min'] + (e_B ** 2) / my_eigen(R)['min']


def local_radius(F_u, K, Q):
    """utils.py:548-564 — epsilon_K = 1 / max_i ||(F_u K)_i||^2_{Q^-1}; Q^-1 comes from the device preparation."""
    F_u = np.atleast_2d(np.asarray(F_u, dtype=np.float64))
    K = np.atleast_2d(np.asarray(K, dtype=np.float64))
    n = K.shape[1]
    Qinv = _rt.problem_for(np.eye(n), np.eye(n, 1), Q, np.eye(1), scratch=True).prepared()['Qinv']
    Mx = F_u @ K
    with np.errstate(divide='ignore'):                     # numpy semantics as upstream: a zero gain gives inf
        return float(np.float64(1.0) / np.float64(max(float(r @ Qinv @ r) for r in Mx)))


def ex_stability_bounds(gamma, epsilon_K, M_V):
    """utils.py:567-584."""
    return {'L_V': max(gamma, M_V / epsilon_K), 'N_0': math.ceil(max(0, M_V / epsilon_K - gamma))}


def _vertices(F_u):
    lo, hi = _rt.box_from_F(F_u)
    if not (np.all(np.isfinite(lo)) and np.all(np.isfinite(hi))):
        raise ValueError("input set is unbounded")
    return lo, hi


def bar_u_solve(F_u):
    """utils.py:592-619 (Gurobi non-convex QP) — the maximum of the convex ||u||^2 over a polytope sits at a vertex:
    closed form for a box, vertex enumeration otherwise."""
    if not _rt.is_box(F_u):
        V = _rt.polytope_vertices(F_u)
        return float(np.max(np.sum(V * V, axis=1)))
    lo, hi = _vertices(F_u)
    return float(np.sum(np.maximum(lo * lo, hi * hi)))


def bar_d_u_solve(F_u):
    """utils.py:622-650 — max ||u1 - u2||^2 over the polytope x itself: a pair of vertices."""
    if not _rt.is_box(F_u):
        V = _rt.polytope_vertices(F_u)
        D = V[:, None, :] - V[None, :, :]
        return float(np.max(np.sum(D * D, axis=2)))
    lo, hi = _vertices(F_u)
    return float(np.sum((hi - lo) ** 2))


def rot_2D(theta):
    """utils.py:658-666."""
    return np.array([[np.cos(theta), -np.sin(theta)], [np.sin(theta), np.cos(theta)]])


def rot_action_2D(x, theta):
    """utils.py:669-680."""
    out = np.zeros([2, theta.shape[0]])
    for i in range(theta.shape[0]):
        out[:, i:i + 1] = rot_2D(theta[i]) @ x
    return out


def circle_generator(N_points, ratio_ext_radius, my_base, Q):
    """utils.py:683-704 — host-side (negligible, 2-D only). The reference inverts scipy's `cho_factor` output whole,
    i.e. the upper Cholesky factor with Q's strict lower triangle left in place; reproduced literally."""
    Q = np.asarray(Q, dtype=np.float64)
    root_Q = np.triu(np.linalg.cholesky(Q).T) + np.tril(Q, -1)
    x0_base = np.array([[ratio_ext_radius * math.sqrt(my_base)], [0.0]])
    my_theta = np.linspace(0, 2 * (1 - 1 / N_points) * math.pi, N_points)
    return np.linalg.inv(root_Q) @ rot_action_2D(x0_base, my_theta)


def N_incremental_test(A, Q, gamma, rho_gamma):
    """utils.py:712-718."""
    iQ = my_eigen(Q)
    n = np.atleast_2d(A).shape[0]
    normA = _detail(A, np.eye(n, 1), np.eye(n), np.eye(1), K=np.zeros((1, n)))['norm_A']
    return math.ceil(-math.log((normA ** 2) * iQ['ratio'] * gamma) / math.log(rho_gamma))


def default_color_generator():
    """utils.py:726-742 (matplotlib's default colour cycle)."""
    cyc = [(31, 119, 180), (255, 127, 14), (44, 160, 44), (214, 39, 40), (148, 103, 189), (140, 86, 75),
           (227, 119, 194), (127, 127, 127), (188, 189, 34), (23, 190, 207)]
    return {'C%d' % i: tuple(c / 255 for c in rgb) for i, rgb in enumerate(cyc)}


def gradient_color(Value, color_base):
    """utils.py:745-757."""
    lo, hi = Value.min(), Value.max()
    return [tuple(((z - lo) / (hi - lo)) * c for c in color_base[:3]) + (1,) for z in Value.flatten()]


def generate_random_matrix(rows, cols, a, b):
    """utils.py:760-776 (unseeded `random.uniform`, as in the reference; see sampling.py for the seeded sampler)."""
    return np.array([[random.uniform(a, b) for _ in range(cols)] for _ in range(rows)])


def random_matrix(M, N_matrix, norm_bound, norm_type):
    """utils.py:779-823: 5*N_matrix matrices with ||.|| <= norm_bound, the first N_matrix on the boundary."""
    from .sampling import reference_random_matrix
    return reference_random_matrix(M, N_matrix, norm_bound, norm_type)


def error_matrix_generator(A, B, error_vec, N_matrix, norm_type):
    """utils.py:826-847 (writes error_A_<t>.npy / error_B_<t>.npy to cwd like the reference)."""
    out_A = np.zeros([A.shape[0], A.shape[1], 5 * N_matrix, len(error_vec)])
    out_B = np.zeros([B.shape[0], B.shape[1], 5 * N_matrix, len(error_vec)])
    for i in range(len(error_vec)):
        out_A[:, :, :, i] = random_matrix(A, N_matrix, error_vec[i], norm_type)
        out_B[:, :, :, i] = random_matrix(B, N_matrix, error_vec[i], norm_type)
    np.save('error_A' + '_' + norm_type + '.npy', out_A)
    np.save('error_B' + '_' + norm_type + '.npy', out_B)
    return {'error_A': out_A, 'error_B': out_B}


def find_closest_index(a_vec, b):
    """utils.py:850-873."""
    if not a_vec.all():
        return None
    if b <= a_vec[0]:
        return 0
    if b >= a_vec[-1]:
        return len(a_vec) - 1
    idx = bisect.bisect_left(a_vec, b)
    return min(idx, idx - 1, key=lambda i: abs(a_vec[i] - b))


def column_statistics(y_data):
    """The four reductions of utils.py:895-898 (max, min, mean, std over axis 0), computed by K5 on the GPU."""
    from .stats import column_stats
    eng = _rt.get_engine()
    t = np.ascontiguousarray(np.asarray(y_data, dtype=np.float64).T)   # [cols][S]
    st = column_stats(eng, t)
    return st['max'], st['min'], st['mean'], st['std']


def statistical_continuous_kernel(ax, x_data, y_data, info_text, info_color, marker=False):
    """utils.py:881-935 — presentation only (out of scope); the statistics it draws come from K5."""
    y_max, y_min, y_mean, y_std = column_statistics(y_data)
    bound_c = tuple(x * 0.75 for x in info_color)
    ax.plot(x_data, y_mean, label=info_text['data'], linewidth=2.5, color=info_color,
            **({'marker': 'x', 'markersize': 10} if marker else {}))
    for y in (y_min, y_max):
        ax.plot(x_data, y, linewidth=1.5, linestyle=':', color=bound_c)
    for y in (y_mean - y_std, y_mean + y_std):
        ax.plot(x_data, y, linewidth=1.5, linestyle='--', color=bound_c)
    ax.fill_between(x_data, y_mean - y_std, y_mean + y_std, color=tuple(x * 0.5 for x in info_color), alpha=0.25)
    ax.fill_between(x_data, y_min, y_max, color=tuple(x * 0.25 for x in info_color), alpha=0.125)


def statistical_continuous(ax, x_data, y_data, info_text, info_color, font_type, font_size, info_zoom,
                           marker=False, x_scale_log=False, y_scale_log=False, set_x_ticks=False):
    """utils.py:938-1027 — presentation only (no zoom inset here; out of scope)."""
    statistical_continuous_kernel(ax, x_data, y_data, info_text, info_color, marker=marker)
    ax.set_title(info_text['title'])
    ax.set_xlabel(info_text['x_label'])
    if set_x_ticks:
        ax.set_xticks(x_data)
    ax.legend(loc='upper left')
    if x_scale_log:
        ax.set_xscale('log')
    if y_scale_log:
        ax.set_yscale('log')
