"""TEST INFRASTRUCTURE ONLY — extended-precision (mpmath, 50 digits) arbitration of the hot path.

Where the CUDA engine and the float64 numpy oracle disagree by more than north_star's 1e-9, neither is trusted: the
quantity is recomputed here in 50-digit arithmetic from the SAME float64 inputs and both are measured against it.
This is what turns "the comparison is ill-conditioned" from an assertion into a measured statement:
  * `k1_truth`   Riccati gain of horizon N on (A+dA, B+dB), closed loop on (A, B), J_inf by a direct solve of the
                 discrete Lyapunov equation (vec form), rho by mpmath's eigenvalue solver, V_N — the quantities of
                 utils_class.py:48-91, 245-285 for an inactive input set;
  * `qp_truth`   the box-constrained condensed QP of utils_class.py:59-88 solved on a GIVEN active set in 50 digits,
                 with the KKT sign conditions re-checked there (so the active set is certified, not assumed).
Slow (pure Python): meant for a handful of samples.
"""
from __future__ import annotations

import mpmath as mp
import numpy as np

mp.mp.dps = 50


def _M(a):
    a = np.atleast_2d(np.asarray(a, dtype=float))
    return mp.matrix(a.tolist())


def _riccati_gain(Ah, Bh, Q, R, Pt, N):
    P = Pt.copy()
    K = None
    for _ in range(N):
        G = R + Bh.T * P * Bh
        H = Bh.T * P * Ah
        K = -mp.lu_solve(G, H)
        P = Q + Ah.T * P * Ah + H.T * K
        P = (P + P.T) / 2
    return K, P


def k1_truth(A, B, Q, R, Pt, dA, dB, x0, N):
    """Returns dict(J, rho, Vn, K0) as Python floats / arrays rounded from 50-digit results."""
    A, B, Q, R, Pt = _M(A), _M(B), _M(Q), _M(R), _M(Pt)
    n, m = A.rows, B.cols
    Ah, Bh = A + _M(dA), B + _M(np.asarray(dB, dtype=float).reshape(n, m))
    x = mp.matrix([mp.mpf(float(v)) for v in np.asarray(x0, dtype=float).reshape(-1)])
    K, P = _riccati_gain(Ah, Bh, Q, R, Pt, N)
    Acl = A + B * K
    W = Q + K.T * R * K
    ev, _ = mp.eig(Acl)
    rho = max(abs(e) for e in ev)
    out = {"rho": float(rho), "Vn": float((x.T * P * x)[0]),
           "K0": np.array([[float(K[i, j]) for j in range(n)] for i in range(m)])}
    if rho >= 1:
        out["J"] = float("inf")
        return out
    # S = W + Acl' S Acl  <=>  (I - Acl' (x) Acl') vec(S) = vec(W)   (row-major vec)
    AT = Acl.T
    big = mp.eye(n * n)
    for i in range(n):
        for j in range(n):
            for k in range(n):
                for l in range(n):
                    big[i * n + j, k * n + l] -= AT[i, k] * AT[j, l]
    vw = mp.matrix([W[i, j] for i in range(n) for j in range(n)])
    vs = mp.lu_solve(big, vw)
    S = mp.matrix(n, n)
    for i in range(n):
        for j in range(n):
            S[i, j] = vs[i * n + j]
    out["J"] = float((x.T * S * x)[0])
    return out


def j_condition_bound(rho):
    """Relative rounding-error allowance for J_inf = x0' S x0 in float64: perturbing A_cl by one rounding error u moves
    rho by ~u and J ~ sum rho^(2k) by dJ/J ~ 2 rho d(rho) / (1 - rho^2), i.e. the relative condition number is
    ~rho / (1 - rho); 16 u rho/(1 - rho) allows a handful of such errors. It exceeds 1e-9 only within 1.8e-6 of the
    stability boundary."""
    u = 2.0 ** -53
    return max(1e-9, 16.0 * u * rho / max(1.0 - rho, 1e-300))


def qp_truth(N, A, B, Q, R, P, lo, hi, x0, active_lo, active_hi):
    """Condensed QP min z'Hz + 2g'z (+ c0 + x0'Qx0) with the components in `active_lo` / `active_hi` clamped, solved in
    50 digits. Returns (u0, V, certified): certified = every free component strictly inside its bounds (to 1e-30) and
    every clamped one with the right multiplier sign."""
    A, B, Q, R, P = _M(A), _M(B), _M(Q), _M(R), _M(P)
    n, m = A.rows, B.cols
    nz = N * m
    # Gamma rows for x_1..x_N and Phi
    Gam = mp.zeros(N * n, nz)
    AB = B.copy()
    for d in range(N):
        for j in range(N - d):
            i = j + d
            for r in range(n):
                for c in range(m):
                    Gam[i * n + r, j * m + c] = AB[r, c]
        AB = A * AB
    x = mp.matrix([mp.mpf(float(v)) for v in np.asarray(x0, dtype=float).reshape(-1)])
    f = mp.zeros(N * n, 1)
    Mk = A.copy()
    for k in range(N):
        v = Mk * x
        for r in range(n):
            f[k * n + r] = v[r]
        Mk = A * Mk
    Wb = mp.zeros(N * n, N * n)
    for k in range(N):
        blk = P if k == N - 1 else Q
        for r in range(n):
            for c in range(n):
                Wb[k * n + r, k * n + c] = blk[r, c]
    H = Gam.T * Wb * Gam
    for k in range(N):
        for r in range(m):
            for c in range(m):
                H[k * m + r, k * m + c] += R[r, c]
    g = Gam.T * Wb * f
    c0 = (f.T * Wb * f)[0] + (x.T * Q * x)[0]
    lo_t = [mp.mpf(float(lo[j % m])) for j in range(nz)]
    hi_t = [mp.mpf(float(hi[j % m])) for j in range(nz)]
    z = mp.zeros(nz, 1)
    fixed = {}
    for j in active_lo:
        fixed[int(j)] = lo_t[int(j)]
    for j in active_hi:
        fixed[int(j)] = hi_t[int(j)]
    free = [j for j in range(nz) if j not in fixed]
    for j, v in fixed.items():
        z[j] = v
    if free:
        Hff = mp.matrix(len(free), len(free))
        rhs = mp.matrix(len(free), 1)
        for a, ja in enumerate(free):
            acc = -g[ja]
            for jb, v in fixed.items():
                acc -= H[ja, jb] * v
            rhs[a] = acc
            for b, jb in enumerate(free):
                Hff[a, b] = H[ja, jb]
        zf = mp.lu_solve(Hff, rhs)
        for a, ja in enumerate(free):
            z[ja] = zf[a]
    grad = 2 * (H * z + g)
    ok = True
    for j in free:
        if not (lo_t[j] - mp.mpf(10) ** -30 <= z[j] <= hi_t[j] + mp.mpf(10) ** -30):
            ok = False
    for j in active_lo:
        if grad[int(j)] < -mp.mpf(10) ** -30:
            ok = False
    for j in active_hi:
        if grad[int(j)] > mp.mpf(10) ** -30:
            ok = False
    V = (z.T * H * z)[0] + 2 * (g.T * z)[0] + c0
    return np.array([float(z[j]) for j in range(m)]), float(V), ok


# ------------------------------------------------------------------------------------------------ longdouble (n = 32)
def _ld_solve(G, H):
    """Gaussian elimination with partial pivoting in numpy longdouble (x87 80-bit: 64-bit mantissa)."""
    G = G.copy()
    H = H.copy()
    k = G.shape[0]
    for c in range(k):
        p = c + int(np.argmax(np.abs(G[c:, c])))
        if p != c:
            G[[c, p]] = G[[p, c]]
            H[[c, p]] = H[[p, c]]
        for r in range(c + 1, k):
            f = G[r, c] / G[c, c]
            G[r, c:] -= f * G[c, c:]
            H[r] -= f * H[c]
    X = np.zeros_like(H)
    for r in range(k - 1, -1, -1):
        X[r] = (H[r] - G[r, r + 1:] @ X[r + 1:]) / G[r, r]
    return X


def _k1_ld(A, B, Q, R, Pt, dA, dB, x0, N):
    ld = np.longdouble
    A, B, Q, R, P = (np.asarray(a, dtype=ld) for a in (A, B, Q, R, Pt))
    n, m = B.shape
    Ah, Bh = A + np.asarray(dA, dtype=ld), B + np.asarray(dB, dtype=ld).reshape(n, m)
    x = np.asarray(x0, dtype=ld).reshape(n)
    K = None
    for _ in range(N):
        G = R + Bh.T @ P @ Bh
        H = Bh.T @ P @ Ah
        K = -_ld_solve(G, H)
        P = Q + Ah.T @ P @ Ah + H.T @ K
        P = (P + P.T) / 2
    Acl = A + B @ K
    W = Q + K.T @ R @ K
    rho = float(np.max(np.abs(np.linalg.eigvals(Acl.astype(np.float64)))))
    out = {"rho": rho, "Vn": float(x @ P @ x)}
    if rho >= 1:
        out["J"] = float("inf")
        return out
    S, M = W.copy(), Acl.copy()
    for _ in range(80):
        T = M.T @ S @ M
        S = S + T
        if float(np.max(np.abs(T))) <= 1e-24 * float(np.max(np.abs(S))):
            break
        M = M @ M
    out["J"] = float(x @ S @ x)
    return out


def k1_truth_ld(A, B, Q, R, Pt, dA, dB, x0, N, delta=1e-13, seed=0):
    """Extended-precision (longdouble) K1 values for larger n, plus the MEASURED relative condition number of each
    value with respect to the sample: kappa_X = |X(dA (1 + delta r)) - X(dA)| / (|X| delta), r ~ U[-1, 1] entrywise.
    A backward-stable float64 evaluation cannot be expected to do better than ~kappa u."""
    base = _k1_ld(A, B, Q, R, Pt, dA, dB, x0, N)
    rng = np.random.default_rng(seed)
    pert = np.asarray(dA, dtype=np.longdouble) * (1 + np.longdouble(delta) * rng.uniform(-1, 1, size=np.shape(dA)))
    other = _k1_ld(A, B, Q, R, Pt, pert, dB, x0, N)
    for k in ("J", "Vn", "rho"):
        if np.isfinite(base[k]) and np.isfinite(other[k]) and base[k] != 0:
            base["kappa_" + k] = abs(other[k] - base[k]) / (abs(base[k]) * delta)
        else:
            base["kappa_" + k] = float("inf")
    return base
