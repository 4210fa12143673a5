"""TEST INFRASTRUCTURE ONLY — batched-numpy oracle of the unconstrained evaluation path (K1).

Same mathematics as oracle/np_oracle.py (`riccati`, `closed_loop_inf_cost`), vectorised over the sample axis with
numpy/LAPACK so that mid-size parity checks (1e4..1e5 samples) and the CPU baseline of bench.py finish in seconds.
Checked against the per-sample oracle in tests/test_oracle_golden.py.  Arrays here are AoS: dA (S,n,n), dB (S,n,m),
x0 (S,n) — the engine's SoA layout is produced by `to_soa`.
"""
from __future__ import annotations

import numpy as np


def to_soa(dA, dB, x0):
    """(S,n,n),(S,n,m),(S,n) -> [n*n][S], [n*m][S], [n][S] contiguous (element-major)."""
    S = dA.shape[0]
    return (np.ascontiguousarray(dA.reshape(S, -1).T), np.ascontiguousarray(dB.reshape(S, -1).T),
            np.ascontiguousarray(x0.reshape(S, -1).T))


def expert_matrix(A, B, Q, R, P_term, N_opc):
    """Cost-to-go matrix of the TRUE model after N_opc Riccati steps (utils_class.py:757,786 for an inactive box);
    N_opc <= 0: iterate to the DARE fixed point."""
    P = np.array(P_term, dtype=float)
    it = N_opc if N_opc > 0 else 100000
    for _ in range(it):
        G = R + B.T @ P @ B
        H = B.T @ P @ A
        Pn = Q + A.T @ P @ A - H.T @ np.linalg.solve(G, H)
        Pn = 0.5 * (Pn + Pn.T)
        if N_opc <= 0 and np.max(np.abs(Pn - P)) <= 1e-16 * np.max(np.abs(Pn)):
            P = Pn
            break
        P = Pn
    return P


def lyapunov_doubling(Acl, W, iters=64):
    """S = sum_k (Acl')^k W Acl^k for a batch of stable Acl (S,n,n)."""
    S = W.copy()
    M = Acl.copy()
    for _ in range(iters):
        T = np.swapaxes(M, 1, 2) @ S @ M
        S = S + T
        if np.max(np.abs(T)) <= 1e-18 * np.max(np.abs(S)):
            break
        M = M @ M
    return 0.5 * (S + np.swapaxes(S, 1, 2))


def eval_batch(A, B, Q, R, P_term, P_exp, dA, dB, x0, N_min, N_max, T=0, want_K=False):
    """Returns dict of (H,S) arrays J, rho, ratio, Vn, JT (T>0), unstable mask and optionally K0 (H,S,m,n)."""
    S = dA.shape[0]
    n, m = B.shape
    H = N_max - N_min + 1
    Ah = A[None] + dA
    Bh = B[None] + dB
    P = np.broadcast_to(np.asarray(P_term, dtype=float), (S, n, n)).copy()
    out = {k: np.zeros((H, S)) for k in ("J", "rho", "ratio", "Vn", "JT")}
    out["unstable"] = np.zeros((H, S), dtype=bool)
    if want_K:
        out["K0"] = np.zeros((H, S, m, n))
    vexp = np.einsum("si,ij,sj->s", x0, P_exp, x0)
    BhT = np.swapaxes(Bh, 1, 2)
    AhT = np.swapaxes(Ah, 1, 2)
    for k in range(1, N_max + 1):
        PB = P @ Bh
        G = R[None] + BhT @ PB
        Hm = np.swapaxes(PB, 1, 2) @ Ah
        K = -np.linalg.solve(G, Hm)
        Pn = Q[None] + AhT @ P @ Ah + np.swapaxes(Hm, 1, 2) @ K
        Pn = 0.5 * (Pn + np.swapaxes(Pn, 1, 2))
        if k >= N_min:
            h = k - N_min
            Acl = A[None] + B[None] @ K
            W = Q[None] + np.swapaxes(K, 1, 2) @ R[None] @ K
            rho = np.max(np.abs(np.linalg.eigvals(Acl)), axis=1)
            stable = rho < 1.0
            J = np.full(S, np.inf)
            if np.any(stable):
                Sl = lyapunov_doubling(Acl[stable], W[stable])
                J[stable] = np.einsum("si,sij,sj->s", x0[stable], Sl, x0[stable])
            out["J"][h] = J
            out["rho"][h] = rho
            out["ratio"][h] = J / vexp
            out["unstable"][h] = ~stable
            out["Vn"][h] = np.einsum("si,sij,sj->s", x0, Pn, x0)
            if want_K:
                out["K0"][h] = K
            if T > 0:
                x = x0.copy()
                cost = np.einsum("si,ij,sj->s", x, Q, x)
                for _ in range(T):
                    u = np.einsum("sij,sj->si", K, x)
                    x = x @ A.T + u @ B.T
                    cost = cost + np.einsum("si,ij,sj->s", x, Q, x) + np.einsum("si,ij,sj->s", u, R, u)
                out["JT"][h] = cost
        P = Pn
    return out


def synth_problem(n=4, m=2, seed=0, rho_target=1.05):
    """cfg-synth: A_true = G * rho_target / rho(G), G ~ N(0,1); B_true ~ N(0,1); Q = I, R = I, P = Q.
    Re-drawn until (A,B) is controllable with cond(ctrb) < 1e6."""
    rng = np.random.default_rng(seed)
    while True:
        G = rng.normal(size=(n, n))
        A = G * (rho_target / np.max(np.abs(np.linalg.eigvals(G))))
        B = rng.normal(size=(n, m))
        C = np.hstack([np.linalg.matrix_power(A, k) @ B for k in range(n)])
        if np.linalg.matrix_rank(C) == n and np.linalg.cond(C @ C.T) < 1e12:
            break
    return A, B, np.eye(n), np.eye(m)


def synth_samples(n, m, S, seed=1, first=0, e=0.01, block=1 << 16):
    """Seeded, shard-invariant samples: block b of `block` samples is drawn from Philox(key=[seed, 0x4c514d50],
    counter=[0,0,0,b]) — dA, dB ~ U[-e, e] entrywise, x0 ~ N(0, I) — so any rank asking for samples
    [first, first+S) sees exactly the values a single-GPU run would."""
    dA = np.empty((S, n, n))
    dB = np.empty((S, n, m))
    x0 = np.empty((S, n))
    b0, b1 = first // block, (first + S - 1) // block
    pos = 0
    for b in range(b0, b1 + 1):
        g = np.random.Generator(np.random.Philox(key=[seed, 0x4C514D50], counter=[0, 0, 0, b]))
        a = g.uniform(-e, e, size=(block, n, n))
        bb = g.uniform(-e, e, size=(block, n, m))
        xx = g.standard_normal(size=(block, n))
        lo = max(first, b * block) - b * block
        hi = min(first + S, (b + 1) * block) - b * block
        cnt = hi - lo
        dA[pos:pos + cnt] = a[lo:hi]
        dB[pos:pos + cnt] = bb[lo:hi]
        x0[pos:pos + cnt] = xx[lo:hi]
        pos += cnt
    return dA, dB, x0
