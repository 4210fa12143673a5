"""TEST INFRASTRUCTURE ONLY — exact dense solver for the strictly convex QP with general linear inequalities
    min_z  z'Hz + 2 g'z   s.t.  C z <= d ,   d > 0 (so z = 0 is strictly feasible),
the problem LQ_MPC_Controller.solve hands to cvxpy when F_u is a general polytope (utils_class.py:59-88:
C = kron(I_N, F_u), d = 1). Textbook primal active-set method (Nocedal & Wright, Alg. 16.3) on dense KKT systems —
deliberately a different route from the engine's stage-wise Riccati sweeps. The minimiser of a strictly convex QP is
unique, so any exact method must agree with whatever cvxpy's back end returns."""
import numpy as np


def ineq_qp(H, g, C, d, tol=1e-13, max_iter=None):
    H = 0.5 * (np.asarray(H, dtype=float) + np.asarray(H, dtype=float).T)
    g = np.asarray(g, dtype=float)
    C = np.atleast_2d(np.asarray(C, dtype=float))
    d = np.asarray(d, dtype=float)
    nz, nc = H.shape[0], C.shape[0]
    z = np.zeros(nz)
    W = []
    max_iter = max_iter or 20 * (nz + nc) + 50
    for _ in range(max_iter):
        # equality-constrained QP on the working set
        if W:
            Cw = C[W]
            KKT = np.block([[2 * H, Cw.T], [Cw, np.zeros((len(W), len(W)))]])
            rhs = np.concatenate([-2 * g, d[W]])
            try:
                sol = np.linalg.solve(KKT, rhs)
            except np.linalg.LinAlgError:
                sol = np.linalg.lstsq(KKT, rhs, rcond=None)[0]
            zs, mu = sol[:nz], sol[nz:]
        else:
            zs, mu = np.linalg.solve(2 * H, -2 * g), np.zeros(0)
        dz = zs - z
        Cz, Cdz = C @ z, C @ dz
        alpha, block = 1.0, -1
        for i in range(nc):
            if i in W:
                continue
            if Cz[i] + Cdz[i] > d[i] + tol * max(1.0, abs(d[i])) and Cdz[i] > 0:
                a = (d[i] - Cz[i]) / Cdz[i]
                if a < alpha:
                    alpha, block = a, i
        if block >= 0:
            z = z + max(alpha, 0.0) * dz
            W.append(block)
            continue
        z = zs
        if len(mu) == 0 or np.min(mu) >= -1e-11 * (np.max(np.abs(mu)) + 1e-300):
            return z, sorted(W)
        W.pop(int(np.argmin(mu)))
    raise RuntimeError("ineq_qp: iteration budget exhausted")


def polytope_vertices(F_u):
    """Vertices of {u : F_u u <= 1} (bounded, m small): every m-subset of rows with a feasible intersection."""
    import itertools
    F = np.atleast_2d(np.asarray(F_u, dtype=float))
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
        raise ValueError("input polytope has no vertex (unbounded or empty)")
    return np.array(out)
