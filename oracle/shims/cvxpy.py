"""Shim (oracle only) for the sliver of cvxpy that LQ_MPC_Controller uses (utils_class.py:46-91):
Variable((m,N)), column indexing u[:, i], affine arithmetic, quad_form, Minimize, Problem.solve, and
constraints of the form  F @ u[:, i] <= 1.

cvxpy (un-pinned, absent here) would canonicalise the same convex QP and hand it to OSQP/Clarabel; the
minimiser of a strictly convex QP is unique, so this shim assembles the dense QP
    min_z  z'Hz + 2 g'z + c0   s.t.  lo <= z <= hi       (z time-major: z[i*m:(i+1)*m] = u[:, i])
and solves it EXACTLY (Cholesky + BVLS active set, tol 1e-15) when every constraint row has a single non-zero — the
only kind the reference's own scripts build (F_u = [[10],[-10]] / +-10*I). Rows with several non-zeros (a general
input polytope) go to the dense primal active-set solver of oracle/qp_dense.py instead.
"""
import numpy as np
import scipy.linalg as sla
from scipy.optimize import lsq_linear


class _Aff:
    """value = M @ z + c"""
    __array_priority__ = 1000

    def __init__(self, var, M, c):
        self.var, self.M, self.c = var, M, c

    def _lift(self, other):
        if isinstance(other, _Aff):
            return other.M, other.c
        o = np.asarray(other, dtype=float)
        return np.zeros((o.shape[0], self.M.shape[1])), o

    def __add__(self, other):
        M, c = self._lift(other)
        return _Aff(self.var, self.M + M, self.c + c)

    __radd__ = __add__

    def __sub__(self, other):
        M, c = self._lift(other)
        return _Aff(self.var, self.M - M, self.c - c)

    def __rsub__(self, other):
        M, c = self._lift(other)
        return _Aff(self.var, M - self.M, c - self.c)

    def __rmatmul__(self, A):
        A = np.asarray(A, dtype=float)
        return _Aff(self.var, A @ self.M, A @ self.c)

    def __le__(self, rhs):
        return ("le", self, rhs)

    @property
    def value(self):
        return self.M @ self.var._z + self.c


class Variable:
    def __init__(self, shape):
        self.m, self.N = shape
        self._z = np.zeros(self.m * self.N)

    def __getitem__(self, idx):
        sl, i = idx
        assert sl == slice(None)
        M = np.zeros((self.m, self.m * self.N))
        M[:, i * self.m:(i + 1) * self.m] = np.eye(self.m)
        return _Aff(self, M, np.zeros(self.m))


class _Quad:
    def __init__(self, terms):
        self.terms = terms

    def __add__(self, other):
        if isinstance(other, _Quad):
            return _Quad(self.terms + other.terms)
        if other == 0:
            return self
        raise TypeError(other)

    __radd__ = __add__


def quad_form(aff, P):
    return _Quad([(aff, np.asarray(P, dtype=float))])


class Minimize:
    def __init__(self, q):
        self.q = q
        self.value = None


class Problem:
    def __init__(self, objective, constraints):
        self.obj, self.cons = objective, constraints

    def solve(self):
        terms = self.obj.q.terms
        var = terms[0][0].var
        nz = var.m * var.N
        H = np.zeros((nz, nz))
        g = np.zeros(nz)
        c0 = 0.0
        for aff, P in terms:
            H += aff.M.T @ P @ aff.M
            g += aff.M.T @ P @ aff.c
            c0 += aff.c @ P @ aff.c
        lo = np.full(nz, -np.inf)
        hi = np.full(nz, np.inf)
        if any(np.count_nonzero(aff.M[r]) > 1 for _, aff, _ in self.cons for r in range(aff.M.shape[0])):
            from oracle.qp_dense import ineq_qp
            C = np.vstack([aff.M for _, aff, _ in self.cons])
            d = np.concatenate([np.broadcast_to(np.asarray(rhs, dtype=float).ravel(), (aff.M.shape[0],)) - aff.c
                                for _, aff, rhs in self.cons])
            H = 0.5 * (H + H.T)
            z, _ = ineq_qp(H, g, C, d)
            var._z = z
            self.obj.value = float(z @ H @ z + 2 * g @ z + c0)
            return self.obj.value
        for kind, aff, rhs in self.cons:
            assert kind == "le"
            rhs = np.broadcast_to(np.asarray(rhs, dtype=float).ravel(), (aff.M.shape[0],))
            for r in range(aff.M.shape[0]):
                nzc = np.flatnonzero(aff.M[r])
                if len(nzc) != 1:
                    raise NotImplementedError("cvxpy shim supports box constraints only")
                j = nzc[0]
                b = (rhs[r] - aff.c[r]) / aff.M[r, j]
                if aff.M[r, j] > 0:
                    hi[j] = min(hi[j], b)
                else:
                    lo[j] = max(lo[j], b)
        H = 0.5 * (H + H.T)
        L = np.linalg.cholesky(H)
        # z'Hz + 2g'z = ||L'z + L^-1 g||^2 - const
        rhs_ls = -sla.solve_triangular(L, g, lower=True)
        res = lsq_linear(L.T, rhs_ls, bounds=(lo, hi), method="bvls", tol=1e-15, max_iter=10 * nz + 50)
        z = res.x
        var._z = z
        self.obj.value = float(z @ H @ z + 2 * g @ z + c0)
        return self.obj.value
