"""ctypes wrappers around the g++ host harness (TEST FIXTURE ONLY; see hostmath.cpp)."""
import ctypes

import numpy as np

from . import build as _build

_P = ctypes.POINTER(ctypes.c_double)
_I32 = ctypes.POINTER(ctypes.c_int32)
_lib = None

BOUND_FIELDS = ['alpha', 'beta', 'xi', 'eta', 'bound', 'E_psi', 'E_u', 'E_psi_u', 'theta_u', 'theta_x_u', 'C_K',
                'rho_K', 'gamma', 'rho_gamma', 'L_V', 'N_0', 'omega_N1', 'omega_N0d5', 'err_th', 'N_min', 'h',
                'epsilon_K', 'norm_A', 'norm_B', 'norm_K', 'norm_Gamma', 'norm_Phi', 'min_H', 'rho_cl', 'bar_u',
                'bar_d_u']


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
    return _lib


def _c(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data_as(_P)


def spectral_radius(M):
    M = _c(M)
    S, n = M.shape[0], M.shape[1]
    rho = np.zeros(S)
    ok = np.zeros(S, dtype=np.int32)
    rc = lib().hm_spectral_radius(n, ctypes.c_int64(S), _p(M), _p(rho), ok.ctypes.data_as(_I32))
    assert rc == 0
    return rho, ok


def spectral_radius_poly(M):
    """Closed-form fast path alone (n = 3, 4): (rho, trusted)."""
    M = _c(M)
    S, n = M.shape[0], M.shape[1]
    rho = np.zeros(S)
    ok = np.zeros(S, dtype=np.int32)
    rc = lib().hm_spectral_radius_poly(n, ctypes.c_int64(S), _p(M), _p(rho), ok.ctypes.data_as(_I32))
    assert rc == 0
    return rho, ok


def eval_batch(A, B, Q, R, Pt, N_opc, dA_soa, dB_soa, x0_soa, Nmin, Nmax, T):
    n, m = B.shape
    S = dA_soa.shape[-1]
    H = Nmax - Nmin + 1
    keep = [_c(a) for a in (A, B, Q, R, Pt, dA_soa, dB_soa, x0_soa)]
    o = {k: np.zeros((H, S)) for k in ("J", "rho", "ratio", "V_N", "J_T")}
    o["flags"] = np.zeros((H, S), dtype=np.int32)
    o["K0"] = np.zeros((H, m * n, S))
    o["prep"] = np.zeros(2 * n * n + 4)
    rc = lib().hm_eval(n, m, *[_p(a) for a in keep[:5]], N_opc, ctypes.c_int64(S), *[_p(a) for a in keep[5:]], Nmin,
                       Nmax, T, _p(o["J"]), _p(o["rho"]), _p(o["ratio"]), _p(o["V_N"]), _p(o["J_T"]),
                       o["flags"].ctypes.data_as(_I32), _p(o["K0"]), _p(o["prep"]))
    assert rc == 0
    return o


def mpc(mode, A, B, Q, R, Pt, lo, hi, dA_soa, dB_soa, N, T=0, pts=None, x0_soa=None, x_ref=None, u_ref=None,
        F_u=None):
    """mode 0: open-loop solves; mode 1: closed-loop simulate. x_ref (n, >=N) / u_ref (m, >=N): shared references.
    F_u (p, m): general input polytope F_u u <= 1 (then lo / hi are ignored by the solver)."""
    xr, ur = _c(x_ref), _c(u_ref)
    Fp = _c(F_u)
    if Fp is not None:
        lib().hm_set_poly(_p(Fp), Fp.shape[0])
    if xr is not None or ur is not None:
        ld = (xr if xr is not None else ur).shape[1]
        assert (xr is None or xr.shape[1] == ld) and (ur is None or ur.shape[1] == ld) and ld >= N
        lib().hm_set_refs(_p(xr), _p(ur), ld)
    try:
        return _mpc(mode, A, B, Q, R, Pt, lo, hi, dA_soa, dB_soa, N, T, pts, x0_soa)
    finally:
        lib().hm_set_refs(None, None, 0)
        lib().hm_set_poly(None, 0)


def _mpc(mode, A, B, Q, R, Pt, lo, hi, dA_soa, dB_soa, N, T=0, pts=None, x0_soa=None):
    n, m = B.shape
    S = 1 if dA_soa is None else dA_soa.shape[-1]
    if dA_soa is None and x0_soa is not None:
        S = x0_soa.shape[-1]
    npts = 0 if pts is None else np.asarray(pts).reshape(-1, n).shape[0]
    Pn = max(npts, 1)
    keep = [_c(a) for a in (A, B, Q, R, Pt, lo, hi, dA_soa, dB_soa, pts, x0_soa)]
    pp = [_p(a) for a in keep]
    o = {"V": np.zeros((Pn, S)), "u0": np.zeros((Pn, m, S)), "M_V": np.zeros(S), "J_T": np.zeros(S),
         "X": np.zeros((T + 1, n, S)), "U": np.zeros((max(T, 1), m, S)), "flags": np.zeros((Pn, S), dtype=np.int32),
         "n_active": np.zeros(S, dtype=np.int32)}
    rc = lib().hm_mpc(mode, n, m, *pp[:7], ctypes.c_int64(S), pp[7], pp[8], N, T, npts, pp[9], pp[10], _p(o["V"]),
                      _p(o["u0"]), _p(o["M_V"]), _p(o["J_T"]), _p(o["X"]), _p(o["U"]),
                      o["flags"].ctypes.data_as(_I32), o["n_active"].ctypes.data_as(_I32))
    assert rc == 0
    o["U"] = o["U"][:T]
    return o


def bounds(A, B, Q, R, lo, hi, dA_soa, dB_soa, N, eA, eB, MV, x_soa, K_in, p, V_expert, strict=1):
    n, m = B.shape
    S = dA_soa.shape[-1]
    keep = [_c(a) for a in (A, B, Q, R, Q, lo, hi, dA_soa, dB_soa, eA, eB, MV, x_soa, K_in, p)]
    pp = [_p(a) for a in keep]
    NF = lib().hm_bounds_fields()
    assert NF == len(BOUND_FIELDS)
    det = np.zeros((NF, S))
    Ko = np.zeros((m * n, S))
    Po = np.zeros((n * n, S))
    fl = np.zeros(S, dtype=np.int32)
    rc = lib().hm_bounds(n, m, *pp[:7], ctypes.c_int64(S), pp[7], pp[8], N, *pp[9:15], ctypes.c_double(V_expert),
                         strict, _p(det), _p(Ko), _p(Po), fl.ctypes.data_as(_I32))
    assert rc == 0
    out = {k: det[i] for i, k in enumerate(BOUND_FIELDS)}
    out.update({"K": Ko, "P": Po, "flags": fl})
    return out


def philox4x32_10(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32); k = np.asarray(key, dtype=np.uint32); o = np.zeros(4, dtype=np.uint32)
    U32 = ctypes.POINTER(ctypes.c_uint32)
    lib().hm_philox4x32_10(c.ctypes.data_as(U32), k.ctypes.data_as(U32), o.ctypes.data_as(U32))
    return o


def sample_error_grid(seed, which, rows, cols, N_sys, levels, n_boundary, norm_type="f", j_first=0):
    lev = _c(levels)
    out = np.zeros((rows * cols, N_sys * lev.size))
    st = np.zeros(2, dtype=np.int64)
    rc = lib().hm_sample_error_grid(ctypes.c_uint64(seed), which, rows, cols, ctypes.c_int64(N_sys),
                                    ctypes.c_int64(j_first), lev.size, _p(lev), ctypes.c_int64(n_boundary),
                                    {"f": 0, "2": 1}[norm_type], _p(out), st.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    assert rc == 0
    return out.reshape(rows, cols, N_sys, lev.size), int(st[0]), int(st[1])


def tridiag_extremes(d, e):
    """(lambda_min, lambda_max) of the symmetric tridiagonal with diagonal d and off-diagonal e[1:] (e[0] unused)."""
    d, e = _c(d), _c(e)
    out = np.zeros(2)
    lib().hm_tridiag_extremes(_p(d), _p(e), len(d), _p(out))
    return out[0], out[1]


def seeded_samples(seed, first, S, n, m, e_A, e_B):
    """lq::seeded_sample (the generator behind lqmpc_eval_seeded): dA (S,n,n), dB (S,n,m), x0 (S,n)."""
    dA, dB, x0 = np.zeros((S, n, n)), np.zeros((S, n, m)), np.zeros((S, n))
    rc = lib().hm_seeded_samples(n, m, ctypes.c_uint64(seed), ctypes.c_int64(first), ctypes.c_int64(S),
                                 ctypes.c_double(e_A), ctypes.c_double(e_B), _p(dA), _p(dB), _p(x0))
    assert rc == 0
    return dA, dB, x0
