"""CPU tier — pins the ORACLE (oracle/np_oracle.py, oracle/np_batched.py) before anything trusts it:
  * the reference's shipped result file data_lq_mpc_multipleSys.npz (all 13 arrays, all 1 500 evals; produced by the
    authors with real cvxpy / python-control / Gurobi; committed copy tests/golden/multiple_sys.npz),
  * answers generated in the build container from the UNTOUCHED reference modules (oracle/make_golden.py ->
    tests/golden/ref_known_answers.json, ref_norm2_subset.npz),
  * analytic cross-checks: Riccati vs condensed QP, Lyapunov doubling vs scipy, scalar closed forms.
When /root/reference is mounted (build container only) the shim-backed reference itself is re-run on a slice.
"""
import math
import os

import numpy as np
import pytest
import scipy.linalg as sla

from oracle import np_batched as nb
from oracle import np_oracle as o
from tests.conftest import relerr

TOL = 1e-9
TABLES = ("alpha_table_error", "beta_table_error", "xi_table_error", "bound_table_error", "true_cost_error",
          "alpha_table_horizon", "beta_table_horizon", "xi_table_horizon", "bound_table_horizon",
          "true_cost_horizon")


def _run_oracle(example, eA, eB, max_sys=None):
    return o.data_generation(example["A"], example["B"], example["Q"], example["R"], example["F_u"], eA, eB,
                             example["info_N"], example["info_e"], 8, 1.5, example["p"], max_sys=max_sys)


def test_oracle_reproduces_shipped_golden_file(golden, example):
    """All 13 arrays of the shipped result file (utils_class.py:944-956), every one of the 1 500 evals."""
    out = _run_oracle(example, golden["error_A_f"], golden["error_B_f"])
    assert np.array_equal(out["error"], golden["error"])
    assert np.array_equal(out["horizon"], golden["horizon"])
    assert abs(out["V_expert"] - float(golden["V_expert"])) < TOL * float(golden["V_expert"])
    for k in TABLES:
        assert out[k].shape == golden[k].shape
        # bound = (alpha V + beta)/(1 - xi - eta) amplifies rounding where 1 - xi - eta is small: 1e-9 still holds
        assert relerr(out[k], golden[k]) < TOL, k


def test_oracle_matches_reference_on_norm2_grids(golden, golden_norm2, example):
    """`_2` grids have no shipped outputs; the untouched reference was run on 5 systems per column."""
    n_sys = golden_norm2["true_cost_error"].shape[0]
    out = _run_oracle(example, golden["error_A_2"], golden["error_B_2"], max_sys=n_sys)
    for k in TABLES:
        assert relerr(out[k], golden_norm2[k]) < TOL, k


def test_oracle_working_example_single(known, example):
    """working_example_single.py:39-108 restated with the oracle vs the untouched reference's printed values."""
    k = known["single"]
    A, B, Q, R, lo, hi = (example[x] for x in ("A", "B", "Q", "R", "lo", "hi"))
    K, P = o.dlqr(A, B, Q, R)
    assert relerr(K, k["K_lqr"]) < TOL and relerr(P, k["P_lqr"]) < TOL
    eps = o.local_radius(lo, hi, -K, Q)
    assert abs(eps - k["epsilon_lqr"]) < TOL * eps
    x0_vec = o.circle_generator(8, 1.5, eps, Q)
    assert np.max(np.abs(x0_vec - np.array(k["x0_vec"]))) < 1e-14
    M_V = o.M_V_of(6, A, B, Q, R, lo, hi, x0_vec)
    assert abs(M_V - k["M_V"]) < TOL * M_V
    for p in range(8):
        u0, V, _ = o.mpc_solve(6, A, B, Q, R, Q, lo, hi, x0_vec[:, p])
        assert abs(V - k["ring_V"][p]) < TOL * V and np.max(np.abs(u0 - k["ring_u0"][p])) < 1e-10
    ex = o.ex_stability_lq(A, B, Q, R, -K)
    for key in ("C_K", "rho_K", "gamma", "rho_gamma"):
        assert abs(ex[key] - k["ex"][key]) < TOL * abs(k["ex"][key]), key
    bar = o.ex_stability_bounds(ex["gamma"], eps, M_V)
    assert abs(bar["L_V"] - k["bar"]["L_V"]) < TOL * bar["L_V"] and bar["N_0"] == k["bar"]["N_0"]
    om = o.fc_omega_eta(6, A, B, Q, R, -K, bar["L_V"], bar["N_0"])
    for key in ("omega_N1", "omega_N0d5", "eta", "err_th", "N_min"):
        assert abs(om[key] - k["omega_eta"][key]) < TOL * abs(k["omega_eta"][key]), key
    dec = o.energy_decreasing(A, B, Q, R, lo, hi, 6, 0.01, 0.01, -K, M_V)
    bnd = o.energy_bound(A, B, Q, R, lo, hi, 6, 0.01, 0.01, x0_vec[:, 0], example["p"])
    assert abs(dec["xi"] - k["decrease"]["xi"]) < TOL * dec["xi"]
    assert abs(dec["eta"] - k["decrease"]["eta"]) < TOL * dec["eta"]
    assert abs(bnd["alpha"] - k["bound"]["alpha"]) < TOL * bnd["alpha"]
    assert abs(bnd["beta"] - k["bound"]["beta"]) < TOL * bnd["beta"]
    assert abs(o.bar_u(lo, hi) - k["bar_u"]) < 1e-15 and abs(o.bar_d_u(lo, hi) - k["bar_d_u"]) < 1e-15


def test_oracle_mpc_test_scenario(known):
    """mpc_test.py:13-56: saturated open-loop solve (N=20) and closed loop (T=20, N=6) on a perturbed plant."""
    k = known["mpc_test"]
    A = np.array([[1, 0.7], [0.12, 0.4]]); B = np.array([[1], [1.2]])
    Q = 2 * np.eye(2); R = np.eye(1); lo, hi = np.array([-0.1]), np.array([0.1])
    x0 = np.array([0.1125, 0.19])
    u0, V, act = o.mpc_solve(20, A, B, Q, R, Q, lo, hi, x0)
    assert act and abs(V - k["V_N"]) < TOL * V and abs(u0[0] - k["u_0"][0]) < 1e-12
    sim = o.simulate(20, 6, A, B, Q, R, Q, lo, hi, x0, np.array([[1.01, 0.7], [0.12, 0.41]]),
                     np.array([[1], [1.21]]))
    assert abs(sim["J_T"] - k["J_T"]) < TOL * k["J_T"]
    assert np.max(np.abs(sim["X"] - np.array(k["X"]))) < 1e-12
    assert np.max(np.abs(sim["U"] - np.array(k["U"]))) < 1e-12


def test_oracle_random_api_cases(known):
    """24 randomised calls of the untouched reference classes (n in {2,3}, m in {1,2}); includes math-domain raises."""
    n_raise = 0
    for c in known["random_api"]:
        n, m, N, T = c["n"], c["m"], c["N"], c["T"]
        A, B, dA, dB = (np.array(c[x]) for x in ("A", "B", "dA", "dB"))
        Q, R = c["q"] * np.eye(n), c["r"] * np.eye(m)
        lo, hi = -c["ub"] * np.ones(m), c["ub"] * np.ones(m)
        x0 = np.array(c["x0"])
        for exact_fast in (True, False):          # the Riccati fast path and the dense QP agree with the reference
            u0, V, _ = o.mpc_solve(N, A + dA, B + dB, Q, R, Q, lo, hi, x0, exact_fast=exact_fast)
            assert abs(V - c["V_N"]) < TOL * abs(c["V_N"])
            assert np.max(np.abs(u0 - np.array(c["u_0"]))) < 1e-10
        sim = o.simulate(T, N, A + dA, B + dB, Q, R, Q, lo, hi, x0, A, B)
        assert abs(sim["J_T"] - c["J_T"]) < TOL * abs(c["J_T"])
        assert np.max(np.abs(sim["U"] - np.array(c["U"]))) < 1e-10
        if "K_dlqr" in c:
            K, _ = o.dlqr(A + dA, B + dB, Q, R)
            assert relerr(K, c["K_dlqr"]) < 1e-8
            if "raises" in c:
                with pytest.raises(ValueError):
                    o.energy_decreasing(A + dA, B + dB, Q, R, lo, hi, N, c["e"], c["e"], -K, c["M_V"])
                n_raise += 1
            else:
                dec = o.energy_decreasing(A + dA, B + dB, Q, R, lo, hi, N, c["e"], c["e"], -K, c["M_V"])
                bnd = o.energy_bound(A + dA, B + dB, Q, R, lo, hi, N, c["e"], c["e"], x0, np.array([0.1, 1, 0.6]))
                for key, v in (("xi", dec["xi"]), ("eta", dec["eta"]), ("alpha", bnd["alpha"]),
                               ("beta", bnd["beta"])):
                    assert abs(v - c[key]) < TOL * abs(c[key]), key
    assert n_raise > 0


def test_oracle_tracking_references(tracking):
    """Non-zero x_ref / u_ref (utils_class.py:62-81) — 16 solves + closed loops of the untouched reference, saturated
    and not; the reference window is wider than N (only its first N columns count)."""
    from tests.conftest import tracking_case
    n_act = 0
    for c in map(tracking_case, tracking):
        Ah, Bh = c["A"] + c["dA"], c["B"] + c["dB"]
        u0, V, act = o.mpc_solve(c["N"], Ah, Bh, c["Q"], c["R"], c["Q"], c["lo"], c["hi"], c["x0"],
                                 x_ref=c["x_ref"], u_ref=c["u_ref"])
        n_act += int(act)
        assert abs(V - c["V_N"]) < TOL * abs(c["V_N"]) and np.max(np.abs(u0 - c["u_0"])) < 1e-10
        sim = o.simulate(c["T"], c["N"], Ah, Bh, c["Q"], c["R"], c["Q"], c["lo"], c["hi"], c["x0"], c["A"], c["B"],
                         x_ref=c["x_ref"], u_ref=c["u_ref"])
        assert abs(sim["J_T"] - c["J_T"]) < TOL * abs(c["J_T"])
        assert np.max(np.abs(sim["U"] - c["U"])) < 1e-10 and np.max(np.abs(sim["X"] - c["X"])) < 1e-10
    assert 0 < n_act < len(tracking)


def test_dense_inequality_qp_oracle():
    """oracle/qp_dense.py (general polytopes): equals the exact box solver on boxes and SLSQP (loosely) otherwise."""
    from scipy.optimize import minimize
    from oracle.qp_dense import ineq_qp, polytope_vertices
    rng = np.random.default_rng(0)
    for _ in range(40):
        nz = int(rng.integers(2, 12))
        M = rng.normal(size=(nz, nz)); H = M @ M.T + 0.1 * np.eye(nz); g = rng.normal(size=nz) * 3
        lo, hi = -rng.uniform(0.1, 1, nz), rng.uniform(0.1, 1, nz)
        z, _ = ineq_qp(H, g, np.vstack((np.diag(1 / hi), np.diag(1 / lo))), np.ones(2 * nz))
        assert np.max(np.abs(z - o.box_qp(H, g, lo, hi))) < 1e-10
    for _ in range(10):
        nz = 6
        M = rng.normal(size=(nz, nz)); H = M @ M.T + 0.1 * np.eye(nz); g = rng.normal(size=nz) * 3
        C = rng.normal(size=(14, nz)); d = np.ones(14)
        z, W = ineq_qp(H, g, C, d)
        r = minimize(lambda x: x @ H @ x + 2 * g @ x, np.zeros(nz), jac=lambda x: 2 * H @ x + 2 * g, method="SLSQP",
                     constraints=[{"type": "ineq", "fun": lambda x: d - C @ x, "jac": lambda x: -C}],
                     options={"ftol": 1e-14, "maxiter": 500})
        assert np.max(np.abs(z - r.x)) < 1e-5 and np.all(C @ z <= 1 + 1e-10)
    V = polytope_vertices(np.array([[1, 1], [-1, 1], [0, -2.0]]))
    assert sorted(map(tuple, np.round(V, 12))) == [(-1.5, -0.5), (0.0, 1.0), (1.5, -0.5)]


def test_oracle_general_polytope_cases(polytope):
    """General F_u (rows coupling the inputs, utils_class.py:81): 16 solves / closed loops / energy_bound calls of the
    untouched reference (its QPs solved by the dense active-set shim), half of them constraint-active."""
    from tests.conftest import polytope_case
    n_act = 0
    for c in map(polytope_case, polytope):
        Ah, Bh = c["A"] + c["dA"], c["B"] + c["dB"]
        u0, V, act = o.mpc_solve(c["N"], Ah, Bh, c["Q"], c["R"], c["Q"], None, None, c["x0"], F_u=c["F_u"])
        assert abs(V - c["V_N"]) < TOL * abs(c["V_N"]) and np.max(np.abs(u0 - c["u_0"])) < 1e-10
        sim = o.simulate(c["T"], c["N"], Ah, Bh, c["Q"], c["R"], c["Q"], None, None, c["x0"], c["A"], c["B"],
                         F_u=c["F_u"])
        n_act += sim["n_active"] > 0
        assert abs(sim["J_T"] - c["J_T"]) < TOL * abs(c["J_T"]) and np.max(np.abs(sim["U"] - c["U"])) < 1e-10
        bu, bdu = o.bar_u_poly(c["F_u"])
        assert abs(bu - c["bar_u"]) < 1e-12 * bu and abs(bdu - c["bar_d_u"]) < 1e-12 * bdu
        bnd = o.energy_bound(Ah, Bh, c["Q"], c["R"], None, None, c["N"], c["e"], c["e"], c["x0"], [0.1, 1, 0.6],
                             F_u=c["F_u"])
        assert abs(bnd["alpha"] - c["alpha"]) < TOL * c["alpha"] and abs(bnd["beta"] - c["beta"]) < TOL * c["beta"]
    assert 4 <= n_act < len(polytope)


# ------------------------------------------------------------------------------------------- analytic cross-checks
def test_oracle_extension_variant(extension_cases):
    """energy_decreasing_extension / fc_omega_eta_extension (utils_class.py:375-406, utils.py:412-466): the oracle
    restatement vs the untouched reference, including the cases where the reference raises (math.log of a
    non-positive number)."""
    n_raise = n_ok = 0
    for c in extension_cases:
        n = c["n"]
        A, B, K, hatK = (np.array(c[k]) for k in ("A", "B", "K", "hatK"))
        Q, R = c["q"] * np.eye(n), c["r"] * np.eye(1)
        lo, hi = np.array([-c["ub"]]), np.array([c["ub"]])
        if "raises" in c:
            with pytest.raises(ValueError):
                o.energy_decreasing_extension(A, B, Q, R, lo, hi, c["N"], c["e"], c["e"], K, hatK, c["M_V"])
            n_raise += 1
            continue
        oe = o.fc_omega_eta_extension(c["N"], A, B, Q, R, K, hatK, c["L_V"], c["N_0"])
        for f in ("omega_N1", "omega_N0d5", "eta", "err_th", "N_min"):
            assert abs(oe[f] - c["omega_eta"][f]) <= TOL * abs(c["omega_eta"][f]), f
        dec = o.energy_decreasing_extension(A, B, Q, R, lo, hi, c["N"], c["e"], c["e"], K, hatK, c["M_V"])
        assert abs(dec["xi"] - c["xi"]) <= TOL * abs(c["xi"]) and abs(dec["eta"] - c["eta"]) <= TOL * abs(c["eta"])
        n_ok += 1
    assert n_ok >= 8 and n_raise >= 1


def test_oracle_cfg4_samples_vs_untouched_reference(cfg4):
    """BASELINE configs[3] shape (n = 4, m = 2, N = 10): the restatement AND the batched K1 oracle vs the untouched
    LQ_MPC_Controller / LQ_MPC_Simulator — first-step gain, V_N, J_T over T = 400 steps (= J_inf to rounding)."""
    A, B = np.array(cfg4["A"]), np.array(cfg4["B"])
    n, m, N, T = cfg4["n"], cfg4["m"], cfg4["N"], cfg4["T"]
    Q, R = np.eye(n), np.eye(m)
    cs = cfg4["cases"]
    dA = np.array([c["dA"] for c in cs]); dB = np.array([c["dB"] for c in cs]); x0 = np.array([c["x0"] for c in cs])
    Pexp = nb.expert_matrix(A, B, Q, R, Q, 30)
    bat = nb.eval_batch(A, B, Q, R, Q, Pexp, dA, dB, x0, N, N, T=T, want_K=True)
    for s, c in enumerate(cs):
        K = o.riccati(A + dA[s], B + dB[s], Q, R, Q, N)[0][0]
        assert np.max(np.abs(K - np.array(c["K0"]))) < 1e-10 * max(1.0, np.max(np.abs(K)))
        assert np.max(np.abs(bat["K0"][0, s] - np.array(c["K0"]))) < 1e-10 * max(1.0, np.max(np.abs(K)))
        J_inf, rho = o.closed_loop_inf_cost(A, B, K, Q, R, x0[s])
        assert rho < 1 and abs(J_inf - c["J_T"]) <= TOL * c["J_T"]
        assert abs(bat["J"][0, s] - c["J_T"]) <= TOL * c["J_T"] and abs(bat["JT"][0, s] - c["J_T"]) <= TOL * c["J_T"]
        assert abs(bat["Vn"][0, s] - c["V_N"]) <= TOL * c["V_N"]
        assert np.max(np.abs(K @ x0[s] - np.array(c["u_0"]))) < 1e-10


def test_riccati_equals_condensed_qp_when_unconstrained():
    """utils_class.py:59-91 unconstrained == Riccati: u_0 = K_0 x0, V_N = x0' P_0 x0."""
    rng = np.random.default_rng(0)
    for n, m, N in [(2, 1, 7), (4, 2, 10), (3, 3, 5), (1, 1, 1)]:
        A = rng.normal(size=(n, n)); B = rng.normal(size=(n, m))
        A *= 1.1 / np.max(np.abs(np.linalg.eigvals(A)))      # mildly unstable plant, well-conditioned condensed Hessian
        Mq = rng.normal(size=(n, n)); Q = Mq @ Mq.T + np.eye(n); R = np.eye(m) * 0.7
        x0 = rng.normal(size=n)
        Ks, Ps = o.riccati(A, B, Q, R, Q, N)
        H, g, c0 = o.condensed_qp(N, A, B, Q, R, Q, x0)
        z = np.linalg.solve(H, -g)
        V = z @ H @ z + 2 * g @ z + c0 + x0 @ Q @ x0
        assert np.max(np.abs(z[:m] - Ks[0] @ x0)) < 1e-10 * max(1, np.max(np.abs(z)))
        assert abs(V - x0 @ Ps[0] @ x0) < 1e-10 * abs(V)


def test_lyapunov_doubling_vs_scipy_and_long_rollout():
    """J_inf = x0' S x0 with S from squared doubling == scipy's Lyapunov solve == the T -> inf limit of J_T."""
    A, B, Q, R = nb.synth_problem(4, 2, seed=0)
    dA, dB, x0 = nb.synth_samples(4, 2, 16, seed=2)
    Pexp = nb.expert_matrix(A, B, Q, R, Q, 30)
    out = nb.eval_batch(A, B, Q, R, Q, Pexp, dA, dB, x0, 10, 10, T=400, want_K=True)
    for s in range(16):
        K = out["K0"][0, s]
        Acl = A + B @ K
        S = sla.solve_discrete_lyapunov(Acl.T, Q + K.T @ R @ K)
        assert abs(out["J"][0, s] - x0[s] @ S @ x0[s]) < 1e-11 * out["J"][0, s]
        assert abs(out["JT"][0, s] - out["J"][0, s]) < 1e-11 * out["J"][0, s]
        assert abs(out["rho"][0, s] - np.max(np.abs(np.linalg.eigvals(Acl)))) < 1e-12
        Jr, rr = o.closed_loop_inf_cost(A, B, K, Q, R, x0[s])
        assert abs(Jr - out["J"][0, s]) < 1e-11 * Jr and abs(rr - out["rho"][0, s]) < 1e-12


def test_scalar_system_closed_forms():
    """BASELINE config 0 calls the example "1-D": scalar plant, everything in closed form."""
    a, b, q, r = 1.2, 0.8, 2.0, 0.5
    A, B, Q, R = np.array([[a]]), np.array([[b]]), np.array([[q]]), np.array([[r]])
    p = q
    for N in range(1, 8):
        k = -(b * p * a) / (r + b * b * p)
        p_next = q + a * a * p - (a * b * p) ** 2 / (r + b * b * p)
        Ks, Ps = o.riccati(A, B, Q, R, Q, N)
        assert abs(Ks[0][0, 0] - k) < 1e-14 and abs(Ps[0][0, 0] - p_next) < 1e-13
        acl = a + b * k
        J, rho = o.closed_loop_inf_cost(A, B, Ks[0], Q, R, np.array([0.3]))
        assert abs(rho - abs(acl)) < 1e-15
        assert abs(J - 0.09 * (q + r * k * k) / (1 - acl * acl)) < 1e-13
        p = p_next
    # DARE fixed point: p = q + a^2 p - (a b p)^2/(r + b^2 p)
    K, P = o.dlqr(A, B, Q, R)
    pp = P[0, 0]
    assert abs(pp - (q + a * a * pp - (a * b * pp) ** 2 / (r + b * b * pp))) < 1e-12
    assert abs(K[0, 0] - a * b * pp / (r + b * b * pp)) < 1e-13


def test_batched_oracle_equals_per_sample_oracle():
    """np_batched (the CPU baseline of bench.py) against the per-sample restatement, incl. an unstable sample."""
    for n, m in [(4, 2), (2, 1), (3, 2)]:
        A, B, Q, R = nb.synth_problem(n, m, seed=1)
        dA, dB, x0 = nb.synth_samples(n, m, 24, seed=3, e=0.3)
        Pexp = nb.expert_matrix(A, B, Q, R, Q, 30)
        Ks, Ps = o.riccati(A, B, Q, R, Q, 30)
        assert relerr(Pexp, Ps[0]) < 1e-13
        out = nb.eval_batch(A, B, Q, R, Q, Pexp, dA, dB, x0, 1, 5, T=12)
        for s in range(24):
            for N in (1, 3, 5):
                K = o.riccati(A + dA[s], B + dB[s], Q, R, Q, N)[0][0]
                J, rho = o.closed_loop_inf_cost(A, B, K, Q, R, x0[s])
                assert abs(out["rho"][N - 1, s] - rho) < 1e-11 * rho
                assert bool(out["unstable"][N - 1, s]) == (rho >= 1.0)
                if math.isinf(J):
                    assert math.isinf(out["J"][N - 1, s])
                elif abs(rho - 1) > 1e-3:
                    assert abs(out["J"][N - 1, s] - J) < TOL * J
                sim = o.simulate(12, N, A + dA[s], B + dB[s], Q, R, Q, np.full(m, -np.inf), np.full(m, np.inf),
                                 x0[s], A, B)
                assert abs(out["JT"][N - 1, s] - sim["J_T"]) < TOL * sim["J_T"]


def test_column_stats_oracle(golden):
    st = o.column_stats(golden["true_cost_error"])
    assert np.array_equal(st["max"], golden["true_cost_error"].max(axis=0))
    assert np.allclose(st["std"], np.std(golden["true_cost_error"], axis=0), rtol=0, atol=0)


# ------------------------------------------------------------------------------- the reference itself (container only)
@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference is mounted in the build container only")
def test_untouched_reference_behind_shims_matches_golden_and_oracle(golden, example):
    """Re-runs the UNMODIFIED reference classes (oracle/ref_oracle.py) on a slice: shipped tables, oracle, and the
    committed known answers all agree — this is what pins the oracle for the GPU box, where the reference is absent."""
    from oracle import ref_oracle as ro
    u, uc = ro.load()
    cfg = ro.example_multiple_config()
    with ro.reference_cwd():
        beh = uc.LQ_RDP_Behavior_Multiple(cfg["info_opc"], cfg["info_N"], cfg["info_e"], cfg["N_matrix"], "f")
    assert np.array_equal(beh.error_A, golden["error_A_f"]) and np.array_equal(beh.error_B, golden["error_B_f"])
    A, B, Q, R, F_u = (example[x] for x in ("A", "B", "Q", "R", "F_u"))
    x0_vec = u.circle_generator(8, 1.5, beh.epsilon_lqr, Q)
    x_start = x0_vec[:, 1]
    for (j, i) in [(0, 0), (7, 9), (50, 4)]:
        Ah = A + golden["error_A_f"][:, :, j, i]; Bh = B + golden["error_B_f"][:, :, j, i]
        sim = uc.LQ_MPC_Simulator(30, 7, Ah, Bh, Q, R, Q, F_u).simulate(x_start, A, B, np.zeros((2, 30)),
                                                                        np.zeros((1, 30)))
        assert abs(sim["J_T"] - golden["true_cost_error"][j, i]) < TOL * sim["J_T"]
        mine = o.simulate(30, 7, Ah, Bh, Q, R, Q, example["lo"], example["hi"], x_start, A, B)
        assert abs(sim["J_T"] - mine["J_T"]) < 1e-12 * sim["J_T"]
        calc = uc.LQ_RDP_Calculator(Ah, Bh, Q, R, F_u)
        e = float(golden["error"][i])
        bnd = calc.energy_bound(7, e, e, x_start, example["p"])
        assert abs(bnd["alpha"] - golden["alpha_table_error"][j, i]) < TOL * bnd["alpha"]
        assert abs(bnd["beta"] - golden["beta_table_error"][j, i]) < TOL * bnd["beta"]
