"""CPU tier — the host-side half of the drop-in layer (no engine call): scalar formulas, matrix builders, input-set
constants, circle / sample generators and layout packing, against the oracle, the untouched reference's answers
(tests/golden/ref_known_answers.json) and — in the build container — the reference module itself."""
import os

import numpy as np
import pytest

from lq_mpc_b200 import runtime as rt
from lq_mpc_b200 import sampling as sp
from lq_mpc_b200 import utils as U
from oracle import np_oracle as o


def test_scalar_error_functions_vs_oracle():
    for (p, i, eA, fA, eB, fB) in [(2, 3, 0.01, 1.2, 0.02, 1.5), (1, 0, 0.1, 0.9, 0.0, 2.0), (2, 7, 1e-3, 1.0, 1e-3, 1.0)]:
        assert U.fc_ec_g_x(p, i, eA, fA) == o.g_x(p, i, eA, fA)
        assert U.fc_ec_g_u(p, i, eA, fA, eB, fB) == o.g_u(p, i, eA, fA, eB, fB)
    for N in (1, 6, 30):
        assert abs(U.fc_ec_bar_g_x(N, 0.01, 1.3) - o.bar_g_x(N, 0.01, 1.3)) <= 1e-15 * abs(o.bar_g_x(N, 0.01, 1.3))
        ref = o.bar_g_u(N, 0.01, 1.3, 0.02, 1.7)
        assert abs(U.fc_ec_bar_g_u(N, 0.01, 1.3, 0.02, 1.7) - ref) <= 1e-15 * abs(ref)
    assert U.ex_stability_bounds(2.2418, 0.045, 0.2023) == o.ex_stability_bounds(2.2418, 0.045, 0.2023)
    assert U.ex_stability_bounds(5.0, 1.0, 2.0) == {'L_V': 5.0, 'N_0': 0}


def test_stacked_matrix_builders_vs_oracle():
    rng = np.random.default_rng(0)
    for n, m, N in [(2, 1, 6), (4, 2, 10), (3, 3, 1)]:
        A, B = rng.normal(size=(n, n)), rng.normal(size=(n, m))
        assert np.allclose(U.sl_syn_Phi(N, A), o.sl_syn_Phi(N, A), rtol=1e-14, atol=0)
        G = U.sl_syn_Gamma(N, A, B)
        assert G.shape == ((N + 1) * n, N * m)
        assert np.allclose(G, o.sl_syn_Gamma(N, A, B), rtol=1e-14, atol=1e-300)
        # meaning: stacked states of x+ = A x + B u from x0 are Phi x0 + Gamma u
        x0, u = rng.normal(size=n), rng.normal(size=(N, m))
        xs = [x0]
        for t in range(N):
            xs.append(A @ xs[-1] + B @ u[t])
        assert np.allclose(np.concatenate(xs), U.sl_syn_Phi(N, A) @ x0 + G @ u.ravel(), rtol=1e-12, atol=1e-12)


def test_input_set_constants_and_box_parsing(known):
    F_u = np.array([[10.0], [-10.0]])
    assert abs(U.bar_u_solve(F_u) - known["single"]["bar_u"]) < 1e-15
    assert abs(U.bar_d_u_solve(F_u) - known["single"]["bar_d_u"]) < 1e-15
    F2 = np.vstack((np.diag([2.0, 4.0]), -np.diag([1.0, 5.0])))
    lo, hi = rt.box_from_F(F2)
    assert np.allclose(lo, [-1.0, -0.2]) and np.allclose(hi, [0.5, 0.25])
    assert abs(U.bar_u_solve(F2) - (1.0 + 0.0625)) < 1e-15
    assert abs(U.bar_d_u_solve(F2) - (1.5 ** 2 + 0.45 ** 2)) < 1e-15
    with pytest.raises(NotImplementedError):
        rt.box_from_F(np.array([[1.0, 1.0]]))
    # general polytopes: vertex enumeration stands in for the two Gurobi models (utils.py:592-650)
    tri = np.array([[1, 1], [-1, 1], [0, -2.0]])
    assert not rt.is_box(tri) and rt.is_box(F2)
    V = rt.polytope_vertices(tri)
    assert sorted(map(tuple, np.round(V, 12))) == [(-1.5, -0.5), (0.0, 1.0), (1.5, -0.5)]
    assert abs(U.bar_u_solve(tri) - 2.5) < 1e-15 and abs(U.bar_d_u_solve(tri) - 9.0) < 1e-15
    with pytest.raises(ValueError):
        rt.polytope_vertices(np.array([[1.0, 1.0], [-1.0, 1.0]]))        # a cone: unbounded
    with pytest.raises(ValueError):
        U.bar_u_solve(np.array([[10.0]]))                 # unbounded below (Gurobi would report unbounded)


def test_numa_binding_is_optional():
    """bind_to_gpu_numa never raises: without NVML / a GPU it returns None and leaves the affinity mask alone."""
    import os
    before = os.sched_getaffinity(0)
    r = rt.bind_to_gpu_numa(0)
    assert r is None or set(r) <= before
    os.sched_setaffinity(0, before)


def test_vertex_constants_vs_reference_answers(polytope):
    """bar_u_solve / bar_d_u_solve on general polytopes vs the untouched reference (Gurobi models replaced by vertex
    enumeration in the oracle's shim; both are maxima of convex functions over the same vertex set)."""
    for c in polytope:
        F = np.array(c["F_u"])
        assert abs(U.bar_u_solve(F) - c["bar_u"]) < 1e-12 * c["bar_u"]
        assert abs(U.bar_d_u_solve(F) - c["bar_d_u"]) < 1e-12 * c["bar_d_u"]


def test_circle_generator_vs_reference_answers(known):
    k = known["single"]
    x0_vec = U.circle_generator(8, 1.5, k["epsilon_lqr"], 2 * np.eye(2))
    assert x0_vec.shape == (2, 8) and np.max(np.abs(x0_vec - np.array(k["x0_vec"]))) < 1e-15
    # every point sits on the Q-ellipse x'Qx = (1.5)^2 eps
    assert np.allclose(2 * np.sum(x0_vec ** 2, axis=0), 2.25 * k["epsilon_lqr"], rtol=1e-13)
    Qn = np.array([[2.0, 0.3], [0.3, 1.0]])            # non-diagonal Q: the literal cho_factor quirk (utils.py:697-699)
    assert np.allclose(U.circle_generator(5, 1.1, 0.3, Qn), o.circle_generator(5, 1.1, 0.3, Qn), rtol=1e-14)
    assert np.allclose(U.rot_2D(0.3) @ U.rot_2D(-0.3), np.eye(2), atol=1e-16)


def test_random_matrix_semantics():
    """utils.py:779-823: 5*N matrices, all ||.|| <= bound, the first N on the boundary; 'f' and '2' norms; other
    norm types leave zeros (as the reference does)."""
    rng = np.random.default_rng(3)
    for nt, ordv in (("f", "fro"), ("2", 2)):
        out = sp.reference_random_matrix(np.zeros((2, 2)), 6, 0.01, nt, rng)
        assert out.shape == (2, 2, 30)
        nv = np.array([np.linalg.norm(out[:, :, k], ord=ordv) for k in range(30)])
        assert np.all(nv <= 0.01 * (1 + 1e-12)) and np.allclose(nv[:6], 0.01, rtol=1e-12)
        assert np.all(np.abs(out) <= 0.01)
        assert nv[6:].std() > 0
    assert not sp.reference_random_matrix(np.zeros((2, 1)), 3, 0.01, "inf").any()
    lit = sp.reference_random_matrix(np.zeros((1, 1)), 2, 0.5, "f", np.random.default_rng(0), literal=False)
    assert np.allclose(np.abs(lit[0, 0, :2]), 0.5)


def test_error_grid_files_and_soa_packing(tmp_path, golden):
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        out = U.error_matrix_generator(np.zeros((2, 2)), np.zeros((2, 1)), np.linspace(1e-3, 1e-2, 4), 2, "f")
        assert sorted(os.listdir(tmp_path)) == ["error_A_f.npy", "error_B_f.npy"]     # reference file names (utils.py:844)
        assert np.array_equal(np.load("error_A_f.npy"), out["error_A"])
    finally:
        os.chdir(cwd)
    assert out["error_A"].shape == (2, 2, 10, 4) and out["error_B"].shape == (2, 1, 10, 4)
    eA, eB = golden["error_A_f"], golden["error_B_f"]
    sA, sB = sp.grids_to_soa(eA, eB)
    assert sA.shape == (4, 1000) and sB.shape == (2, 1000)
    j, i = 37, 6
    assert np.array_equal(sA[:, j * 10 + i], eA[:, :, j, i].ravel())
    assert np.array_equal(sB[:, j * 10 + i], eB[:, :, j, i].ravel())
    lA, lB = sp.grids_to_soa(eA, eB, level=4)
    assert np.array_equal(lA[:, 12], eA[:, :, 12, 4].ravel()) and lB.shape == (2, 100)
    gA, gB = sp.seeded_error_grids(2, 1, np.linspace(1e-3, 1e-2, 3), 4, "2", seed=9)
    hA, hB = sp.seeded_error_grids(2, 1, np.linspace(1e-3, 1e-2, 3), 4, "2", seed=9)
    assert np.array_equal(gA, hA) and np.array_equal(gB, hB) and gA.shape == (2, 2, 20, 3)


def test_find_closest_index_and_colors():
    a = np.array([1e-3, 2e-3, 5e-3, 1e-2])
    assert U.find_closest_index(a, 5e-3) == 2 and U.find_closest_index(a, 0.0) == 0
    assert U.find_closest_index(a, 1.0) == 3 and U.find_closest_index(a, 2.4e-3) == 1
    c = U.default_color_generator()
    assert len(c) == 10 and c["C0"] == (31 / 255, 119 / 255, 180 / 255)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference is mounted in the build container only")
def test_host_functions_vs_reference_module():
    from oracle import ref_oracle as ro
    u, _ = ro.load()
    rng = np.random.default_rng(1)
    A, B = rng.normal(size=(2, 2)), rng.normal(size=(2, 1))
    assert np.allclose(U.sl_syn_Phi(5, A), u.sl_syn_Phi(5, A), rtol=1e-14)
    assert np.allclose(U.sl_syn_Gamma(5, A, B), u.sl_syn_Gamma(5, A, B), rtol=1e-14)
    assert U.fc_ec_g_u(2, 3, 0.01, 1.2, 0.02, 1.5) == u.fc_ec_g_u(2, 3, 0.01, 1.2, 0.02, 1.5)
    assert abs(U.fc_ec_bar_g_u(6, 0.01, 1.2, 0.02, 1.5) - u.fc_ec_bar_g_u(6, 0.01, 1.2, 0.02, 1.5)) < 1e-15
    assert abs(U.fc_ec_bar_g_x(6, 0.01, 1.2) - u.fc_ec_bar_g_x(6, 0.01, 1.2)) < 1e-16
    Q = np.array([[2.0, 0.0], [0.0, 2.0]])
    assert np.allclose(U.circle_generator(8, 1.5, 0.045, Q), u.circle_generator(8, 1.5, 0.045, Q), rtol=1e-15)
    assert U.ex_stability_bounds(2.24, 0.045, 0.2) == u.ex_stability_bounds(2.24, 0.045, 0.2)
    av = np.array([1e-3, 2e-3, 5e-3])
    assert U.find_closest_index(av, 1.9e-3) == u.find_closest_index(av, 1.9e-3)
