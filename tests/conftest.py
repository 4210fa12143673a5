import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long CPU test")


@pytest.fixture(scope="session")
def golden():
    """The reference's shipped result file + input grids (tests/golden/multiple_sys.npz)."""
    return dict(np.load(os.path.join(GOLD, "multiple_sys.npz")))


@pytest.fixture(scope="session")
def known():
    """Answers produced by the untouched reference behind shims (oracle/make_golden.py)."""
    with open(os.path.join(GOLD, "ref_known_answers.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def tracking():
    """Non-zero-reference solves / closed loops of the untouched reference (oracle/make_golden.py --tracking)."""
    with open(os.path.join(GOLD, "ref_tracking_cases.json")) as f:
        return json.load(f)


def tracking_case(c):
    """Arrays of one tests/golden/ref_tracking_cases.json entry."""
    n, m = c["n"], c["m"]
    d = {k: np.array(c[k]) for k in ("A", "B", "dA", "dB", "x0", "x_ref", "u_ref", "u_0", "X", "U")}
    d.update(n=n, m=m, N=c["N"], T=c["T"], Q=c["q"] * np.eye(n), R=c["r"] * np.eye(m), lo=-c["ub"] * np.ones(m),
             hi=c["ub"] * np.ones(m), F_u=np.vstack((np.eye(m), -np.eye(m))) / c["ub"], V_N=c["V_N"], J_T=c["J_T"])
    return d


@pytest.fixture(scope="session")
def polytope():
    """General-polytope F_u solves / closed loops / bounds of the untouched reference (make_golden.py --polytope)."""
    with open(os.path.join(GOLD, "ref_polytope_cases.json")) as f:
        return json.load(f)


def polytope_case(c):
    n, m = c["n"], c["m"]
    d = {k: np.array(c[k]) for k in ("A", "B", "dA", "dB", "x0", "F_u", "u_0", "X", "U")}
    d.update(n=n, m=m, p=c["p"], N=c["N"], T=c["T"], Q=c["q"] * np.eye(n), R=c["r"] * np.eye(m), e=c["e"],
             **{k: c[k] for k in ("V_N", "J_T", "bar_u", "bar_d_u", "alpha", "beta")})
    return d


def random_polytope(rng, m, p, lo=0.1, hi=0.45):
    """Bounded polytope {u : F u <= 1} around the origin: p unit normals that positively span R^m / support distances."""
    while True:
        D = rng.normal(size=(p, m))
        D /= np.linalg.norm(D, axis=1, keepdims=True)
        probe = rng.normal(size=(4000, m))
        probe /= np.linalg.norm(probe, axis=1, keepdims=True)
        if np.min(np.max(probe @ D.T, axis=1)) > 0.1:
            return D / rng.uniform(lo, hi, size=(p, 1))


@pytest.fixture(scope="session")
def golden_norm2():
    return dict(np.load(os.path.join(GOLD, "ref_norm2_subset.npz")))


@pytest.fixture(scope="session")
def example():
    """Constants of working_example_single.py:20-27 / working_example_multiple.py:13-76."""
    A = np.array([[1, 0.7], [0.12, 0.4]])
    B = np.array([[1], [1.2]])
    return dict(A=A, B=B, Q=2 * np.eye(2), R=np.eye(1), lo=np.array([-0.1]), hi=np.array([0.1]),
                F_u=np.array([[10.0], [-10.0]]), p=np.array([0.1, 1, 0.6]),
                info_N={'N_min': 6, 'N_max': 10, 'N_nominal': 7, 'N_opc': 30, 'N_mpc': 30},
                info_e={'e_min': 1e-3, 'e_max': 1e-2, 'e_nominal': 5e-3})


@pytest.fixture(scope="session")
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from lq_mpc_b200.engine import Engine
    return Engine(0)


def relerr(a, b):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin), "finite masks differ"
    if not fin.any():
        return 0.0
    return float(np.max(np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1e-300)))


@pytest.fixture(scope="session")
def extension_cases():
    """energy_decreasing_extension / fc_omega_eta_extension runs of the untouched reference (make_golden.py --round2)."""
    with open(os.path.join(GOLD, "ref_extension_cases.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def cfg4():
    """n = 4, m = 2, N = 10 (BASELINE configs[3]) through the untouched LQ_MPC_Controller / LQ_MPC_Simulator
    (make_golden.py --round2): K0, V_N, J_T(T = 400)."""
    with open(os.path.join(GOLD, "ref_cfg4_cases.json")) as f:
        return json.load(f)
