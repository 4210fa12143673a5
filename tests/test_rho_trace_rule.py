"""CPU check of the RULE behind the early spectral-radius acceptance of k_group.cu / k_tiled.cu (trace-based dominant
pair; numpy restatement in scripts/rho_trace_study.py — the kernels themselves are checked on the GPU): whatever the
rule accepts must be within 1e-9 of LAPACK on well-conditioned spectra, defective spectra must never be accepted, and
typical closed loops must be accepted after a few squarings (that is the point of the rule)."""
import importlib.util
import os

import numpy as np

_spec = importlib.util.spec_from_file_location(
    "rho_trace_study", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "rho_trace_study.py"))
study = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(study)


def test_accepted_candidates_are_accurate_and_early():
    rng = np.random.default_rng(3)
    ks = []
    for n in (6, 8, 16, 32):
        for _ in range(150):
            M = rng.standard_normal((n, n)) * rng.uniform(0.1, 0.5)
            ref = np.abs(np.linalg.eigvals(M)).max()
            r, k, ok = study.rho_trace(M)
            assert ok, "a random matrix with a simple dominant eigenvalue / pair must be accepted"
            assert abs(r - ref) <= 1e-9 * ref
            ks.append(k)
    assert np.median(ks) <= 14 and max(ks) <= 30


def test_special_spectra():
    n = 8
    rng = np.random.default_rng(4)
    for _ in range(100):                                 # defective dominant eigenvalue: never accepted
        V = rng.standard_normal((n, n)); D = np.diag(rng.uniform(-0.5, 0.5, n))
        D[0, 0] = D[1, 1] = D[2, 2] = 0.9; D[0, 1] = D[1, 2] = 1.0
        assert not study.rho_trace(V @ D @ np.linalg.inv(V))[2]
    r, k, ok = study.rho_trace(np.zeros((n, n)))         # zero / nilpotent: rho = 0 from the vanishing power
    assert ok and r == 0.0
    N = np.diag(np.ones(n - 1), 1)
    r, k, ok = study.rho_trace(N)
    assert ok and r == 0.0 and k <= 4
    for th in (0.3, 1.0, 2.5):                           # rotation-dominated: complex pair, exact answer known
        M = np.zeros((n, n)); c, s = 0.95 * np.cos(th), 0.95 * np.sin(th)
        M[:2, :2] = [[c, s], [-s, c]]; M[2:, 2:] = np.diag(rng.uniform(-0.5, 0.5, n - 2))
        Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        r, k, ok = study.rho_trace(Q @ M @ Q.T)
        assert ok and abs(r - 0.95) <= 1e-12 and k <= 12
    M = np.diag([0.9, -0.9, 0.3, 0.2, 0.1, 0.0, -0.1, -0.2])    # exact +- pair: double root of the power model,
    r, k, ok = study.rho_trace(M)                                # the guard refuses it (falls through to the norm test)
    assert not ok
