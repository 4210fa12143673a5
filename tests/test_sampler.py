"""Model-error sampler (K6): Philox4x32-10 known answers, the numpy restatement (oracle/np_sampler.py) vs the
engine's own sampler core compiled for the host (CPU tier) and vs the device kernel (GPU tier), and the semantics of
the reference's `random_matrix` (utils.py:779-823)."""
import numpy as np
import pytest

from oracle import np_sampler as ns

# Random123 kat_vectors, philox4x32-10: (counter, key) -> output
KAT = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
       ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
       ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
LEVELS = np.linspace(1e-3, 1e-2, 10)


def test_philox_known_answers_oracle_and_engine_core():
    hm = pytest.importorskip("tests.hostmath.api")
    for ctr, key, out in KAT:
        got = ns.philox4x32_10(*ctr, *key)
        assert [int(x) for x in got] == list(out)
        assert [int(x) for x in hm.philox4x32_10(ctr, key)] == list(out)


def test_uniform_construction_is_exact_and_in_range():
    w = np.array([0, 0xffffffff, 0x80000000, 123456789], dtype=np.uint32)
    u = ns.symm(w, w[::-1].copy())
    assert np.all(u >= -1.0) and np.all(u < 1.0)
    assert ns.symm(np.uint32([0]), np.uint32([0]))[0] == -1.0
    assert ns.symm(np.uint32([0xffffffff]), np.uint32([0xffffffff]))[0] == 1.0 - 2.0 ** -52


@pytest.mark.parametrize("rows,cols,norm", [(2, 2, "f"), (2, 1, "f"), (2, 2, "2"), (2, 1, "2"), (4, 2, "2")])
def test_engine_sampler_core_matches_oracle_bit_for_bit(rows, cols, norm):
    hm = pytest.importorskip("tests.hostmath.api")
    N_sys, nb = 400, 80
    ref, rej_r, proj_r = ns.sample_error_grid(2024, 1 if cols != rows else 0, rows, cols, N_sys, LEVELS, nb, norm)
    got, rej_g, proj_g = hm.sample_error_grid(2024, 1 if cols != rows else 0, rows, cols, N_sys, LEVELS, nb, norm)
    assert ref.shape == (rows, cols, N_sys, 10) and proj_r == proj_g == 0
    # interior samples: identical accept decisions => identical bits; boundary ones carry one rescaling (1 ulp)
    assert np.array_equal(ref[:, :, nb:, :], got[:, :, nb:, :])
    # (the spectral norm comes from two different algorithms: a few ulp)
    assert np.max(np.abs(ref[:, :, :nb, :] - got[:, :, :nb, :])) <= 4e-15 * LEVELS.max()
    assert rej_r == rej_g


def test_sampler_semantics_and_shard_invariance():
    g, rej, proj = ns.sample_error_grid(7, 0, 2, 2, 300, LEVELS, 60, "f")
    nv = np.sqrt((g ** 2).sum(axis=(0, 1)))                              # (N_sys, n_err)
    assert np.allclose(nv[:60], LEVELS[None, :], rtol=1e-14)             # boundary block (utils.py:803)
    assert np.all(nv[60:] <= LEVELS[None, :]) and np.all(np.abs(g) <= LEVELS[None, None, None, :])
    assert proj == 0 and 0.5 < (240 * 10) / (240 * 10 + rej) < 0.9 or rej > 0   # acceptance of the 4-ball in the cube ~ 0.31
    g2, _, _ = ns.sample_error_grid(7, 0, 2, 2, 300, LEVELS, 60, "2")
    n2 = np.linalg.norm(np.moveaxis(g2, (0, 1), (2, 3)), ord=2, axis=(2, 3))
    assert np.allclose(n2[:60], LEVELS[None, :], rtol=1e-13) and np.all(n2[60:] <= LEVELS[None, :] * (1 + 1e-15))
    part, _, _ = ns.sample_error_grid(7, 0, 2, 2, 100, LEVELS, 60, "f", j_first=150)
    assert np.array_equal(part, g[:, :, 150:250, :])                     # counter = global perturbation index
    gB, _, _ = ns.sample_error_grid(7, 1, 2, 1, 300, LEVELS, 60, "f")
    assert not np.array_equal(gB[0, 0], g[0, 0])                         # `which` separates the A and B streams
    # mean ~ 0, entries spread over the level's range
    assert abs(g[:, :, 60:, 9].mean()) < 2e-4 and g[:, :, 60:, 9].std() > 2e-3


def test_seeded_evaluation_samples_core_matches_oracle():
    """The generator behind lqmpc_eval_seeded (csrc/sampler.cuh: seeded_sample, compiled here with g++) vs its numpy
    restatement: uniforms bit-exact, Box-Muller normals to rounding; N(0, I) / U[-e, e) moments; counter = global
    sample index (a shard starting at `first` reproduces the corresponding slice)."""
    from tests.hostmath import api as hm
    for n, m in ((4, 2), (2, 1), (3, 3), (1, 1)):
        dA, dB, x0 = hm.seeded_samples(77, 5, 4096, n, m, 0.01, 0.02)
        rA, rB, rx = ns.seeded_samples(77, 5, 4096, n, m, 0.01, 0.02)
        assert np.array_equal(dA, rA) and np.array_equal(dB, rB)
        assert np.max(np.abs(x0 - rx)) < 1e-14
        assert np.max(np.abs(dA)) <= 0.01 and np.max(np.abs(dB)) <= 0.02
    a_all = ns.seeded_samples(3, 0, 3000, 4, 2, 0.01, 0.01)
    a_mid = ns.seeded_samples(3, 1000, 500, 4, 2, 0.01, 0.01)
    for u, v in zip(a_all, a_mid):
        assert np.array_equal(u[1000:1500], v)
    big = ns.seeded_samples(11, 1 << 33, 200_000, 4, 2, 0.01, 0.01)            # indices beyond 2^32
    assert abs(big[2].mean()) < 5e-3 and abs(big[2].std() - 1.0) < 5e-3
    assert abs(big[0].mean()) < 1e-4 and abs(big[0].std() - 0.01 / np.sqrt(3)) < 1e-4
    other = ns.seeded_samples(12, 1 << 33, 100, 4, 2, 0.01, 0.01)
    assert not np.array_equal(other[0], big[0][:100])


@pytest.mark.gpu
@pytest.mark.parametrize("rows,cols,norm", [(2, 2, "f"), (2, 1, "2"), (4, 4, "2"), (3, 2, "f")])
def test_device_sampler_matches_oracle(engine, rows, cols, norm):
    N_sys, nb = 2003, 401
    which = 0 if rows == cols else 1
    out, st = engine.sample_error_grid(99, which, rows, cols, N_sys, LEVELS, nb, norm, want_stats=True)
    got = out.cpu().numpy().reshape(rows, cols, N_sys, 10)
    ref, rej, proj = ns.sample_error_grid(99, which, rows, cols, N_sys, LEVELS, nb, norm)
    assert st["projected"] == proj and abs(st["rejected"] - rej) <= 2
    if rows * cols <= 6:                                # rejection terminates quickly only for small matrices:
        assert proj == 0                                # a 4 x 4 draw almost never has ||T||_2 <= e -> projected
        same = np.all(ref[:, :, nb:, :] == got[:, :, nb:, :], axis=(0, 1))
        assert same.mean() > 0.9999                      # borderline ||T|| == e decisions may differ by rounding
    else:
        assert proj > 0.9 * (N_sys - nb) * 10
    # rescaled samples (boundary block, projections) carry one division by a norm from two different algorithms
    assert np.max(np.abs(ref - got)) <= 1e-14 * LEVELS.max()
    part = engine.sample_error_grid(99, which, rows, cols, 500, LEVELS, nb, norm, j_first=1000)
    assert np.array_equal(part.cpu().numpy().reshape(rows, cols, 500, 10), got[:, :, 1000:1500, :])
