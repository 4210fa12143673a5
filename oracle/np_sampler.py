"""TEST INFRASTRUCTURE ONLY — numpy restatement of the engine's counter-based model-error sampler (K6,
lq_mpc_b200/csrc/sampler.cuh), i.e. of the reference's `random_matrix` / `error_matrix_generator`
(utils.py:779-847: entries uniform in [-e, e], accepted when ||T|| <= e in the 'f' or '2' norm, the first n_boundary
perturbations of each level on the norm boundary) driven by Philox4x32-10 instead of the unseeded `random.uniform`.

Pinned on the three known-answer vectors of the Random123 distribution (kat_vectors: philox4x32-10) in
tests/test_sampler.py; the integer stream and the uniforms are bit-exact against the device, accept/reject decisions
can differ only for ||T|| within rounding of e.
"""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
STREAM = 0x4C514D50
MAX_ATTEMPTS = 256
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (Salmon et al. 2011). Counters: arrays or scalars (uint32 values); returns 4 uint32 arrays."""
    c = [np.asarray(x, dtype=np.uint64) & _MASK for x in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = np.uint64(k0) & _MASK, np.uint64(k1) & _MASK
    for _ in range(10):
        p0 = np.uint64(M0) * c[0]
        p1 = np.uint64(M1) * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ k0, p1 & _MASK, (p0 >> np.uint64(32)) ^ c[3] ^ k1, p0 & _MASK]
        k0 = (k0 + np.uint64(W0)) & _MASK
        k1 = (k1 + np.uint64(W1)) & _MASK
    return [x.astype(np.uint32) for x in c]


def symm(w0, w1):
    """[-1, 1) from 53 random bits: (k - 2^52) 2^-52, exact."""
    u53 = (w0 >> np.uint32(5)).astype(np.float64) * 67108864.0 + (w1 >> np.uint32(6)).astype(np.float64)
    return (u53 - 4503599627370496.0) * (1.0 / 4503599627370496.0)


def _draw(seed, which, j, level, attempt, rows, cols):
    """Candidate matrices for arrays j, level (same shape) at one attempt number: (len, rows*cols) entries in [-1, 1)."""
    k0 = seed & 0xFFFFFFFF
    k1 = ((seed >> 32) & 0xFFFFFFFF) ^ STREAM
    j = np.asarray(j, dtype=np.uint64)
    c0 = j & _MASK
    c1 = np.asarray(level, dtype=np.uint64) | np.uint64(which << 16) | ((j >> np.uint64(32)) << np.uint64(20))
    out = np.empty((j.size, rows * cols))
    for p in range((rows * cols + 1) // 2):
        x = philox4x32_10(c0, c1, np.uint64(attempt), np.uint64(p), k0, k1)
        out[:, 2 * p] = symm(x[0], x[1])
        if 2 * p + 1 < rows * cols:
            out[:, 2 * p + 1] = symm(x[2], x[3])
    return out


def sample_error_grid(seed, which, rows, cols, N_sys, levels, n_boundary, norm_type="f", j_first=0):
    """Returns (grid, n_rejected, n_projected); grid has the reference layout (rows, cols, N_sys, n_err)."""
    levels = np.asarray(levels, dtype=np.float64)
    n_err = levels.size
    jj, ii = np.meshgrid(np.arange(j_first, j_first + N_sys), np.arange(n_err), indexing="ij")
    jj, ii = jj.ravel(), ii.ravel()                                  # s = j*n_err + i
    e = levels[ii]
    boundary = jj < n_boundary
    T = np.zeros((jj.size, rows * cols))
    pending = np.ones(jj.size, dtype=bool)
    rejected = 0
    nv_last = np.zeros(jj.size)
    for a in range(MAX_ATTEMPTS):
        idx = np.flatnonzero(pending)
        if idx.size == 0:
            break
        cand = _draw(seed, which, jj[idx], ii[idx], a, rows, cols) * e[idx, None]
        M = cand.reshape(-1, rows, cols)
        nv = np.sqrt((cand * cand).sum(axis=1)) if norm_type == "f" else np.linalg.norm(M, ord=2, axis=(1, 2))
        ok = np.where(boundary[idx], nv > 0.0, nv <= e[idx])
        scale = np.where(boundary[idx], e[idx] / np.where(nv > 0, nv, 1.0), 1.0)
        T[idx[ok]] = cand[ok] * scale[ok, None]
        T[idx[~ok]] = cand[~ok]
        nv_last[idx] = nv
        pending[idx[ok]] = False
        rejected += int((~ok).sum())
    proj = np.flatnonzero(pending)
    T[proj] *= (e[proj] / nv_last[proj])[:, None]
    grid = np.ascontiguousarray(T.T).reshape(rows, cols, N_sys, n_err)
    return grid, rejected, int(proj.size)


SEEDED_STREAM = 0x53454544


def seeded_samples(seed, first, S, n, m, e_A, e_B):
    """Restatement of lq::seeded_sample (csrc/sampler.cuh; entry point lqmpc_eval_seeded): global samples
    [first, first + S) -> dA (S, n, n), dB (S, n, m) uniform in [-e, e) (bit-exact against the device), x0 (S, n)
    ~ N(0, I) by Box-Muller (agrees with the device to the last digits of log / cos / sin)."""
    k0 = seed & 0xFFFFFFFF
    k1 = ((seed >> 32) & 0xFFFFFFFF) ^ SEEDED_STREAM
    s = np.arange(first, first + S, dtype=np.uint64)
    c0, c1 = s & _MASK, s >> np.uint64(32)
    nu = n * n + n * m
    flat = np.empty((S, nu))
    for p in range((nu + 1) // 2):
        x = philox4x32_10(c0, c1, np.uint64(p), np.uint64(0), k0, k1)
        flat[:, 2 * p] = symm(x[0], x[1])
        if 2 * p + 1 < nu:
            flat[:, 2 * p + 1] = symm(x[2], x[3])
    dA = (e_A * flat[:, :n * n]).reshape(S, n, n)
    dB = (e_B * flat[:, n * n:]).reshape(S, n, m)
    x0 = np.empty((S, n))
    for p in range((n + 1) // 2):
        x = philox4x32_10(c0, c1, np.uint64((nu + 1) // 2 + p), np.uint64(0), k0, k1)
        u = ((x[0] >> np.uint32(5)).astype(np.float64) * 67108864.0 + (x[1] >> np.uint32(6)).astype(np.float64) + 1.0) \
            * (1.0 / 9007199254740992.0)
        v = ((x[2] >> np.uint32(5)).astype(np.float64) * 67108864.0 + (x[3] >> np.uint32(6)).astype(np.float64)) \
            * (1.0 / 9007199254740992.0)
        r, th = np.sqrt(-2.0 * np.log(u)), 6.283185307179586476925286766559 * v
        x0[:, 2 * p] = r * np.cos(th)
        if 2 * p + 1 < n:
            x0[:, 2 * p + 1] = r * np.sin(th)
    return dA, dB, x0
