"""Development study (CPU, numpy): the trace-based dominant-pair acceptance of the spectral radius used by
k_group.cu (lane-group K1) and k_tiled.cu (K4a), restated in numpy and run against LAPACK on random, adversarial,
defective and LQR closed-loop matrices. This is how kRhoEta, kRhoMinRoot and the near-double-root guard were chosen;
the kernels themselves are tested on the GPU (tests/test_gpu_parity.py: *_hard_spectra, k4 spectral-radius sets).

    python scripts/rho_trace_study.py [n] [trials]

Prints, per family: worst relative error among ACCEPTED samples, how many were accepted, and the squarings needed
(median / p90 / p99 / max). Not accepted = falls through to the 40-squaring two-estimate test and the QR kernel."""
import sys
import numpy as np
import scipy.linalg as sl

ETA, MIN_ROOT, DISC_GUARD, KMAX = 1e-9, 0.02, 1e-4, 40


def rho_trace(M):
    C = M.copy(); lacc = 0.0; wgt = 1.0; prev = None; streak = 0
    for kk in range(KMAX):
        mx = np.abs(C).max()
        if mx == 0 or not np.isfinite(mx):
            return 0.0, kk, mx == 0
        e = int(np.floor(np.log2(mx)))
        X = C * 2.0 ** (-e)                      # max|X| in [1, 2)
        C = X @ X
        lacc += wgt * e
        tau, taup = np.trace(X), np.trace(C)
        d, disc = (tau * tau - taup) / 2, 2 * taup - tau * tau
        r = np.sqrt(d) if disc < 0 else (abs(tau) + np.sqrt(disc)) / 2
        if r > MIN_ROOT and np.isfinite(r) and not abs(disc) < DISC_GUARD * tau * tau:
            good = prev is not None and abs(r - prev * prev * 2.0 ** (-e)) <= ETA * r
            if good and streak >= 1 and ETA * wgt <= 2e-10:
                return float(np.exp(lacc * np.log(2) + wgt * np.log(r))), kk + 1, True
            streak = streak + 1 if good else 0
            prev = r
        else:
            prev = None; streak = 0
        wgt *= 0.5
    return float("nan"), KMAX, False


def family(name, mats):
    worst, ks, acc = 0.0, [], 0
    for M in mats:
        ref = np.abs(np.linalg.eigvals(M)).max()
        r, k, ok = rho_trace(M)
        if ok:
            acc += 1; ks.append(k)
            worst = max(worst, abs(r - ref) / max(ref, 1e-300))
    pct = np.percentile(ks, [50, 90, 99, 100]) if ks else [0, 0, 0, 0]
    print("%-12s worst accepted error %.2e   accepted %d / %d   squarings median %d p90 %d p99 %d max %d"
          % ((name, worst, acc, len(mats)) + tuple(int(v) for v in pct)))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
    rng = np.random.default_rng(0)
    family("random", [rng.standard_normal((n, n)) * rng.uniform(0.1, 0.5) for _ in range(T)])

    def structured():
        V = rng.standard_normal((n, n)); D = np.zeros((n, n)); i = 0
        r0 = rng.uniform(0.5, 1.2); kind = rng.integers(0, 4)
        mods = [r0, r0 * (1 - 10 ** rng.uniform(-12, -1)), r0 * (1 - 10 ** rng.uniform(-6, -0.5))]
        th = rng.uniform(0, np.pi)
        if kind == 0: spec = [("c", mods[0], th), ("r", mods[1] * rng.choice([-1, 1])), ("c", mods[2], rng.uniform(0, np.pi))]
        elif kind == 1: spec = [("r", mods[0]), ("r", -mods[1]), ("c", mods[2], th)]
        elif kind == 2: spec = [("c", mods[0], th), ("c", mods[1], rng.uniform(0, np.pi)), ("r", mods[2])]
        else: spec = [("r", mods[0]), ("r", mods[1]), ("r", mods[2])]
        for sp in spec:
            if sp[0] == "c" and i + 2 <= n:
                a, b = sp[1] * np.cos(sp[2]), sp[1] * np.sin(sp[2]); D[i:i + 2, i:i + 2] = [[a, b], [-b, a]]; i += 2
            elif i < n:
                D[i, i] = sp[1]; i += 1
        while i < n:
            D[i, i] = rng.uniform(-0.4, 0.4) * r0; i += 1
        return V @ D @ np.linalg.inv(V)
    family("close moduli", [structured() for _ in range(T)])

    def jordan():
        V = rng.standard_normal((n, n)); D = np.diag(rng.uniform(-0.5, 0.5, n))
        D[0, 0] = D[1, 1] = D[2, 2] = 0.9; D[0, 1] = D[1, 2] = 1.0
        return V @ D @ np.linalg.inv(V)
    family("defective", [jordan() for _ in range(T // 5)])

    def closed_loop():
        m = 2
        while True:
            A = rng.standard_normal((n, n)); A *= rng.uniform(0.8, 1.3) / np.abs(np.linalg.eigvals(A)).max()
            B = rng.standard_normal((n, m))
            Ah = A + rng.uniform(-1, 1, (n, n)) * 0.1; Bh = B + rng.uniform(-1, 1, (n, m)) * 0.1
            try:
                P = sl.solve_discrete_are(Ah, Bh, np.eye(n), np.eye(m))
            except Exception:
                continue
            K = -np.linalg.solve(np.eye(m) + Bh.T @ P @ Bh, Bh.T @ P @ Ah)
            return A + B @ K
    family("LQR loops", [closed_loop() for _ in range(T // 5)])


if __name__ == "__main__":
    main()
