"""Development probe (GPU box): the sweep's two K2 launches — ring solves (M_V over 8 ring points) and the closed
loop (T = 30) — on the shipped 2-state example at several horizons, for every library given on the command line
(A/B builds from scripts/unit_variants.sh; the engine is loaded through LQMPC_LIB in a sub-process each).
usage: python scripts/k2_probe.py [lib.so ...]        Not part of the product."""
import json, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(S=1_000_000):
    import torch
    from lq_mpc_b200.engine import Engine
    eng = Engine(0)
    A = np.array([[1, 0.7], [0.12, 0.4]]); B = np.array([[1], [1.2]])
    eng.set_problem(A, B, 2 * np.eye(2), np.eye(1), 2 * np.eye(2), [-0.1], [0.1], 30)
    K = eng.dlqr_batch(S=1)["K"].cpu().numpy()[:, 0]
    eps = 1.0 / (float(K @ K) / 2.0 / 0.01)
    th = np.linspace(0, 2 * (1 - 1 / 8) * np.pi, 8)
    ring = (1.5 * np.sqrt(eps) / np.sqrt(2.0)) * np.vstack([np.cos(th), np.sin(th)])
    x_start = ring[:, 1]
    g = torch.Generator(device="cuda").manual_seed(3)
    lev = torch.linspace(1e-3, 1e-2, 10, device="cuda", dtype=torch.float64).repeat(S // 10)
    dA = (torch.rand((4, S), device="cuda", dtype=torch.float64, generator=g) - 0.5) * lev
    dB = (torch.rand((2, S), device="cuda", dtype=torch.float64, generator=g) - 0.5) * lev

    def timed(fn, reps=5):
        fn(); torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best
    out = {}
    chk = 0.0
    for N in (5, 10, 25, 50):
        out["ring_N%d_ms" % N] = round(timed(lambda: eng.mpc_solve_batch(dA, dB, N, pts=ring.T, want=("M_V",))), 3)
        out["sim_N%d_ms" % N] = round(timed(lambda: eng.simulate_batch(dA, dB, N, 30, x0_shared=x_start, want=("J_T",))), 3)
        chk += float(eng.mpc_solve_batch(dA, dB, N, pts=ring.T, want=("M_V",))["M_V"].sum())
        chk += float(eng.simulate_batch(dA, dB, N, 30, x0_shared=x_start, want=("J_T",))["J_T"].sum())
    out["ring_total_ms"] = round(sum(v for k, v in out.items() if k.startswith("ring")), 3)
    out["sim_total_ms"] = round(sum(v for k, v in out.items() if k.startswith("sim")), 3)
    out["checksum"] = repr(chk)
    return out


if __name__ == "__main__":
    if os.environ.get("K2_PROBE_CHILD"):
        print(json.dumps(one()))
        sys.exit(0)
    libs = sys.argv[1:] or [""]
    for lib in libs:
        env = dict(os.environ, K2_PROBE_CHILD="1")
        if lib:
            env["LQMPC_LIB"] = os.path.abspath(lib)
        r = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True)
        print(os.path.basename(lib) or "default", r.stdout.strip() or r.stderr[-400:], flush=True)
