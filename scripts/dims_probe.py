"""Development probe (GPU box): K1 / K2 / K3 throughput by state dimension and route — the register-resident
thread-per-sample instantiations (spilling for n >= 6) vs the run-time-dimension warp-per-sample route
(LQMPC_FORCE_DYN=1 routes a compiled pair through it). Not part of the product."""
import json, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def flops_per_eval(n, m, N, lyap_iters=8):
    f_ric = 4 * n ** 3 + 6 * n * n * m + 4 * n * m * m + m ** 3 / 3.0
    return N * f_ric + 2 * n * n * m + (2 * n * n * m + 2 * n * m * m) + lyap_iters * 6 * n ** 3 + 10 * n ** 3 + (2 * n * n + 2 * n)


def one(n, m, N, S):
    import torch
    from lq_mpc_b200.engine import Engine
    from lq_mpc_b200 import sampling as sp
    eng = Engine(0)
    A, B, Q, R = sp.synth_problem(n, m, seed=0)
    lo, hi = -0.5 * np.ones(m), 0.5 * np.ones(m)
    eng.set_problem(A, B, Q, R, Q, lo, hi, 30)
    g = torch.Generator(device="cuda").manual_seed(1)
    dA = (torch.rand((n * n, S), device="cuda", dtype=torch.float64, generator=g) - 0.5) * 0.02
    dB = (torch.rand((n * m, S), device="cuda", dtype=torch.float64, generator=g) - 0.5) * 0.02
    x0 = torch.randn((n, S), device="cuda", dtype=torch.float64, generator=g) * 0.3

    def timed(fn, reps=3):
        fn(); torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best
    t1 = timed(lambda: eng.eval_batch(dA, dB, x0, N, N))
    t2 = timed(lambda: eng.mpc_solve_batch(dA, dB, N, x0=x0, want=("V", "flags")))
    t3 = timed(lambda: eng.bounds_batch(dA, dB, N, 0.01, 0.01, 0.5, x0, (0.1, 1, 0.6), 1.0))
    peak = eng.fp64_tensor_peak()
    return {"n": n, "m": m, "N": N, "S": S, "route": "dyn" if os.environ.get("LQMPC_FORCE_DYN") else "auto",
            "k1_ms": t1, "k1_evals_per_s": S / t1 * 1e3, "k1_frac_fp64": S * flops_per_eval(n, m, N) / (t1 * 1e-3) / 1e12 / peak,
            "k2_solve_ms": t2, "k2_solves_per_s": S / t2 * 1e3, "k3_ms": t3, "k3_evals_per_s": S / t3 * 1e3, "peak_tf": peak}


if __name__ == "__main__":
    if len(sys.argv) > 1:
        n, m, N, S = (int(v) for v in sys.argv[1:5])
        print(json.dumps(one(n, m, N, S)))
        sys.exit(0)
    rows = []
    for (n, m, S_thread, S_dyn) in [(4, 2, 2_000_000, 200_000), (6, 2, 1_000_000, 200_000), (8, 2, 500_000, 200_000),
                                    (5, 2, 0, 200_000), (7, 3, 0, 200_000), (16, 4, 0, 50_000), (32, 8, 0, 10_000)]:
        for route, S in (("auto", S_thread), ("dyn", S_dyn)):
            if S == 0:
                continue
            env = dict(os.environ)
            env.pop("LQMPC_FORCE_DYN", None)
            if route == "dyn":
                env["LQMPC_FORCE_DYN"] = "1"
            r = subprocess.run([sys.executable, __file__, str(n), str(m), "10", str(S)], env=env, capture_output=True, text=True)
            line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:]
            print(line, flush=True)
