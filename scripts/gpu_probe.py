"""Development probe (GPU box): parity of K1 vs the batched oracle + first timings. Not part of the product."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lq_mpc_b200.engine import Engine
from oracle import np_batched as nb

out = {}
eng = Engine(0)
print("dims", eng.supported_dims())
out["fp64_peak_tflops"] = eng.fp64_peak()
print("fp64 peak TFLOP/s", out["fp64_peak_tflops"])
for (n, m, e) in [(4, 2, 0.01), (2, 1, 0.05), (3, 2, 0.3), (8, 2, 0.02)]:
    A, B, Q, R = nb.synth_problem(n, m, seed=0)
    eng.set_problem(A, B, Q, R, Q, None, None, 30)
    S = 20000
    dA, dB, x0 = nb.synth_samples(n, m, S, seed=1, e=e)
    Pexp = nb.expert_matrix(A, B, Q, R, Q, 30)
    ref = nb.eval_batch(A, B, Q, R, Q, Pexp, dA, dB, x0, 3, 10, T=30)
    a, b, x = nb.to_soa(dA, dB, x0)
    got = eng.eval_batch(a, b, x, 3, 10, T=30, want=("J", "rho", "ratio", "flags", "V_N", "J_T"))
    torch.cuda.synchronize()
    def rel(k, kr):
        g = got[k].cpu().numpy(); r = ref[kr]; msk = np.isfinite(r)
        assert np.array_equal(np.isfinite(g), msk), k
        return float(np.max(np.abs(g[msk] - r[msk]) / np.maximum(np.abs(r[msk]), 1e-300)))
    res = {k: rel(k, kr) for k, kr in [("J", "J"), ("rho", "rho"), ("ratio", "ratio"), ("V_N", "Vn"), ("J_T", "JT")]}
    res["Pexp"] = float(np.max(np.abs(eng.prepared()["Pexp"] - Pexp)))
    res["flags"] = np.unique(got["flags"].cpu().numpy()).tolist()
    print(n, m, res)
    out["parity_%dx%d" % (n, m)] = res

# K2 on the golden grid: true costs + M_V of data_lq_mpc_multipleSys.npz (error sweep, N=7, T=30)
g = np.load("tests/golden/multiple_sys.npz")
A2 = np.array([[1, 0.7], [0.12, 0.4]]); B2 = np.array([[1], [1.2]])
eng.set_problem(A2, B2, 2 * np.eye(2), np.eye(1), 2 * np.eye(2), [-0.1], [0.1], 30)
eA = g["error_A_f"].reshape(4, -1); eB = g["error_B_f"].reshape(2, -1)   # [(i,j)][sample*10+level]
import json as _json
ans = _json.load(open("tests/golden/ref_known_answers.json"))
x0_vec = np.array(ans["single"]["x0_vec"]); x_start = x0_vec[:, 1]
sim = eng.simulate_batch(eA, eB, 7, 30, x0_shared=x_start)
Jt = sim["J_T"].cpu().numpy().reshape(100, 10)
print("golden true_cost_error rel", np.max(np.abs(Jt - g["true_cost_error"]) / g["true_cost_error"]),
      "n_active", np.unique(sim["n_active"].cpu().numpy()))
vexp = eng.mpc_solve_batch(None, None, 30, pts=x_start[None], S=1)["V"].cpu().numpy()[0, 0]
print("V_expert", vexp, float(g["V_expert"]), abs(vexp - float(g["V_expert"])) / vexp)
out["golden_true_cost_rel"] = float(np.max(np.abs(Jt - g["true_cost_error"]) / g["true_cost_error"]))

# timing n=4 m=2 N=10
n, m = 4, 2
A, B, Q, R = nb.synth_problem(n, m, seed=0)
eng.set_problem(A, B, Q, R, Q, None, None, 30)
for S in (1 << 20, 1 << 22, 12_500_000):
    g = torch.Generator(device="cuda").manual_seed(0)
    dA = (torch.rand((n * n, S), device="cuda", dtype=torch.float64, generator=g) - 0.5) * 0.02
    dB = (torch.rand((n * m, S), device="cuda", dtype=torch.float64, generator=g) - 0.5) * 0.02
    x0 = torch.randn((n, S), device="cuda", dtype=torch.float64, generator=g)
    for Nmin, Nmax in ((10, 10), (1, 10)):
        for _ in range(3):
            eng.eval_batch(dA, dB, x0, Nmin, Nmax)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        K = 5
        for _ in range(K):
            r = eng.eval_batch(dA, dB, x0, Nmin, Nmax)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        ev = S * (Nmax - Nmin + 1) / (ms * 1e-3)
        print("S", S, "N", Nmin, Nmax, "ms", ms, "evals/s %.3e" % ev, "unstable", int((r["flags"] & 1).sum()))
        out["time_S%d_N%d_%d" % (S, Nmin, Nmax)] = {"ms": ms, "evals_per_s": ev}
# e2e host path
S = 1 << 22
hA = torch.empty((n * n, S), dtype=torch.float64).pin_memory(); hA.uniform_(-0.01, 0.01)
hB = torch.empty((n * m, S), dtype=torch.float64).pin_memory(); hB.uniform_(-0.01, 0.01)
hx = torch.empty((n, S), dtype=torch.float64).pin_memory(); hx.normal_()
outb = None
for chunk in (1 << 18, 1 << 19, 1 << 20):
    for _ in range(2):
        outb = eng.eval_batch_host(hA, hB, hx, 10, 10, out=outb, chunk=chunk)
    t = time.perf_counter()
    for _ in range(3):
        outb = eng.eval_batch_host(hA, hB, hx, 10, 10, out=outb, chunk=chunk)
    dt = (time.perf_counter() - t) / 3
    print("e2e chunk", chunk, "ms", dt * 1e3, "evals/s %.3e" % (S / dt), "GB/s in", S * 224 / dt / 1e9)
    out["e2e_chunk%d" % chunk] = {"ms": dt * 1e3, "evals_per_s": S / dt}
d = eng.eval_batch(hA.cuda(), hB.cuda(), hx.cuda(), 10, 10)
print("host-vs-device J diff", float((d["J"].cpu() - outb["J"]).abs().max()))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
