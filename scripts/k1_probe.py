"""Development probe (GPU box): K1 parity vs the batched oracle and timing of launch-bound variants (needs a build with
LQMPC_NVCC_EXTRA=-DLQ_K1_VARIANTS for the LQMPC_K1_MINB switch to have an effect). Not part of the product."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lq_mpc_b200.engine import Engine
from oracle import np_batched as nb

out = {}
eng = Engine(0)
out["fp64_peak_tflops"] = eng.fp64_peak()
n, m = 4, 2
A, B, Q, R = nb.synth_problem(n, m, seed=0)
eng.set_problem(A, B, Q, R, Q, None, None, 30)
S = 12_500_000
g = torch.Generator(device="cuda").manual_seed(0)
dA = (torch.rand((n * n, S), device="cuda", dtype=torch.float64, generator=g) - 0.5) * 0.02
dB = (torch.rand((n * m, S), device="cuda", dtype=torch.float64, generator=g) - 0.5) * 0.02
x0 = torch.randn((n, S), device="cuda", dtype=torch.float64, generator=g)
Sp = 20000
pdA, pdB, px0 = nb.synth_samples(n, m, Sp, seed=1, e=0.01)
Pexp = nb.expert_matrix(A, B, Q, R, Q, 30)
ref = nb.eval_batch(A, B, Q, R, Q, Pexp, pdA, pdB, px0, 3, 10, T=30)
soa = nb.to_soa(pdA, pdB, px0)
for mb in sys.argv[1:] or ["2"]:
    os.environ["LQMPC_K1_MINB"] = mb
    got = eng.eval_batch(*soa, 3, 10, T=30, want=("J", "rho", "ratio", "flags", "V_N", "J_T"))
    errs = {}
    for k, kr in [("J", "J"), ("rho", "rho"), ("ratio", "ratio"), ("V_N", "Vn"), ("J_T", "JT")]:
        gg = got[k].cpu().numpy(); r = ref[kr]
        errs[k] = float(np.max(np.abs(gg - r) / np.abs(r)))
    res = {"parity": errs, "flags": np.unique(got["flags"].cpu().numpy()).tolist()}
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
        res["N%d_%d" % (Nmin, Nmax)] = {"ms": ms, "evals_per_s": S * (Nmax - Nmin + 1) / (ms * 1e-3)}
    print("MINB", mb, json.dumps(res))
    out["minb_" + mb] = res
json.dump(out, open("gpurun_out/k1_probe.json", "w"), indent=1)
