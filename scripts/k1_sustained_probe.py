"""Development probe (GPU box, -DLQ_K1_VARIANTS build): SUSTAINED K1 time per launch for the register-budget variants
(the kernel runs into the board's power cap: the variant with the fewest instructions per eval may win there even if
it ties from a cold power state)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lq_mpc_b200.engine import Engine
from oracle import np_batched as nb
eng = Engine(0)
n, m = 4, 2
A, B, Q, R = nb.synth_problem(n, m, seed=0)
eng.set_problem(A, B, Q, R, Q, None, None, 30)
S = 12_500_000
g = torch.Generator(device="cuda").manual_seed(0)
dA = (torch.rand((n * n, S), device="cuda", dtype=torch.float64, generator=g) - 0.5) * 0.02
dB = (torch.rand((n * m, S), device="cuda", dtype=torch.float64, generator=g) - 0.5) * 0.02
x0 = torch.randn((n, S), device="cuda", dtype=torch.float64, generator=g)
for mb in (sys.argv[1:] or ["2", "4", "5"]):
    os.environ["LQMPC_K1_MINB"] = mb
    for (Nmin, reps) in ((10, 200), (1, 50)):
        for _ in range(reps * 3 // 4):
            r = eng.eval_batch(dA, dB, x0, Nmin, 10)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            r = eng.eval_batch(dA, dB, x0, Nmin, 10)
        e1.record(); torch.cuda.synchronize()
        print("MINB", mb, "horizons %d..10" % Nmin, "sustained ms %.4f" % (e0.elapsed_time(e1) / reps))
