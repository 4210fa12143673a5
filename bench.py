#!/usr/bin/env python
"""bench.py — headline benchmark of the lq_mpc hot path on B200 (see DESIGN.md "Measurement").

Metric (BASELINE.json): MPC performance evals/sec, 1 eval = one (sample, horizon) pair = Riccati gain of that
horizon on the estimated model + closed loop on the true plant + J_inf (Lyapunov doubling) + spectral-radius
stability check + performance ratio.  Workload: cfg-synth-4-2-10 (n=4, m=2, N=10, Q=I, R=I, unconstrained),
1.25e7 seeded samples (dA, dB, x0) PER GPU (= BASELINE's 1e8 samples at 8 GPUs; weak scaling), followed by the
per-column worst-case statistics (K5, one pass) and — at N>1 — the engine's only collective, one tiny NCCL
all-gather of the per-column moments.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm  (torchrun launches it for N > 1)
  python bench.py --impl reference ...                         the CPU arm: the numpy oracle port on all host cores

One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_DIM, M_DIM, HORIZON = 4, 2, 10
S_PER_GPU = 12_500_000
SEED_PROBLEM, SEED_SAMPLES = 0, 1
# BASELINE.json configs[3] is the headline (the >= 1e7 evals/s target is quoted on it); configs[4] is selectable:
WORKLOADS = {
    "cfg-synth-4-2-10": dict(n=4, m=2, N=10, S=12_500_000, e=0.01, tiled=False, kernel="eval_kernel<4,2>"),
    "cfg-synth-32-8-30": dict(n=32, m=8, N=30, S=125_000, e=1e-3, tiled=True,
                              kernel="tiled_eval_kernel<32,8> (+ tiled_rho_kernel<32> for entries it hands over)"),
}


def set_workload(name):
    global N_DIM, M_DIM, HORIZON, S_PER_GPU, WL
    WL = dict(WORKLOADS[name], name=name)
    N_DIM, M_DIM, HORIZON, S_PER_GPU = WL["n"], WL["m"], WL["N"], WL["S"]


WL = dict(WORKLOADS["cfg-synth-4-2-10"], name="cfg-synth-4-2-10")


def flops_per_eval(n, m, N, lyap_iters=8):
    """dense algorithmic FP64 flops of one eval (FMA = 2 flops, no symmetry savings)."""
    f_ric = 4 * n ** 3 + 6 * n * n * m + 4 * n * m * m + m ** 3 / 3.0
    return N * f_ric + 2 * n * n * m + (2 * n * n * m + 2 * n * m * m) + lyap_iters * 6 * n ** 3 + 10 * n ** 3 + (
        2 * n * n + 2 * n)


def bytes_per_eval(n, m):
    """compulsory HBM traffic of one eval: read dA, dB, x0; write J, rho, ratio, flags."""
    return (n * n + n * m + n) * 8 + 4 * 8


# ---------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except (ValueError, IndexError):
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------- CPU arm
def _cpu_worker(args):
    os.environ["OMP_NUM_THREADS"] = "1"
    try:                                    # the GPU arm may have bound its process to one NUMA node: the CPU arm
        os.sched_setaffinity(0, range(os.cpu_count() or 1))            # gets every core of the box
    except OSError:
        pass
    first, count, (n, m, N, e) = args
    import numpy as np  # noqa
    from oracle import np_batched as nb
    A, B, Q, R = nb.synth_problem(n, m, seed=SEED_PROBLEM)
    Pexp = nb.expert_matrix(A, B, Q, R, Q, 30)
    dA, dB, x0 = nb.synth_samples(n, m, count, seed=SEED_SAMPLES, first=first, e=e)
    t = time.perf_counter()
    out = nb.eval_batch(A, B, Q, R, Q, Pexp, dA, dB, x0, N, N)
    return time.perf_counter() - t, float(out["ratio"].max())


def cpu_port_throughput(per_worker, workers=None):
    """The oracle port (oracle/np_batched.py) on `workers` host processes, `per_worker` samples each.
    Returns (evals/s over the wall clock of the pool, workers, samples)."""
    import multiprocessing as mp
    workers = workers or max(1, min(os.cpu_count() or 1, 64))
    ctx = mp.get_context("spawn")
    cfg = (N_DIM, M_DIM, HORIZON, WL["e"])
    if WL["tiled"]:
        per_worker = max(64, per_worker // 40)                          # ~800x the flops per eval of the n=4 case
    with ctx.Pool(workers) as pool:
        pool.map(_cpu_worker, [(0, 32 if WL["tiled"] else 256, cfg)] * workers)   # spawn + import warm-up
        t = time.perf_counter()
        pool.map(_cpu_worker, [(w * per_worker, per_worker, cfg) for w in range(workers)])
        wall = time.perf_counter() - t
    total = workers * per_worker
    return total / wall, workers, total


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    per_worker = 20_000
    vals = []
    for _ in range(args.warmup):
        cpu_port_throughput(2_000)
    workers = total = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, workers, total = cpu_port_throughput(per_worker)
        vals.append(v)
    wall = time.perf_counter() - t0
    value = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": "mpc_evals_per_sec", "value": value, "unit": "evals/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * wall / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.gpus, note="CPU arm: bounded sample per step"),
            "cpu_baseline": {"value": value, "unit": "evals/s", "cores": workers, "kind": "port",
                             "sample": "%d samples per step (%d per process) of the same seeded workload; oracle/"
                                       "np_batched.py (batched numpy/LAPACK), one process per host core" %
                                       (total, total // max(1, workers))},
            "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def workload_config(n_gpus, note=None):
    in_gb = S_PER_GPU * (N_DIM * N_DIM + N_DIM * M_DIM + N_DIM) * 8 / 1e9
    cfg = {"workload": "%s: n=%d m=%d N=%d Q=I R=I unconstrained, %.3g seeded (dA,dB,x0) samples per GPU "
                       "(%.3g total), J_inf + rho + ratio + flags per sample, then per-column worst-case stats"
                       % (WL["name"], N_DIM, M_DIM, HORIZON, S_PER_GPU, S_PER_GPU * n_gpus),
           "n": N_DIM, "m": M_DIM, "N": HORIZON, "samples_per_gpu": S_PER_GPU, "evals_per_sample": 1,
           "l2": "inputs (%.2f GB per step) exceed the 126 MB L2; no flush needed" % in_gb,
           "parallelism": "sample-sharded x%d, one final all-gather of per-column moments (6 doubles/column)" % n_gpus}
    if note:
        cfg["note"] = note
    return cfg


# ---------------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from lq_mpc_b200 import sampling as sp
    from lq_mpc_b200.engine import Engine
    from lq_mpc_b200.stats import column_moments_device, merge_moments

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    from lq_mpc_b200.runtime import bind_to_gpu_numa
    numa_cpus = bind_to_gpu_numa(local)          # pinned host shards below are then first-touched on the GPU's node
    eng = Engine(local)
    A, B, Q, R = sp.synth_problem(N_DIM, M_DIM, seed=SEED_PROBLEM)
    tiled = WL["tiled"]
    if tiled:
        eng.set_problem_tiled(A, B, Q, R, Q, 30)
    else:
        eng.set_problem(A, B, Q, R, Q, None, None, 30)
    S = S_PER_GPU
    n, m = N_DIM, M_DIM

    # ---- this rank's shard of the seeded workload, generated on the host into pinned buffers
    # (K1: struct-of-arrays [element][sample]; K4: array-of-matrices [sample][element], one sample contiguous for TMA)
    shp = (lambda k: (S, k)) if tiled else (lambda k: (k, S))
    hA = torch.empty(shp(n * n), dtype=torch.float64).pin_memory()
    hB = torch.empty(shp(n * m), dtype=torch.float64).pin_memory()
    hx = torch.empty(shp(n), dtype=torch.float64).pin_memory()
    sp.synth_samples_soa(n, m, S, seed=SEED_SAMPLES, first=rank * S, e=WL["e"], aos=tiled,
                         out=(hA.numpy(), hB.numpy(), hx.numpy()))
    dA, dB, x0 = hA.cuda(non_blocking=True), hB.cuda(non_blocking=True), hx.cuda(non_blocking=True)
    torch.cuda.synchronize()
    evaluate = (lambda a, b, c: eng.eval_batch_tiled(a, b, c, HORIZON, HORIZON)) if tiled else \
        (lambda a, b, c: eng.eval_batch(a, b, c, HORIZON, HORIZON))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        r = evaluate(dA, dB, x0)                              # K1/K4 write J, rho, ratio into one [3][S] table
        # K5 (one pass) + the only collective (all-gather), all enqueued on the device: the [world][3][6] moments
        # stay in HBM; they are merged on the host after the timed region (the e2e leg reads results back per step)
        return r, column_moments_device(eng, r["table"])

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- (1) device-resident throughput: W warm-up steps, K timed steps, CUDA events, max over ranks
    # the clock sampler (one `nvidia-smi -lms` process) starts BEFORE the warm-up and gets time to initialise (its NVML
    # start-up takes driver locks for a few hundred ms: not at the edge of the timed region)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(1.0)
    # W warm-up steps, then as many more as keep the GPU under this load for ~0.5 s (the kernel draws the board's power
    # cap: the timed region should not start on a cold power state). The count is agreed across ranks — every step
    # carries the all-gather, so ranks must not leave the warm-up at different steps.
    n_warm = max(3, args.warmup)
    for _ in range(n_warm):
        r, st = step()      # bound exactly as in the timed loop: two result sets alternate, so the second 350 MB block
    barrier()               # is allocated HERE (a cudaMalloc of it inside the timed region cost 6-150 ms at step 2)
    t_w = time.perf_counter()
    r, st = step()
    torch.cuda.synchronize()
    extra = int(max_over_ranks(float(min(400, int(0.5 / max(time.perf_counter() - t_w, 1e-4))))))
    for _ in range(extra):
        r, st = step()
    n_warm += 1 + extra
    barrier()
    if rank == 0:
        sampler.rows.clear()                      # samples from here on: the three timed regions
    launches0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    trace = os.environ.get("LQMPC_BENCH_TRACE") == "1"
    tev, thost = [], []
    e0.record()
    for _ in range(args.steps):
        r, st = step()
        if trace:                                  # development: where does a slow region lose its time?
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            tev.append(ev)
            thost.append(time.perf_counter())
    e1.record()
    barrier()
    ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    if trace and rank == 0:
        d = [e0.elapsed_time(tev[0])] + [tev[i].elapsed_time(tev[i + 1]) for i in range(len(tev) - 1)]
        med = sorted(d)[len(d) // 2]
        sys.stderr.write("trace: median %.3f ms, slow steps %s, host enqueue span %.1f ms\n" % (
            med, [(i, round(x, 1)) for i, x in enumerate(d) if x > 1.3 * med], (thost[-1] - thost[0]) * 1e3))
    st = merge_moments(st.cpu().numpy())
    launches = eng.launch_count - launches0
    # ---- (2) the dominant kernel alone (K1), same data, for the roofline
    barrier()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(args.steps):
        evaluate(dA, dB, x0)
    k1.record()
    barrier()
    ms_kernel = k0.elapsed_time(k1) / args.steps
    # ---- (2b) supplementary: every horizon 1..N emitted from ONE nested Riccati recursion per sample (the metric's
    #      "samples x horizons" reading; not the headline, which counts one eval per sample at the quoted horizon)
    ms_nested = None
    if not tiled:
        nsteps = max(2, args.steps // 4)
        for _ in range(2):
            eng.eval_batch(dA, dB, x0, 1, HORIZON)
        barrier()
        n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0.record()
        for _ in range(nsteps):
            eng.eval_batch(dA, dB, x0, 1, HORIZON)
        n1.record()
        barrier()
        ms_nested = max_over_ranks(n0.elapsed_time(n1) / nsteps)
    # ---- (3) end to end through the host-buffer entry point: pinned host -> H2D -> K1 -> D2H, every step
    outb = None
    if tiled:       # K4: chunked H2D -> K4a + K4b -> D2H pipeline from the pinned shard (lqmpc_eval_batch_tiled_host)
        def host_step(prev):
            return eng.eval_batch_tiled_host(hA, hB, hx, HORIZON, HORIZON, out=prev, chunk=16384)
    else:
        def host_step(prev):
            return eng.eval_batch_host(hA, hB, hx, HORIZON, HORIZON, out=prev, chunk=1 << 19)
    for _ in range(2):
        outb = host_step(outb)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        outb = host_step(outb)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
    clocks = sampler.stop() if rank == 0 else None        # sampled across all three timed regions (GPU busy throughout)
    same = bool(torch.equal(outb["J"][0, :4096], r["J"][0, :4096].cpu()))
    peak_fp64 = eng.fp64_peak() if rank == 0 else 0.0
    peak_dmma = eng.fp64_tensor_peak() if rank == 0 else 0.0
    unstable = int((r["flags"] & 1).sum().item())
    if world > 1:
        dist.barrier()

    if rank == 0:
        evals = S * world
        value = evals / (ms_step * 1e-3)
        fl = flops_per_eval(n, m, HORIZON)
        by = bytes_per_eval(n, m)
        ach_tf = S * fl / (ms_kernel * 1e-3) / 1e12
        ach_gb = S * by / (ms_kernel * 1e-3) / 1e9
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic = None
        try:
            tf_name = "k4_traffic.json" if tiled else "k1_traffic.json"
            traffic = json.load(open(os.path.join(ROOT, "profiles", tf_name))).get("dram_bytes_per_launch")
            # (latest `ncu --set full` capture of eval_kernel<4,2> on this workload; see profiles/README.md)
        except (OSError, ValueError):
            pass
        line = {
            "metric": "mpc_evals_per_sec", "value": value, "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "warmup_total_steps": n_warm, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world),
            "roofline": {
                # Both workloads are bound by the SM's FP64 pipe. On B200 the FP64 tensor-core MMA (mma.sync.m8n8k4.f64)
                # executes on that same pipe (scripts/probes/fp64_pipe_probe.cu: one MMA holds it for 16 cycles = 256
                # FMAs at the DFMA rate), so the pipe's peak is the MMA-stream figure (= 148 SMs x 64 FMA/clk); a
                # DFMA-only stream tops out ~9 % lower (issue / register bandwidth). `peak` is the higher, harder one for
                # both kernels; the DFMA-stream figure and the fraction against it are reported beside it.
                "bound": "tensor", "bound_detail": "FP64 pipe: " + (
                    "tensor-core MMAs (DMMA)" if tiled else
                    "scalar DFMA code, no MMA issued; the FP64 tensor-core peak is the peak of this same pipe"),
                "achieved": ach_tf, "peak": max(peak_dmma, peak_fp64), "unit": "TFLOP/s",
                "frac": (ach_tf / max(peak_dmma, peak_fp64)) if peak_fp64 else None, "traffic": traffic,
                "fp64_vector_peak": peak_fp64, "frac_of_dfma_stream_peak": (ach_tf / peak_fp64) if peak_fp64 else None,
                "kernel": WL["kernel"], "kernel_ms": ms_kernel,
                "algorithmic_flops_per_eval": fl, "algorithmic_bytes_per_eval": by,
                "peak_source": "measured in this run: mma.sync.m8n8k4.f64 stream (lqmpc_fp64_tensor_peak) and DFMA-chain "
                               "stream (lqmpc_fp64_peak) micro-benchmarks; MEASURED_PEAKS.json has no FP64 figure",
                "hbm": {"achieved": ach_gb, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gb / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650"}},
            "e2e": {"value": evals / e2e_s, "unit": "evals/s",
                    "h2d_bytes_per_step": int(S * (n * n + n * m + n) * 8),
                    "d2h_bytes_per_step": int(S * (3 * 8 + 4)), "matches_device_path": same,
                    "host_numa_binding": ("process pinned to the %d cores NVML reports local to its GPU" % len(numa_cpus))
                    if numa_cpus else "none (NVML affinity unavailable)",
                    "api": "lqmpc_eval_batch_host (pinned host SoA in, J/rho/ratio/flags tables out)" if not tiled
                    else "lqmpc_eval_batch_tiled_host (pinned host array-of-matrices in, J/rho/ratio/flags tables out)"},
            "gpu_launches": int(launches),
            "nested_horizons": None if ms_nested is None else {
                "what": "horizons 1..%d of every sample from one nested Riccati recursion (K1 alone, device-resident)"
                        % HORIZON, "evals_per_s": S * world * HORIZON / (ms_nested * 1e-3), "ms": ms_nested},
            "clocks": clocks,
            "worst_case": {"ratio_max": float(st["max"][2]), "ratio_mean": float(st["mean"][2]),
                           "ratio_std": float(st["std"][2]), "rho_max": float(st["max"][1]), "unstable": unstable},
        }
        if world == 1 and not args.no_cpu_baseline:
            v, workers, total = cpu_port_throughput(8_000)
            line["cpu_baseline"] = {"value": v, "unit": "evals/s", "cores": workers, "kind": "port",
                                    "sample": "%d samples of the same seeded workload (8000 per process); "
                                              "oracle/np_batched.py, one process per host core" % total}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="cfg-synth-4-2-10", choices=sorted(WORKLOADS))
    args = ap.parse_args()
    set_workload(args.workload)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
