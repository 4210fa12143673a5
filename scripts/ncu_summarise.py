#!/usr/bin/env python
"""Turn gpurun_out/ ncu artefacts into the tracked summaries under profiles/ (run in the build container).
Usage: python scripts/ncu_summarise.py <tag> <kernel-report-basename> [out-prefix]
  launches_<tag>.csv                       -> profiles/<tag>_launches.csv (verbatim) + <tag>_launch_shares.txt
  <basename>.ncu-rep (ncu --set full)      -> profiles/<prefix>_metrics.csv (selected raw metrics),
                                              profiles/<prefix>_lines.txt (per-source-line instruction/stall shares),
                                              profiles/<prefix>_traffic.json (dram bytes per launch, read by bench.py)
"""
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KEEP = re.compile(r"^(dram__bytes_(read|write)\.sum|gpu__time_duration\.sum|launch__(registers_per_thread|grid_size|"
                  r"block_size|occupancy_limit_\w+|waves_per_multiprocessor)|sm__warps_active\.avg\.pct_of_peak_sustained_"
                  r"active|sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|sm__inst_executed_pipe_"
                  r"(fp64|alu|fma|xu|cbu|lsu|uniform)\.avg\.pct_of_peak_sustained_active|smsp__issue_active\.avg\.pct_of_"
                  r"peak_sustained_active|smsp__inst_executed\.sum|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|dram__throughput\.avg\.pct_of_peak_sustained_"
                  r"elapsed|smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|sm__cycles_elapsed\.max|"
                  r"smsp__sass_inst_executed_op_local_(ld|st)\.sum|sm__pipe_tensor\w*cycles_active\.avg\.pct_of_peak_"
                  r"sustained_active|l1tex__t_bytes\.sum|lts__t_bytes\.sum|sm__inst_executed_pipe_tensor_subpipe_dmma\."
                  r"avg\.pct_of_peak_sustained_active|smsp__thread_inst_executed_per_inst_executed\.ratio)$")


def launches(tag):
    src = os.path.join(GO, "launches_%s.csv" % tag)
    if not os.path.exists(src):
        return
    shutil.copy(src, os.path.join(PR, "%s_launches.csv" % tag))
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    d = defaultdict(list)
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)      # -> microseconds
        d[r[ki]].append(v)
    tot = sum(sum(v) for v in d.values())
    with open(os.path.join(PR, "%s_launch_shares.txt" % tag), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write("# %-88s %6s %12s %12s %7s\n" % ("kernel", "n", "total_us", "avg_us", "share"))
        for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            f.write("%-90s %6d %12.1f %12.1f %7.4f\n" % (k[:90], len(v), sum(v), sum(v) / len(v), sum(v) / tot))


def full(base, prefix):
    rep = os.path.join(GO, base + ".ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u = rows[0], rows[1]
    out = []
    traffic = {}
    for row in rows[2:]:
        name = row[h.index("Kernel Name")]
        for a, b, c in zip(h, u, row):
            if KEEP.match(a):
                out.append((name, a, b, c))
        rd = float(row[h.index("dram__bytes_read.sum")].replace(",", ""))
        wr = float(row[h.index("dram__bytes_write.sum")].replace(",", ""))
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd *= scale[u[h.index("dram__bytes_read.sum")]]
        wr *= scale[u[h.index("dram__bytes_write.sum")]]
        traffic = {"kernel": name, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
                   "source": "profiles/%s_metrics.csv (ncu --set full, one launch)" % prefix}
    with open(os.path.join(PR, prefix + "_metrics.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "metric", "unit", "value"])
        w.writerows(out)
    with open(os.path.join(PR, prefix + "_traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    tmp = os.path.join("/tmp", prefix + "_cs.csv")
    open(tmp, "w").write(src)
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_lines.py"), tmp, "60"],
                         capture_output=True, text=True).stdout
    open(os.path.join(PR, prefix + "_lines.txt"), "w").write(txt)


if __name__ == "__main__":
    tag, base = sys.argv[1], sys.argv[2]
    prefix = sys.argv[3] if len(sys.argv) > 3 else base
    os.makedirs(PR, exist_ok=True)
    launches(tag)
    full(base, prefix)
    print(sorted(os.listdir(PR)))
