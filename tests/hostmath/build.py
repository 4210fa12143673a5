"""Builds the g++ test harness (tests/hostmath/hostmath.cpp) — test fixture only."""
import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libhostmath.so")


def build():
    os.makedirs(OUT, exist_ok=True)
    csrc = os.path.join(ROOT, "lq_mpc_b200", "csrc")
    deps = [os.path.join(HERE, "hostmath.cpp")] + [os.path.join(csrc, f) for f in sorted(os.listdir(csrc))
                                                    if f.endswith(".cuh")]
    h = hashlib.sha256()
    for d in deps:
        h.update(open(d, "rb").read())
    stamp = os.path.join(OUT, "stamp")
    if os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == h.hexdigest():
        return LIB
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-x", "c++",
           os.path.join(HERE, "hostmath.cpp"), "-o", LIB, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n" + r.stderr)
    open(stamp, "w").write(h.hexdigest())
    return LIB


if __name__ == "__main__":
    print(build())
