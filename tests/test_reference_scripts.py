"""The reference's own driver scripts, executed UNCHANGED on the GPU engine ("what must keep working").

`__graft_entry__.build()` stages byte-for-byte copies of working_example_single.py, working_example_multiple.py,
mpc_test.py and behavior_test.py under oracle/_ref/scripts/ (git-ignored, travels to the GPU box). Each is run with
`runpy` with `lq_mpc_b200/dropin` first on sys.path, so its `import utils`, `import utils_class`, `import control`
resolve to the drop-in modules; what it prints is compared with the answers the UNTOUCHED reference gave for the same
script (tests/golden/ref_known_answers.json, the shipped data_lq_mpc_multipleSys.npz).
"""
import contextlib
import io
import os
import re
import runpy
import sys

import numpy as np
import pytest

from tests.conftest import ROOT, relerr

pytestmark = pytest.mark.gpu
SCRIPTS = os.path.join(ROOT, "oracle", "_ref", "scripts")
DROPIN = os.path.join(ROOT, "lq_mpc_b200", "dropin")
NUM = r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?"


def _have(name):
    return os.path.isfile(os.path.join(SCRIPTS, name))


def _run(name, cwd, source=None):
    """Execute a staged script (or `source`, a transformed copy of its text) with the drop-in modules first on
    sys.path; returns (stdout, globals)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not _have(name):
        pytest.skip("oracle/_ref/scripts not staged (build() ran without /root/reference)")
    saved_path, saved_cwd = list(sys.path), os.getcwd()
    saved_mods = {k: sys.modules.pop(k, None) for k in ("utils", "utils_class", "control")}
    sys.path.insert(0, DROPIN)
    os.chdir(cwd)
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            if source is None:
                g = runpy.run_path(os.path.join(SCRIPTS, name), run_name="__main__")
            else:
                g = {"__name__": "__main__", "__file__": os.path.join(SCRIPTS, name)}
                exec(compile(source, name, "exec"), g)
    finally:
        os.chdir(saved_cwd)
        sys.path[:] = saved_path
        for k, v in saved_mods.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v
    return buf.getvalue(), g


def _after(text, label):
    m = re.search(re.escape(label) + r"[^\n]*?(" + NUM + ")", text)
    assert m, "label %r not printed" % label
    return float(m.group(1))


def test_working_example_single_script_unchanged(tmp_path, known):
    out, g = _run("working_example_single.py", tmp_path)
    k = known["single"]
    for label, want in (("The chosen energy bound is", k["M_V"]), ("The coefficient C_K", k["ex"]["C_K"]),
                        ("The coefficient rho_K", k["ex"]["rho_K"]), ("The coefficient gamma", k["ex"]["gamma"]),
                        ("The coefficient rho_gamma", k["ex"]["rho_gamma"]),
                        ("The radius epsilon corresponding to the LQR is", k["epsilon_lqr"]),
                        ("The critical prediction horizon is", k["bar"]["N_0"]),
                        ("The minimum required prediction horizon:", k["omega_eta"]["N_min"]),
                        ("The error threshold:", k["omega_eta"]["err_th"])):
        assert abs(_after(out, label) - want) <= 1e-9 * abs(want), label
    assert abs(g["info_decrease"]["xi"] - k["decrease"]["xi"]) <= 1e-9 * k["decrease"]["xi"]
    assert abs(g["info_decrease"]["eta"] - k["decrease"]["eta"]) <= 1e-9 * k["decrease"]["eta"]
    assert abs(g["info_bound"]["alpha"] - k["bound"]["alpha"]) <= 1e-9 * k["bound"]["alpha"]
    assert abs(g["info_bound"]["beta"] - k["bound"]["beta"]) <= 1e-9 * k["bound"]["beta"]
    assert relerr(g["K_lqr"], k["K_lqr"]) < 1e-9 and np.max(np.abs(g["x0_vec"] - np.array(k["x0_vec"]))) < 1e-12
    assert str(g["info_decrease"]) in out and str(g["info_bound"]) in out        # the script's own two print() calls


def test_mpc_test_script_unchanged(tmp_path, known):
    out, g = _run("mpc_test.py", tmp_path)
    k = known["mpc_test"]
    assert abs(_after(out, "Closed-loop cost:") - k["J_T"]) <= 1e-9 * k["J_T"]
    assert abs(g["MPC_traj_info"]["J_T"] - k["J_T"]) <= 1e-9 * k["J_T"]
    assert np.max(np.abs(g["MPC_traj_info"]["U"] - np.array(k["U"]))) < 1e-12
    assert np.max(np.abs(g["MPC_traj_info"]["X"] - np.array(k["X"]))) < 1e-12


def test_behavior_test_script_unchanged(tmp_path, known):
    out, g = _run("behavior_test.py", tmp_path)
    k = known["behavior"]
    assert abs(g["M_V"] - k["M_V"]) <= 1e-9 * k["M_V"]
    for name, key in (("data_xi", "xi"), ("data_alpha", "alpha"), ("data_beta", "beta")):
        for part in ("error", "horizon"):
            assert relerr(g[name][part], k[key][part]) < 1e-9, (name, part)
    for f in ("X", "Y", "J_MPC_true", "J_MPC_bound", "V_OPC"):                       # behavior_test.py:98-103
        assert relerr(g["data_surface"][f], k["mesh"][f]) < 1e-9, f


def test_working_example_multiple_script_unchanged_and_recomputing(tmp_path, golden):
    """(1) the shipped script as it is: loads data_lq_mpc_multipleSys.npz from cwd, hands the tables to the plotter
    (headless stand-in printing K5's column statistics), prints rho(A) and V_expert; (2) the same script with its own
    commented-out "recompute the data" block (working_example_multiple.py:95-102) enabled by removing the two
    quote lines around it: LQ_RDP_Behavior_Multiple(...).data_generation(...) runs on the GPU and must reproduce the
    shipped file."""
    keys = ["error", "horizon", "V_expert", "alpha_table_error", "beta_table_error", "xi_table_error",
            "bound_table_error", "true_cost_error", "alpha_table_horizon", "beta_table_horizon",
            "xi_table_horizon", "bound_table_horizon", "true_cost_horizon"]
    np.savez(tmp_path / "data_lq_mpc_multipleSys.npz", **{k: golden[k] for k in keys})
    for f in ("error_A_f", "error_B_f"):
        np.save(tmp_path / (f + ".npy"), golden[f])
    out, g = _run("working_example_multiple.py", tmp_path)
    assert abs(_after(out, "The spectral radius of matrix A is") - 1.117133072292284) < 1e-12
    assert abs(float(out.strip().splitlines()[-1]) - float(golden["V_expert"])) < 1e-12
    assert "[plotter_error] table (100, 10)" in out and "[plotter_true_cost]" in out
    mx = golden["bound_table_error"].max(axis=0)
    assert ("max %s" % mx) in out                                      # K5 statistics of the table the plotter received
    # (2) enable the script's own recompute block
    src = open(os.path.join(SCRIPTS, "working_example_multiple.py")).read()
    lines = src.split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith("# ----------------- Main commands"))
    q = [i for i in range(start, len(lines)) if lines[i].strip() == "'''"][:2]
    assert len(q) == 2
    del lines[q[1]], lines[q[0]]
    lines = [l for l in lines if not l.startswith("data_table = np.load(")]
    os.remove(tmp_path / "data_lq_mpc_multipleSys.npz")
    out2, g2 = _run("working_example_multiple.py", tmp_path, source="\n".join(lines))
    saved = dict(np.load(tmp_path / "data_lq_mpc_multipleSys.npz"))
    for k in keys:
        assert relerr(g2["data_table"][k], golden[k]) < 1e-9, k
        assert relerr(saved[k], golden[k]) < 1e-9, k
