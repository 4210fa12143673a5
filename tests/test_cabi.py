"""CPU tier — the drop-in boundary: liblqmpc_b200.so loads, exports every symbol include/lqmpc_b200.h declares, the
ctypes table in lq_mpc_b200/engine.py binds exactly that set, and — there being NO CPU path — every entry fails
loudly without a CUDA device. No compute call is made here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lqmpc_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lqmpc_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from lq_mpc_b200 import _build
    return ctypes.CDLL(_build.build())


def test_header_declares_the_expected_entry_points():
    syms = _declared_symbols()
    for s in ("lqmpc_create", "lqmpc_destroy", "lqmpc_set_problem", "lqmpc_eval_batch", "lqmpc_eval_batch_host",
              "lqmpc_mpc_solve_batch", "lqmpc_simulate_batch", "lqmpc_bounds_batch", "lqmpc_dlqr_batch",
              "lqmpc_column_stats", "lqmpc_column_sqdev", "lqmpc_column_moments", "lqmpc_last_error", "lqmpc_abi_version"):
        assert s in syms


def test_library_exports_every_declared_symbol(lib):
    for s in _declared_symbols():
        assert hasattr(lib, s), "missing export: " + s


def test_ctypes_binding_covers_exactly_the_header():
    from lq_mpc_b200 import engine
    assert sorted(engine.ABI.keys()) == _declared_symbols()
    engine.load_library()                               # binds every symbol; raises on a missing one


def test_header_compiles_as_plain_c(tmp_path):
    """The boundary is a C ABI: the header must be consumable by a C compiler, no C++/torch types."""
    c = tmp_path / "t.c"
    c.write_text('#include "lqmpc_b200.h"\nint main(void){ return LQMPC_OK + (int)sizeof(lqmpc_ctx*) * 0; }\n')
    import subprocess
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(c), "-o",
                        str(tmp_path / "t.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_abi_version_and_dims(lib):
    lib.lqmpc_abi_version.restype = ctypes.c_int
    lib.lqmpc_supported_dims.restype = ctypes.c_char_p
    assert lib.lqmpc_abi_version() >= 1
    dims = lib.lqmpc_supported_dims().decode().split(",")
    assert "4x2" in dims and "2x1" in dims and "1x1" in dims


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for the GPU-less container")
    h = ctypes.c_void_p()
    lib.lqmpc_create.restype = ctypes.c_int
    rc = lib.lqmpc_create(ctypes.byref(h), 0, None)
    assert rc == -2 and not h.value                      # LQMPC_ENODEVICE
    from lq_mpc_b200.engine import Engine, EngineError
    with pytest.raises(EngineError):
        Engine(0)
    from lq_mpc_b200.utils_class import LQ_MPC_Controller
    import numpy as np
    with pytest.raises(EngineError):
        LQ_MPC_Controller(3, np.eye(2), np.ones((2, 1)), np.eye(2), np.eye(1), np.eye(2),
                          np.array([[10.0], [-10.0]])).solve(np.ones(2), np.zeros((2, 3)), np.zeros((1, 3)))


def test_product_never_imports_the_oracle():
    """Only tests/, smoke() and bench.py's CPU legs may touch oracle/ (a product path through it voids parity)."""
    pkg = os.path.join(ROOT, "lq_mpc_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), os.path.join(dp, f)
                if f.endswith(".py"):                   # (a .cuh comment may mention the test harness)
                    assert "hostmath" not in txt, os.path.join(dp, f)
