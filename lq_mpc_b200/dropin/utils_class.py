"""Put this directory first on sys.path / PYTHONPATH and the reference's scripts (`import utils_class`) run on the
B200 engine unchanged."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
from lq_mpc_b200.utils_class import *  # noqa: F401,F403,E402
import lq_mpc_b200.utils_class as _m  # noqa: E402

globals().update({k: getattr(_m, k) for k in dir(_m) if not k.startswith('__')})


class _HeadlessPlotter:
    """Stand-in for the reference's Plotter_* classes (presentation, out of scope): every plotting method prints
    the per-column statistics (computed by K5 on the GPU) of the 2-D tables it is handed instead of drawing."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        def method(*args, **kwargs):
            import numpy as np
            from lq_mpc_b200.utils import column_statistics
            for a in args:
                if isinstance(a, np.ndarray) and a.ndim == 2 and a.shape[0] > 1:
                    mx, mn, mean, std = column_statistics(a)
                    print("[%s] table %s: max %s min %s mean %s std %s" % (name, a.shape, mx, mn, mean, std))
        return method


Plotter_MPC = Plotter_PF_LQMPC = Plotter_PF_LQMPC_Multiple = _HeadlessPlotter
