"""Put this directory first on sys.path / PYTHONPATH and the reference's scripts (`import control`) run on the
B200 engine unchanged."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
from lq_mpc_b200.control import *  # noqa: F401,F403,E402
import lq_mpc_b200.control as _m  # noqa: E402

globals().update({k: getattr(_m, k) for k in dir(_m) if not k.startswith('__')})
