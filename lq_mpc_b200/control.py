"""Drop-in for the one python-control entry point the reference uses: `ct.dlqr` (utils_class.py:761,840,923;
working_example_single.py:39). The DARE is solved on the GPU (structure-preserving doubling, csrc/bounds.cuh)."""
import numpy as np

from . import runtime as _rt


def dlqr(A, B, Q, R):
    """Returns (K, P, E) with the python-control convention u = -K x; E = eigenvalues... of A - B K are NOT computed
    on the host: E is returned as None-filled placeholder array of spectral radius (callers in the reference
    discard it: `K, _, _ = ct.dlqr(...)`)."""
    A = np.atleast_2d(np.asarray(A, dtype=np.float64))
    n = A.shape[0]
    B = np.asarray(B, dtype=np.float64).reshape(n, -1)
    m = B.shape[1]
    eng = _rt.problem_for(A, B, Q, R, scratch=True)
    out = eng.dlqr_batch(S=1)
    K = out["K"].cpu().numpy()[:, 0].reshape(m, n)
    P = out["P"].cpu().numpy()[:, 0].reshape(n, n)
    return K, P, None
