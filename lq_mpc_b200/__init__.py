"""lq_mpc_b200 — B200-native batch engine for the hot path of lcrekko/lq_mpc (certainty-equivalent LQ MPC under
model mismatch). `utils` / `utils_class` mirror the reference's modules; `engine` is the ctypes binding of the C ABI
(include/lqmpc_b200.h); `stats`, `sweep`, `sampling` are the batch-level API. Nothing here computes on the CPU."""
__version__ = "0.1.0"
