"""python-mpc_b200 — B200-native batched linear-MPC QP path (drop-in for the OSQP hot path of
hynkis/Python-MPC).  Import as ``python_mpc_b200`` (the directory name has a hyphen).

Layout: csrc/ (hand-written sm_100a CUDA kernels + the C ABI of include/mpc_b200.h),
_lib.py (ctypes binding), solver.py (batched OSQP-like solver object), vehicle_models.py and
mpc.py (host-side mirror of the reference's vehicle models and MPC entry points).
"""
from ._lib import MpcError, cuda_backend, LIB_PATH, EXPORTED_SYMBOLS  # noqa: F401
from .solver import BatchSolver, SolveInfo, OSQP_DEFAULTS  # noqa: F401
