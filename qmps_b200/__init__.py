"""qmps_b200 -- B200-native classical inner loop of qmps (fergusfinn/qmps).

Layout
  csrc/               hand-written sm_100a CUDA kernels + the C ABI (include/qmps_b200.h)
  _lib.py             ctypes binding of that ABI (no fallback: raises if not built / no GPU)
  batched.py          batched device API on torch CUDA tensors
  tools.py, time_evolve_tools.py, represent.py, ground_state.py, rotosolve.py,
  exact_loschmidt.py, loschmidts/   drop-in mirrors of the reference's qmps.* modules
  dist.py             batch sharding over ranks + the final NCCL argmin
"""
from ._lib import QmpsError, LIB_PATH  # noqa: F401

__version__ = "0.1.0"
