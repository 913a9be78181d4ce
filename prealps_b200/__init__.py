"""prealps_b200 -- B200-native ECG + block-Jacobi (the preAlps hot path).

The product is two native libraries built in-tree by `make` (see __graft_entry__.build):
  prealps_b200/lib/libprealps_cuda.so   hand-written sm_100a kernels behind a C ABI (include/prealps_cuda.h)
  prealps_b200/lib/libprealps_b200.so   the preAlps API in C (include/operator.h, block_jacobi.h, ecg.h)
This Python package is only the ctypes binding used by tests/ and bench.py.  There is no
Python or CPU implementation of the hot path: importing `capi` fails loudly if the libraries
are missing, and every compute entry point aborts without a CUDA device.
"""
from . import capi  # noqa: F401
