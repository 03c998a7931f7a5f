"""colvars_b200 -- B200-native coordination-number engine behind the Colvars CVC interface.

csrc/      hand-written sm_100a kernels + the C-ABI implementation (libb200cv.so)
host/      C++ `colvar::cvc` shim for the unmodified Colvars host
binding.py ctypes plumbing used by tests/ and bench.py
"""
from .binding import (CoordNum, B200cvError, load, comm_unique_id, plan_items, fp64_peak, selftest_rcp,
                      kernel_launches,
                      EF_NULL, EF_GRADIENTS, EF_USE_PAIRLIST, EF_REBUILD_PAIRLIST, LIB_PATH,
                      SYMBOLS)

__all__ = ["CoordNum", "B200cvError", "load", "comm_unique_id", "plan_items", "fp64_peak", "selftest_rcp",
           "kernel_launches",
           "EF_NULL", "EF_GRADIENTS", "EF_USE_PAIRLIST", "EF_REBUILD_PAIRLIST", "LIB_PATH",
           "SYMBOLS"]
