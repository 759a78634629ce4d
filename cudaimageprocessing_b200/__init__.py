"""cudaimageprocessing_b200 -- B200-native guided filter behind the CudaImageProcessing call
surface (GuidedFilter/guided_filter.h, guided_filter_d.h).

The product is the shared library `libgf_b200.so` (hand-written CUDA for sm_100a + a C ABI,
include/gf_b200.h).  This Python package is a thin harness over that ABI: it builds/loads the
library and mirrors the reference's host interface for tests and benchmarks.  There is no CPU
path: if the library cannot be built or loaded, importing `api()` raises.
"""
from __future__ import annotations

import ctypes
import os

from ._capi import (BORDER_REFLECT, BORDER_REFLECT101, BORDER_TRUNCATE, GfApi, GfError)

_API = None
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libgf_b200.so")


def api() -> GfApi:
    """Loads libgf_b200.so (building it with nvcc if it is missing or stale)."""
    global _API
    if _API is None:
        from .build import build
        # GF_LIB_PATH: developer switch for A/B-testing differently compiled builds of the SAME
        # sources (bench_tools/); it must still be a libgf_b200 build -- there is no other backend.
        path = os.environ.get("GF_LIB_PATH") or build()
        _API = GfApi(ctypes.CDLL(path))
    return _API


__all__ = ["api", "GfApi", "GfError", "BORDER_REFLECT101", "BORDER_TRUNCATE", "BORDER_REFLECT", "LIB_PATH"]
