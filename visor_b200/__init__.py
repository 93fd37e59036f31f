"""visor_b200 — B200-native draw-execution path for visor (CUDA kernels + C-ABI).

The product is the shared library ``libvisor_b200.so`` built from ``csrc/`` (see ``Makefile``); its
C-ABI is declared in ``include/visor_b200.h``.  This module only locates/loads it.  There is no
Python or CPU implementation of the path: if the library is missing or no CUDA device is usable,
loading/initialising fails loudly.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvisor_b200.so")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include", "visor_b200.h")


def lib() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is not built: run `make -C visor_b200` "
                           "(or python -c 'import __graft_entry__ as g; g.build()')")
    return ctypes.CDLL(LIB_PATH)
