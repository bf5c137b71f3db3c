"""Input generators of the drivers (host side), backed by csrc/workloads.cpp.

`fcc` is the reference drivers' `init()` (make_list.cpp:51-77, make_list.cu:42-66); `uniform` and `clustered` are the
synthetic distributions of SURVEY.md §8d.  They produce numpy float64 arrays of shape (n, stride).
"""
from __future__ import annotations

import numpy as np

from . import _lib


def fcc(density: float, L: float = 50.0, sx: int = 0, sy: int = 0, sz: int = 0, seed: int = 2,
        stride: int = 4) -> np.ndarray:
    lib = _lib.lib()
    n = lib.nlb200_workload_fcc(density, L, sx, sy, sz, seed, None, stride, 0)
    if n < 0:
        raise ValueError("bad FCC workload parameters")
    q = np.zeros((n, stride), dtype=np.float64)
    got = lib.nlb200_workload_fcc(density, L, sx, sy, sz, seed, q.ctypes.data, stride, n)
    assert got == n
    return q


def uniform(n: int, L: float, seed: int = 2, stride: int = 4) -> np.ndarray:
    q = np.zeros((n, stride), dtype=np.float64)
    if _lib.lib().nlb200_workload_uniform(n, L, seed, q.ctypes.data, stride) != n:
        raise ValueError("bad uniform workload parameters")
    return q


def clustered(n: int, L: float, blobs: int = 32, seed: int = 2, stride: int = 4) -> np.ndarray:
    q = np.zeros((n, stride), dtype=np.float64)
    if _lib.lib().nlb200_workload_clustered(n, L, blobs, seed, q.ctypes.data, stride) != n:
        raise ValueError("bad clustered workload parameters")
    return q
