"""md_neighbor_list_b200 — B200-native Verlet neighbor-list builder (drop-in for kohnakagawa/md_neighbor_list's
list-build path).  The product is libnlist_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/nlist_b200.h); this package is the thin host-side mirror of the reference's class interface."""
from ._lib import (F32, F64, FULL_CSR, FULL_ELL_TRANSPOSED, HALF_CSR, LIB_PATH, NlistError, Stats, SYMBOLS)  # noqa: F401

__all__ = ["VerletListB200", "NeighListGPU", "NeighList", "PeriodicVerletList", "PeriodicSlabDecomposition", "workloads", "NlistError", "LIB_PATH",
           "SYMBOLS"]


def __getattr__(name):
    # torch is imported lazily so that `import md_neighbor_list_b200` stays cheap for symbol checks
    import importlib
    if name in ("VerletListB200", "NeighListGPU", "NeighList"):
        return getattr(importlib.import_module(__name__ + ".neighlist"), name)
    if name in ("PeriodicVerletList", "PeriodicSlabDecomposition"):
        return getattr(importlib.import_module(__name__ + ".periodic"), name)
    if name in ("workloads", "neighlist", "parallel", "periodic"):
        return importlib.import_module(__name__ + "." + name)
    raise AttributeError(name)
