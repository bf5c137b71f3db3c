"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes front-end to oracle/_build/liboracle.so (the plain-C restatement of the reference's list build,
oracle/nlist_oracle.c) and to oracle/_ref/libref_*.so (the reference's own classes compiled from /root/reference
by oracle/Makefile).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package md_neighbor_list_b200 never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REFERENCE_SRC = "/root/reference"

_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_vp = C.c_void_p


def build(with_ref: bool | None = None) -> None:
    """Compile the oracle (and, where /root/reference exists, the reference classes).  Building is not using."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if with_ref is None:
        with_ref = os.path.isdir(REFERENCE_SRC)
    if with_ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build(with_ref=False)
        L = C.CDLL(LIB_PATH)
        L.orc_gen_fcc.restype = C.c_int64
        L.orc_gen_fcc.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_uint32, _vp, C.c_int,
                                  C.c_int64]
        L.orc_gen_uniform.restype = C.c_int64
        L.orc_gen_uniform.argtypes = [C.c_int64, C.c_double, C.c_uint64, _vp, C.c_int]
        L.orc_fnv1a64.restype = C.c_uint64
        L.orc_fnv1a64.argtypes = [_vp, C.c_int64]
        L.orc_sort_rows.restype = None
        L.orc_sort_rows.argtypes = [_vp, _vp, C.c_int64]
        L.orc_ell_from_csr.restype = C.c_int
        L.orc_ell_from_csr.argtypes = [_vp, _vp, C.c_int64, C.c_int32, _vp]
        L.orc_free.restype = None
        L.orc_free.argtypes = [_vp]
        for suf in ("_f64", "_f32"):
            f = getattr(L, "orc_build_half" + suf)
            f.restype = C.c_int64
            f.argtypes = [_vp, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                          _vp, _vp, C.POINTER(_vp), _i64p]
            f = getattr(L, "orc_build_full" + suf)
            f.restype = C.c_int64
            f.argtypes = [_vp, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                          _vp, _vp, C.POINTER(_vp), _i64p]
            f = getattr(L, "orc_bruteforce" + suf)
            f.restype = C.c_int64
            f.argtypes = [_vp, C.c_int64, C.c_int, C.c_double, C.c_int, C.c_int, _vp, _vp, C.POINTER(_vp)]
            f = getattr(L, "orc_bin" + suf)
            f.restype = C.c_int
            f.argtypes = [_vp, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                          _vp, _vp, _vp]
            f = getattr(L, "orc_band_report" + suf)
            f.restype = C.c_int64
            f.argtypes = [_vp, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                          _i64p, _i64p, _vp, C.c_int64]
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_vp)


def _suf(q: np.ndarray) -> str:
    if q.dtype == np.float64:
        return "_f64"
    if q.dtype == np.float32:
        return "_f32"
    raise TypeError(f"positions must be float32/float64, got {q.dtype}")


@dataclass
class CSR:
    """number_of_partners[n] int32, offsets[n+1] int64, partners[P] int32 (+ candidates tested)."""
    number_of_partners: np.ndarray
    offsets: np.ndarray
    partners: np.ndarray
    candidates: int = 0

    @property
    def number_of_pairs(self) -> int:
        return int(self.offsets[-1])

    def sorted_rows(self) -> "CSR":
        out = self.partners.copy()
        lib().orc_sort_rows(_ptr(out), _ptr(self.offsets), len(self.number_of_partners))
        return CSR(self.number_of_partners, self.offsets, out, self.candidates)


def fnv1a64(a: np.ndarray) -> str:
    a = np.ascontiguousarray(a)
    return f"{lib().orc_fnv1a64(_ptr(a), a.nbytes):016x}"


def gen_fcc(density: float, L: float = 50.0, sx: int = 0, sy: int = 0, sz: int = 0, seed: int = 2,
            stride: int = 4) -> np.ndarray:
    """make_list.cpp:51-77 workload (FCC + U[0,0.1) jitter from mt19937(seed)); returns (n, stride) float64."""
    n = lib().orc_gen_fcc(density, L, sx, sy, sz, seed, None, stride, 0)
    q = np.zeros((n, stride), dtype=np.float64)
    got = lib().orc_gen_fcc(density, L, sx, sy, sz, seed, _ptr(q), stride, n)
    assert got == n
    return q


def gen_uniform(n: int, L: float, seed: int = 2, stride: int = 4) -> np.ndarray:
    """U[0,L)^3 from mt19937_64(seed)."""
    q = np.zeros((n, stride), dtype=np.float64)
    lib().orc_gen_uniform(n, L, seed, _ptr(q), stride)
    return q


def _take_list(p: _vp, total: int) -> np.ndarray:
    if total > 0:
        arr = np.ctypeslib.as_array(C.cast(p, _i32p), shape=(total,)).copy()
    else:
        arr = np.zeros(0, dtype=np.int32)
    lib().orc_free(p)
    return arr


def _build(kind: str, q: np.ndarray, sl: float, box, order: int) -> CSR:
    q = np.ascontiguousarray(q)
    n, stride = q.shape
    np_ = np.zeros(n, dtype=np.int32)
    off = np.zeros(n + 1, dtype=np.int64)
    lp = _vp()
    cand = C.c_int64(0)
    f = getattr(lib(), f"orc_build_{kind}" + _suf(q))
    total = f(_ptr(q), n, stride, sl, box[0], box[1], box[2], order, _ptr(np_), _ptr(off), C.byref(lp),
              C.byref(cand))
    if total < 0:
        raise RuntimeError(f"oracle build_{kind} failed: {total}")
    return CSR(np_, off, _take_list(lp, total), cand.value)


def build_half(q, sl, box, order: int = 0) -> CSR:
    """neighlist_cpu.hpp:417-435 semantics (half list, key = min(i,j)), CSR in discovery order."""
    return _build("half", q, sl, box, order)


def build_full(q, sl, box, order: int = 0) -> CSR:
    """kernel_impl.cuh:3-35 semantics (every j != i within SL), CSR in discovery order."""
    return _build("full", q, sl, box, order)


def bruteforce(q, sl, full: bool, order: int = 0) -> CSR:
    """make_list.cpp:79-99 (half) / make_list.cu:79-98 (full)."""
    q = np.ascontiguousarray(q)
    n, stride = q.shape
    np_ = np.zeros(n, dtype=np.int32)
    off = np.zeros(n + 1, dtype=np.int64)
    lp = _vp()
    f = getattr(lib(), "orc_bruteforce" + _suf(q))
    total = f(_ptr(q), n, stride, sl, int(full), order, _ptr(np_), _ptr(off), C.byref(lp))
    if total < 0:
        raise RuntimeError("oracle bruteforce failed")
    return CSR(np_, off, _take_list(lp, total))


def mesh_dims(sl, box, dtype=np.float64):
    """neighlist_cpu.hpp:384-387: mesh_size = int(L / search_length) in the working precision."""
    t = np.dtype(dtype).type
    return [int(t(box[d]) / t(sl)) for d in range(3)]


def bin_particles(q, sl, box, gpu_clamp: bool = False):
    """neighlist_cpu.hpp:134-165: returns (mesh_index[M+1] int64, ptcl_id_in_mesh[n] int32, cell_of[n] int32)."""
    q = np.ascontiguousarray(q)
    n, stride = q.shape
    mesh = mesh_dims(sl, box, q.dtype)
    M = mesh[0] * mesh[1] * mesh[2]
    mi = np.zeros(M + 1, dtype=np.int64)
    pid = np.zeros(max(n, 1), dtype=np.int32)
    cell = np.zeros(max(n, 1), dtype=np.int32)
    f = getattr(lib(), "orc_bin" + _suf(q))
    rc = f(_ptr(q), n, stride, sl, box[0], box[1], box[2], int(gpu_clamp), _ptr(mi), _ptr(pid), _ptr(cell))
    if rc:
        raise RuntimeError(f"oracle bin failed: {rc}")
    return mi, pid[:n], cell[:n]


def band_report(q, sl, box, cap: int = 1024):
    """Pairs whose verdict depends on the rounding order / lies within 1 ulp of SL^2 (north_star's '1-ulp band')."""
    q = np.ascontiguousarray(q)
    n, stride = q.shape
    nod, nulp = C.c_int64(0), C.c_int64(0)
    pairs = np.zeros((cap, 2), dtype=np.int32)
    f = getattr(lib(), "orc_band_report" + _suf(q))
    nb = f(_ptr(q), n, stride, sl, box[0], box[1], box[2], C.byref(nod), C.byref(nulp), _ptr(pairs), cap)
    if nb < 0:
        raise RuntimeError("oracle band_report failed")
    return {"order_dependent": nod.value, "within_1ulp": nulp.value, "pairs": pairs[:nb].copy()}


def ell_from_csr(csr: CSR, rows: int) -> np.ndarray:
    """kernel_impl.cuh:30 layout list[k*N+i], padded -1."""
    n = len(csr.number_of_partners)
    ell = np.empty((rows, n), dtype=np.int32)
    rc = lib().orc_ell_from_csr(_ptr(csr.partners), _ptr(csr.offsets), n, rows, _ptr(ell))
    if rc:
        raise RuntimeError("row longer than ELL capacity")
    return ell


# ---------------------------------------------------------------------------------------------------------------
# The reference's own classes (oracle/_ref/libref_*.so)
# ---------------------------------------------------------------------------------------------------------------
REF_VARIANTS = {"scalar": "libref_scalar.so", "scalar_swp": "libref_scalar_swp.so",
                "avx2_4x1": "libref_avx2.so", "avx512_8x1": "libref_avx512.so"}


def _cpu_flags() -> set:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def ref_available(variant: str) -> bool:
    if not os.path.exists(os.path.join(REF_DIR, REF_VARIANTS[variant])):
        return False
    flags = _cpu_flags()
    need = {"avx2", "fma"}
    if variant in ("avx2_4x1", "avx512_8x1"):
        need |= {"avx512f", "avx512dq", "avx512bw", "avx512vl", "avx512cd"}
    return need <= flags


_ref_libs: dict = {}


def ref_build(variant: str, q: np.ndarray, sl: float, box, loops: int = 1, warmup: int = 0):
    """Run the reference class `variant` on q (n x 4 float64): `warmup` untimed builds, then `loops` timed ones on the
    same instance (the reference's own protocol, make_list.cpp:152-157).  Returns (CSR with int64 offsets,
    ms_per_build)."""
    if variant not in _ref_libs:
        L = C.CDLL(os.path.join(REF_DIR, REF_VARIANTS[variant]))
        L.ref_build_warm.restype = C.c_int
        L.ref_build_warm.argtypes = [_vp, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int32,
                                     C.c_int32, _vp, _vp, _vp, C.c_int64, _i64p, C.POINTER(C.c_double)]
        L.ref_pair_capacity.restype = C.c_int64
        L.ref_pair_capacity.argtypes = [C.c_int64]
        _ref_libs[variant] = L
    L = _ref_libs[variant]
    q = np.ascontiguousarray(q, dtype=np.float64)
    n, stride = q.shape
    assert stride == 4
    cap = L.ref_pair_capacity(n)
    np_ = np.zeros(n, dtype=np.int32)
    kp = np.zeros(n + 1, dtype=np.int32)
    lst = np.zeros(cap, dtype=np.int32)
    npairs = C.c_int64(0)
    ms = C.c_double(0)
    rc = L.ref_build_warm(_ptr(q), n, sl, box[0], box[1], box[2], warmup, loops, _ptr(np_), _ptr(kp), _ptr(lst),
                          cap, C.byref(npairs), C.byref(ms))
    if rc:
        raise RuntimeError(f"reference {variant} failed: {rc}")
    return CSR(np_, kp.astype(np.int64), lst[: npairs.value].copy()), ms.value
