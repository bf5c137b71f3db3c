"""Host-side mirror of the reference's list-builder classes on top of the C ABI (include/nlist_b200.h).

Reference interface being mirrored:
  NeighListGPU<Vec,Dtype>(search_length, Lx, Ly, Lz); Initialize(N); MakeNeighList(q, N, sync, ...);
      number_of_pairs(); neigh_list(); number_of_partners()                      (neighlist_gpu.hpp:236-487)
  NeighList*<Vec>(search_length, Lx, Ly, Lz); Initialize(N); MakeNeighList(q, N);
      number_of_pairs(); sorted_list(); key_pointer(); number_of_partners()      (neighlist_cpu.hpp:380-463)

PyTorch is used only for device memory and streams; all computation happens in libnlist_b200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import (F32, F64, FULL_CSR, FULL_ELL_TRANSPOSED, HALF_CSR, OPT_ELL_ROWS, OPT_EXACT_ONLY,
                   OPT_KERNEL_VARIANT, OPT_MAX_IN_CELL, OPT_PDL, OPT_POSITION_STRIDE, OPT_PROFILE, OPT_SORT_ROWS,
                   OPT_USE_GRAPH, NlistError,
                   Stats, check)

_MODES = {"half_csr": HALF_CSR, "full_csr": FULL_CSR, "full_ell_transposed": FULL_ELL_TRANSPOSED}
_DTYPES = {"f64": F64, "f32": F32, torch.float64: F64, torch.float32: F32, np.float64: F64, np.float32: F32}


class _DevView:
    """Zero-copy torch view of a library-owned device buffer (via __cuda_array_interface__)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def _view(ptr, n: int, typestr: str, device) -> torch.Tensor:
    if n == 0 or not ptr:
        dt = {"<i4": torch.int32, "<i8": torch.int64}[typestr]
        return torch.empty(0, dtype=dt, device=device)
    return torch.as_tensor(_DevView(int(ptr), int(n), typestr), device=device)


class VerletListB200:
    """The B200 builder: one handle of libnlist_b200.so."""

    def __init__(self, search_length: float, Lx: float, Ly: float, Lz: float, dtype="f64", mode="full_csr",
                 position_stride: int = 4, sort_rows: bool = False, ell_rows: int = 200, exact_only: bool = False,
                 use_graph: bool = True, kernel_variant: int = 0, profile: bool = False, max_in_cell: int = 0,
                 cell_window=None, pdl: bool | None = None):
        self._lib = _lib.lib()
        self._h = C.c_void_p()
        self.dtype = _DTYPES[dtype]
        self.mode = _MODES[mode] if isinstance(mode, str) else int(mode)
        self.stride = int(position_stride)
        self.ell_rows = int(ell_rows)
        st = self._lib.nlb200_create(search_length, Lx, Ly, Lz, self.dtype, self.mode, C.byref(self._h))
        if st != _lib.OK:
            raise NlistError(st, "nlb200_create: invalid box / search length (need >= 3 cells per axis)")
        for opt, val in ((OPT_POSITION_STRIDE, self.stride), (OPT_SORT_ROWS, int(sort_rows)),
                         (OPT_ELL_ROWS, self.ell_rows), (OPT_EXACT_ONLY, int(exact_only)),
                         (OPT_USE_GRAPH, int(use_graph)), (OPT_KERNEL_VARIANT, int(kernel_variant)),
                         (OPT_PROFILE, int(profile)), (OPT_MAX_IN_CELL, int(max_in_cell))):
            check(self._h, self._lib.nlb200_set_option(self._h, opt, val))
        if pdl is not None:
            check(self._h, self._lib.nlb200_set_option(self._h, OPT_PDL, int(pdl)))
        if cell_window is not None:
            # (axis, first_cell, n_cells): this handle bins only that window of the global grid (a slab rank)
            axis, first, count = cell_window
            check(self._h, self._lib.nlb200_set_cell_window(self._h, int(axis), int(first), int(count)))
        self.n = 0
        self.device = None
        self._q_keepalive = None

    # -- lifecycle ------------------------------------------------------------------------------------------------
    def initialize(self, max_particles: int, max_entries: int = 0) -> None:
        check(self._h, self._lib.nlb200_initialize(self._h, int(max_particles), int(max_entries)))
        self.device = torch.device("cuda", torch.cuda.current_device())

    def reserve(self, max_entries: int) -> None:
        check(self._h, self._lib.nlb200_reserve(self._h, int(max_entries)))

    def reserve_cell_capacity(self, max_in_cell: int) -> None:
        check(self._h, self._lib.nlb200_reserve_cell_capacity(self._h, int(max_in_cell)))

    def close(self) -> None:
        if self._h:
            self._lib.nlb200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- the hot path ---------------------------------------------------------------------------------------------
    def build(self, q: torch.Tensor, n: int | None = None, stream: torch.cuda.Stream | None = None,
              n_owned: int | None = None, global_ids: torch.Tensor | None = None) -> None:
        """Asynchronous build from a CUDA tensor of shape (n, stride)."""
        if not q.is_cuda or not q.is_contiguous():
            raise ValueError("q must be a contiguous CUDA tensor")
        if _DTYPES[q.dtype] != self.dtype:
            raise ValueError("q dtype does not match the handle's dtype")
        if q.dim() != 2 or q.shape[1] != self.stride:
            raise ValueError(f"q must have shape (n, {self.stride})")
        n = q.shape[0] if n is None else int(n)
        s = stream if stream is not None else torch.cuda.current_stream()
        self._q_keepalive = (q, global_ids)
        if n_owned is None and global_ids is None:
            check(self._h, self._lib.nlb200_build(self._h, q.data_ptr(), n, s.cuda_stream))
            self.n = n
        else:
            own = n if n_owned is None else int(n_owned)
            gid = 0
            if global_ids is not None:
                if global_ids.dtype != torch.int32 or not global_ids.is_cuda or global_ids.numel() < n:
                    raise ValueError("global_ids must be an int32 CUDA tensor with n entries")
                gid = global_ids.data_ptr()
            check(self._h, self._lib.nlb200_build_subset(self._h, q.data_ptr(), n, own, gid, s.cuda_stream))
            self.n = own

    def _mark_pending(self, stream: torch.cuda.Stream) -> None:
        """A build captured in an outer CUDA graph was replayed on `stream` (nlb200_mark_enqueued)."""
        check(self._h, self._lib.nlb200_mark_enqueued(self._h, stream.cuda_stream))

    def synchronize(self) -> Stats:
        check(self._h, self._lib.nlb200_synchronize(self._h))
        return self.stats()

    def build_host(self, q: np.ndarray, want_partners: bool = True):
        """Host buffers in, host buffers out (H2D + build + D2H inside the call).  Returns (np, offsets, partners)."""
        q = np.ascontiguousarray(q)
        if _DTYPES[q.dtype.type] != self.dtype or q.ndim != 2 or q.shape[1] != self.stride:
            raise ValueError("bad host position array")
        n = q.shape[0]
        npart = np.empty(n, dtype=np.int32)
        off = np.empty(n + 1, dtype=np.int64)
        total = C.c_int64(0)
        # first call learns the size (counts + offsets only), second fetches the list
        check(self._h, self._lib.nlb200_build_host(self._h, q.ctypes.data, n, npart.ctypes.data, off.ctypes.data,
                                                  None, 0, C.byref(total)))
        self.n = n
        if not want_partners:
            return npart, off, None
        lst = np.empty(total.value, dtype=np.int32)
        return npart, off, self.copy_partners_to(lst)

    def copy_partners_to(self, out: np.ndarray) -> np.ndarray:
        total = self.number_of_pairs()
        t = self.partners()
        if total:
            out_t = torch.from_numpy(out[:total])
            out_t.copy_(t)
        return out[:total]

    # -- callers either side of the build -----------------------------------------------------------
    def track(self, q: torch.Tensor, stream: torch.cuda.Stream | None = None) -> None:
        """Remember the positions the current list was built from (Verlet-list lifetime, nlb200_track_reference)."""
        s = stream if stream is not None else torch.cuda.current_stream()
        check(self._h, self._lib.nlb200_track_reference(self._h, q.data_ptr(), q.shape[0], s.cuda_stream))

    def max_displacement(self, q: torch.Tensor, stream: torch.cuda.Stream | None = None) -> float:
        """max_i |q[i] - q_tracked[i]|: rebuild when it exceeds margin / 2."""
        s = stream if stream is not None else torch.cuda.current_stream()
        out = C.c_double(0.0)
        check(self._h, self._lib.nlb200_max_displacement(self._h, q.data_ptr(), q.shape[0], s.cuda_stream,
                                                         C.byref(out)))
        return float(out.value)

    def lj_forces(self, q: torch.Tensor, rc: float, epsilon: float = 1.0, sigma: float = 1.0,
                  stream: torch.cuda.Stream | None = None):
        """(forces (n, 3), per-particle energies (n,)) of a Lennard-Jones fluid over the FULL rows of the last build."""
        s = stream if stream is not None else torch.cuda.current_stream()
        f = torch.empty((self.n, 3), dtype=torch.float64, device=q.device)
        e = torch.empty(self.n, dtype=torch.float64, device=q.device)
        check(self._h, self._lib.nlb200_lj_forces(self._h, q.data_ptr(), rc, epsilon, sigma, f.data_ptr(),
                                                  e.data_ptr(), s.cuda_stream))
        return f, e

    def gather_sorted(self, src: torch.Tensor, stream: torch.cuda.Stream | None = None) -> torch.Tensor:
        """src[sorted_ids] for a per-particle CUDA array (n, width) of 4- or 8-byte elements: the cell-ordered copy
        the reference stubbed out (SortPtclData / CopyGather)."""
        s = stream if stream is not None else torch.cuda.current_stream()
        src = src.contiguous()
        width = 1 if src.dim() == 1 else src.shape[1]
        dst = torch.empty_like(src)
        check(self._h, self._lib.nlb200_gather_sorted(self._h, src.data_ptr(), src.element_size(), width,
                                                      dst.data_ptr(), s.cuda_stream))
        return dst

    # -- accessors ------------------------------------------------------------------------------------------------
    def stats(self) -> Stats:
        s = Stats()
        check(self._h, self._lib.nlb200_get_stats(self._h, C.byref(s)))
        return s

    def stage_times(self) -> dict:
        """Per-stage device milliseconds of the last synchronized build (profile=True handles only)."""
        ms = (C.c_float * 16)()
        ids = (C.c_int32 * 16)()
        k = self._lib.nlb200_get_stage_times(self._h, ms, ids, 16)
        if k < 0:
            raise NlistError(_lib.ERR_STATE, "stage times need profile=True and a synchronized build")
        return {self._lib.nlb200_stage_name(ids[i]).decode(): float(ms[i]) for i in range(k)}

    def number_of_pairs(self) -> int:
        return int(self._lib.nlb200_number_of_pairs(self._h))

    def number_of_partners(self) -> torch.Tensor:
        return _view(self._lib.nlb200_number_of_partners(self._h), self.n, "<i4", self.device)

    def offsets(self) -> torch.Tensor:
        return _view(self._lib.nlb200_offsets(self._h), self.n + 1, "<i8", self.device)

    def offsets32(self) -> torch.Tensor:
        p = self._lib.nlb200_offsets32(self._h)
        if not p:
            raise NlistError(_lib.ERR_INVALID, "total exceeds INT32_MAX: no 32-bit offsets view")
        return _view(p, self.n + 1, "<i4", self.device)

    def partners(self) -> torch.Tensor:
        return _view(self._lib.nlb200_partners(self._h), max(self.number_of_pairs(), 0), "<i4", self.device)

    def ell_transposed(self) -> torch.Tensor:
        p = self._lib.nlb200_ell_transposed(self._h)
        if not p:
            raise NlistError(_lib.ERR_STATE, "handle was not created in full_ell_transposed mode")
        return _view(p, self.ell_rows * self.n, "<i4", self.device).view(self.ell_rows, self.n)

    def cell_start(self) -> torch.Tensor:
        m = self.stats().mesh
        return _view(self._lib.nlb200_cell_start(self._h), m[0] * m[1] * m[2] + 1, "<i4", self.device)

    def sorted_ids(self, n_total: int | None = None) -> torch.Tensor:
        return _view(self._lib.nlb200_sorted_ids(self._h), self.n if n_total is None else n_total, "<i4", self.device)


class NeighListGPU:
    """Drop-in mirror of the reference's NeighListGPU (neighlist_gpu.hpp:43-488): same method names, same meaning.

    neigh_list() returns the reference layout list[k*N + i] (-1 padded, MAX_PARTNERS rows)."""

    MAX_PARTNERS = 200  # neighlist_gpu.hpp:70

    def __init__(self, search_length: float, Lx: float, Ly: float, Lz: float, dtype="f64"):
        self._impl = VerletListB200(search_length, Lx, Ly, Lz, dtype=dtype, mode="full_ell_transposed",
                                    ell_rows=self.MAX_PARTNERS)

    def Initialize(self, particle_number: int) -> None:
        self._impl.initialize(particle_number)

    def MakeNeighList(self, q: torch.Tensor, particle_number: int, sync: bool = True, tblock_size: int = 128,
                      smem_hei: int = 7) -> None:
        del tblock_size, smem_hei  # launch shapes are chosen by the library
        self._impl.build(q, particle_number)
        if sync:
            self._impl.synchronize()

    def number_of_pairs(self) -> int:
        self._impl.synchronize()
        return self._impl.number_of_pairs()

    def neigh_list(self) -> torch.Tensor:
        return self._impl.ell_transposed()

    def number_of_partners(self) -> torch.Tensor:
        return self._impl.number_of_partners()


class NeighList:
    """Drop-in mirror of the reference's CPU classes NeighList / NeighListAVX2 / NeighListAVX512
    (neighlist_cpu.hpp:380-463): host arrays in, half list in CSR out (key = smaller index)."""

    def __init__(self, search_length: float, Lx: float, Ly: float, Lz: float, position_stride: int = 4):
        self._impl = VerletListB200(search_length, Lx, Ly, Lz, dtype="f64", mode="half_csr",
                                    position_stride=position_stride)
        self._np = self._off = self._list = None

    def Initialize(self, particle_number: int) -> None:
        self._impl.initialize(particle_number)

    def MakeNeighList(self, q: np.ndarray, particle_number: int) -> None:
        self._np, self._off, self._list = self._impl.build_host(q[:particle_number])

    def number_of_pairs(self) -> int:
        return int(self._off[-1])

    def sorted_list(self) -> np.ndarray:
        return self._list

    def key_pointer(self) -> np.ndarray:
        """int32 like the reference's key_pointer_ when it fits, else int64."""
        return self._off.astype(np.int32) if self._off[-1] <= np.iinfo(np.int32).max else self._off

    def number_of_partners(self) -> np.ndarray:
        return self._np
