"""Slab decomposition of the list build over the GPUs of one box (no reference counterpart — the
reference is single-process, single-GPU).

One process per GPU.  The box is cut into `world` slabs of equal thickness along one axis; rank r owns the particles
whose coordinate on that axis lies in [lo, hi).  A list row depends only on the particles within the search length of
its owner, so the only exchange is the ghost layer: before a build every rank sends the owned particles within
`search_length` of a face to the rank on the other side of that face ({x, y, z, w} records + global ids, one grouped
send/recv per face over NCCL — NVLink 5 / NVSwitch on a B200 box), then builds rows for its owned particles only:

    q_all      = [ owned | ghosts from below | ghosts from above ]
    global_ids = [ base + arange(n_owned) | received ids ]
    nlb200_build_subset(handle, q_all, n_total, n_owned, global_ids)      (include/nlist_b200.h)

Every rank bins on the GLOBAL cell grid (the handle is created with the global box), so the rows it emits are exactly
the rows a single-GPU build of the whole system would emit for those particles — same partners, same order.  HALF
lists: the row of the smaller global id keeps the pair, so each pair is emitted once, by the rank that owns that
particle (ghosts are needed from both faces).  Open boundary (the reference measures distances without minimum image,
neighlist_cpu.hpp:219-223): the end slabs have one neighbour.

Device path: selection and gather run in the library's CUDA kernels (nlb200_select_slab / nlb200_gather_records).
The same partition / exchange / ownership logic also accepts CPU tensors (torch ops + the `gloo` backend); that
branch exists so that the world_size-2 tests can exercise the logic without GPUs — it is not a compute fallback: the
list build itself is injected by the caller (`build_fn`) and is the CUDA library in the product.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


class SlabDecomposition:
    def __init__(self, world: int, rank: int, box, search_length: float, axis: int = 2, stride: int = 4,
                 group=None):
        if not (0 <= rank < world):
            raise ValueError("rank outside [0, world)")
        self.world, self.rank, self.axis, self.stride = int(world), int(rank), int(axis), int(stride)
        self.box = tuple(float(b) for b in box)
        self.sl = float(search_length)
        self.thickness = self.box[axis] / world
        if world > 1 and self.thickness < self.sl:
            raise ValueError("slab thinner than the search length: ghosts would come from second neighbours")
        self.lo = rank * self.thickness
        self.hi = (rank + 1) * self.thickness if rank + 1 < world else float("inf")
        if rank == 0:
            self.lo = -float("inf")
        self.group = group
        self._buf = None  # device workspaces, allocated on first build
        self._qall = self._gall = None
        self._last = None
        self.last_counts = (0, 0, 0, 0)  # sent below, sent above, received from below, received from above

    # -- partitioning ---------------------------------------------------------------------------------------------
    def owns(self, q: np.ndarray) -> np.ndarray:
        """Boolean mask of the particles of a global array that this rank owns."""
        z = q[:, self.axis]
        return (z >= self.lo) & (z < self.hi)

    def partition(self, q_global: np.ndarray):
        """(owned positions, their global ids) of this rank, ids ascending."""
        m = self.owns(q_global)
        return np.ascontiguousarray(q_global[m]), np.nonzero(m)[0].astype(np.int32)

    def local_fcc_slab(self, density: float, L: float, seed: int = 2):
        """Weak-scaling workload: every rank owns one L^3 block of the reference's jittered FCC system
        (make_list.cpp:51-77), shifted to its slab; blocks are generated independently (seed + rank), so the global
        system is their concatenation and global ids are rank * n + local index."""
        from . import workloads
        q = workloads.fcc(density, L, seed=seed + self.rank, stride=self.stride)
        q[:, self.axis] += self.rank * self.thickness
        return q

    def max_ghosts(self, n_owned: int) -> int:
        if self.world == 1:
            return 0
        frac = min(1.0, self.sl / self.thickness)
        return int(2 * n_owned * frac * 1.5) + 4096

    # -- exchange -------------------------------------------------------------------------------------------------
    def _select(self, q: torch.Tensor, lo: float, hi: float) -> torch.Tensor:
        """Indices (ascending) of the rows of q with lo <= q[:, axis] < hi."""
        n = q.shape[0]
        if not q.is_cuda:
            z = q[:, self.axis]
            return torch.nonzero((z >= lo) & (z < hi)).flatten().to(torch.int32)
        L = _lib.lib()
        ws_bytes = L.nlb200_select_slab_workspace(n)
        b = self._buf
        if b is None or b["ws"].numel() < ws_bytes or b["idx"].numel() < n:
            self._buf = b = {"ws": torch.empty(ws_bytes, dtype=torch.uint8, device=q.device),
                             "idx": torch.empty(max(n, 1), dtype=torch.int32, device=q.device),
                             "cnt": torch.zeros(1, dtype=torch.int64, device=q.device)}
        dtype = _lib.F64 if q.dtype == torch.float64 else _lib.F32
        s = torch.cuda.current_stream().cuda_stream
        st = L.nlb200_select_slab(q.data_ptr(), n, dtype, self.stride, self.axis, lo, hi, b["idx"].data_ptr(), n,
                                  b["cnt"].data_ptr(), b["ws"].data_ptr(), ws_bytes, s)
        if st != _lib.OK:
            raise _lib.NlistError(st, "nlb200_select_slab failed")
        cnt = int(b["cnt"].item())  # host needs the count to size the send (one small D2H per face)
        return b["idx"][:cnt].clone()

    def _gather(self, q: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        if not q.is_cuda:
            return q[idx.long()].contiguous()
        out = torch.empty((idx.numel(), self.stride), dtype=q.dtype, device=q.device)
        if idx.numel():
            dtype = _lib.F64 if q.dtype == torch.float64 else _lib.F32
            st = _lib.lib().nlb200_gather_records(q.data_ptr(), idx.data_ptr(), idx.numel(), dtype, self.stride,
                                                  out.data_ptr(), torch.cuda.current_stream().cuda_stream)
            if st != _lib.OK:
                raise _lib.NlistError(st, "nlb200_gather_records failed")
        return out

    def exchange(self, q_owned: torch.Tensor, gid_owned: torch.Tensor):
        """Returns (q_all, gid_all, n_owned): owned records followed by the ghosts from below and from above."""
        n = q_owned.shape[0]
        if self.world == 1:
            return q_owned, gid_owned, n
        dev = q_owned.device
        below, above = self.rank - 1, self.rank + 1
        send = {}
        if below >= 0:
            idx = self._select(q_owned, -float("inf"), self.lo + self.sl)
            send[below] = (self._gather(q_owned, idx), gid_owned[idx.long()].contiguous())
        if above < self.world:
            idx = self._select(q_owned, self.hi - self.sl, float("inf"))
            send[above] = (self._gather(q_owned, idx), gid_owned[idx.long()].contiguous())
        # 1. counts (one int64 per face), 2. records + ids — each a single grouped send/recv
        cnt_out = {p: torch.tensor([send[p][0].shape[0]], dtype=torch.int64, device=dev) for p in send}
        cnt_in = {p: torch.zeros(1, dtype=torch.int64, device=dev) for p in send}
        ops = []
        for p in sorted(send):
            ops.append(dist.P2POp(dist.isend, cnt_out[p], p, group=self.group))
            ops.append(dist.P2POp(dist.irecv, cnt_in[p], p, group=self.group))
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        # persistent assembly buffers (stable device pointers: identical builds replay the library's CUDA graph);
        # ghosts are received straight into their final place behind the owned records
        n_in = {p: int(cnt_in[p].item()) for p in send}
        n_total = n + sum(n_in.values())
        if (self._qall is None or self._qall.shape[0] < n_total or self._qall.device != dev
                or self._qall.dtype != q_owned.dtype):
            cap = max(n_total, n + self.max_ghosts(n))
            self._qall = torch.empty((cap, self.stride), dtype=q_owned.dtype, device=dev)
            self._gall = torch.empty(cap, dtype=torch.int32, device=dev)
        self._qall[:n].copy_(q_owned)
        self._gall[:n].copy_(gid_owned)
        recv, at = {}, n
        for p in (below, above):
            if p in send:
                recv[p] = (self._qall[at:at + n_in[p]], self._gall[at:at + n_in[p]])
                at += n_in[p]
        ops = []
        for p in sorted(send):
            for k in (0, 1):
                if send[p][k].numel():
                    ops.append(dist.P2POp(dist.isend, send[p][k], p, group=self.group))
                if recv[p][k].numel():
                    ops.append(dist.P2POp(dist.irecv, recv[p][k], p, group=self.group))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        self.last_counts = (send[below][0].shape[0] if below in send else 0,
                            send[above][0].shape[0] if above in send else 0,
                            n_in.get(below, 0), n_in.get(above, 0))
        self._last = (self._qall[:n_total], self._gall[:n_total], n)
        return self._last

    def last_assembled(self):
        """(q_all, gid_all, n_owned) of the last exchange — lets one rank rebuild (e.g. under a profiler) without a new
        exchange, which would need every rank."""
        return self._last

    # -- build ----------------------------------------------------------------------------------------------------
    def global_ids(self, n_owned: int, device) -> torch.Tensor:
        """Global ids of equally sized slabs (local_fcc_slab): rank * n + local index."""
        return torch.arange(n_owned, dtype=torch.int32, device=device) + self.rank * n_owned

    def build(self, nl, q_owned: torch.Tensor, stream=None, gid_owned: torch.Tensor | None = None, build_fn=None):
        """Ghost exchange followed by the list build of the owned rows.  `nl` is a VerletListB200 created with the
        GLOBAL box; `build_fn(q_all, n_owned, gid_all)` replaces nl.build in the CPU logic tests."""
        ctx = torch.cuda.stream(stream) if (stream is not None and q_owned.is_cuda) else _null()
        with ctx:
            if gid_owned is None:
                gid_owned = self.global_ids(q_owned.shape[0], q_owned.device)
            q_all, gid_all, n_owned = self.exchange(q_owned, gid_owned)
            if build_fn is not None:
                return build_fn(q_all, n_owned, gid_all)
            nl.build(q_all, n_owned=n_owned, global_ids=gid_all if self.world > 1 else None, stream=stream)
        return None


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
