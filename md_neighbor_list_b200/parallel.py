"""Slab decomposition of the list build over the GPUs of one box (no reference counterpart — the
reference is single-process, single-GPU).

One process per GPU.  The box is cut into `world` slabs of equal thickness along one axis; rank r owns the particles
whose coordinate on that axis lies in [lo, hi).  A list row depends only on the particles within the search length of
its owner, so the only exchange is the ghost layer: before a build every rank sends the owned particles within
`search_length` of a face to the rank on the other side of that face ({x, y, z, w} records + global ids, ONE grouped
send/recv over NCCL — NVLink 5 / NVSwitch on a B200 box), then builds rows for its owned particles only:

    q_all      = [ owned | ghost slots from below | ghost slots from above ]      (fixed capacity per face)
    global_ids = [ own ids | received ids ]
    nlb200_build_subset(handle, q_all, n_total, n_owned, global_ids)      (include/nlist_b200.h)

The ghost buffers have a FIXED capacity per face and unused slots hold NaN records, which the build treats as absent
(include/nlist_b200.h, nlb200_pack_slab): no rank has to learn a count before it posts its receive or launches its
build, so an exchange + build is enqueued without a single host synchronisation and replays the library's CUDA graph.

Every rank assigns cells on the GLOBAL cell grid (the handle is created with the global box), so the rows it emits are
exactly the rows a single-GPU build of the whole system would emit for those particles — same partners, same order
(a cell's particles are sorted by global id) — but bins, sorts and searches only its own WINDOW of that grid: the
cells of its slab plus the ghost layer (`cell_window()`, nlb200_set_cell_window), not a grid that is mostly empty.  HALF
lists: the row of the smaller global id keeps the pair, so each pair is emitted once, by the rank that owns that
particle (ghosts are needed from both faces).  Open boundary (the reference measures distances without minimum image,
neighlist_cpu.hpp:219-223): the end slabs have one neighbour.

Device path: selection, packing and padding of both faces run in ONE kernel of the library (nlb200_pack_faces).  The same
partition / exchange / ownership logic also accepts CPU tensors (torch ops + the `gloo` backend); that branch exists so
that the world_size-2 tests can exercise the logic without GPUs — it is not a compute fallback: the list build itself
is injected by the caller (`build_fn`) and is the CUDA library in the product.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


class SlabDecomposition:
    def __init__(self, world: int, rank: int, box, search_length: float, axis: int = 2, stride: int = 4,
                 group=None, slack: float = 1.5):
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
        self.slack = float(slack)  # ghost capacity per face = expected count * slack + 1024
        self._cap = None
        self._qall = self._gall = self._sq = self._sg = self._cnt = self._cnt_host = None
        self._last = None
        self._gid_default = None
        self._cnt2 = self._state = None

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

    def cell_window(self):
        """(axis, first_cell, n_cells): the cells of the global grid this rank can hold particles in — its slab plus one
        search length of ghosts on either side, plus one cell of slack for the rounding of int(q * ims) — for
        VerletListB200(..., cell_window=...).  None on one rank."""
        if self.world == 1:
            return None
        m = int(self.box[self.axis] / self.sl)  # neighlist_cpu.hpp:384-387
        ms = self.box[self.axis] / m
        lo = 0 if self.rank == 0 else int((self.lo - self.sl) / ms) - 1
        hi = m - 1 if self.rank + 1 == self.world else int((self.hi + self.sl) / ms) + 1
        lo, hi = max(lo, 0), min(hi, m - 1)
        if hi - lo + 1 == 3 and m > 3:  # a window of exactly 3 cells would be taken for a 3-cell (wrapped) axis
            if hi + 1 < m:
                hi += 1
            else:
                lo -= 1
        return (self.axis, lo, hi - lo + 1)

    def n_faces(self) -> int:
        return (1 if self.rank > 0 else 0) + (1 if self.rank + 1 < self.world else 0)

    def ghost_capacity(self, n_owned: int) -> int:
        """Records moved per face.  Fixed, so that an exchange needs no host synchronisation (nobody has to learn a
        count before posting a receive or launching the build): unused slots travel as NaN records, which
        nlb200_build_subset treats as absent."""
        if self.world == 1:
            return 0
        if self._cap is None:
            # both sides of a face must move the same number of records: agree on the largest estimate once
            # (a collective — every rank makes its first ghost_capacity / max_ghosts / exchange call together)
            frac = min(1.0, self.sl / self.thickness)
            cap = (int(n_owned * frac * self.slack) + 1024 + 31) // 32 * 32
            dev = "cuda" if dist.get_backend(self.group) == "nccl" else "cpu"
            t = torch.tensor([cap], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            self._cap = int(t.item())
        return self._cap

    def max_ghosts(self, n_owned: int) -> int:
        """Ghost slots behind the owned records: initialise the handle for n_owned + max_ghosts(n_owned)."""
        return self.n_faces() * self.ghost_capacity(n_owned)

    # -- exchange -------------------------------------------------------------------------------------------------
    def _pack(self, q: torch.Tensor, gid: torch.Tensor, lo: float, hi: float, out_q: torch.Tensor,
              out_g: torch.Tensor, count: torch.Tensor) -> None:
        """out_q/out_g[0:k] = the records / global ids with lo <= q[:, axis] < hi (ascending), NaN records behind;
        count[0] = k (may exceed the capacity: overflow, reported by check())."""
        cap = out_q.shape[0]
        if not q.is_cuda:  # logic tests on CPU tensors (gloo); same contract as the CUDA kernels
            z = q[:, self.axis]
            idx = torch.nonzero((z >= lo) & (z < hi)).flatten()
            count[0] = idx.numel()
            idx = idx[:cap]
            out_q.fill_(float("nan"))
            out_q[:idx.numel()] = q[idx]
            out_g[:idx.numel()] = gid[idx]
            return
        raise RuntimeError("CUDA tensors are packed by nlb200_pack_faces (exchange), not here")

    def _alloc(self, n_total, cap, peers, dtype, dev):
        # persistent buffers: stable device pointers and sizes, so identical builds replay the library's CUDA graph
        self._qall = torch.empty((n_total, self.stride), dtype=dtype, device=dev)
        self._gall = torch.zeros(n_total, dtype=torch.int32, device=dev)
        self._sq = {p: torch.empty((cap, self.stride), dtype=dtype, device=dev) for p in peers}
        self._sg = {p: torch.zeros(cap, dtype=torch.int32, device=dev) for p in peers}
        self._cnt = {p: torch.zeros(1, dtype=torch.int64, device=dev) for p in peers}
        self._cnt_host = {p: (torch.zeros(1, dtype=torch.int64).pin_memory() if dev.type == "cuda"
                              else torch.zeros(1, dtype=torch.int64)) for p in peers}

    def exchange(self, q_owned: torch.Tensor, gid_owned: torch.Tensor):
        """Returns (q_all, gid_all, n_owned): the owned records followed by `ghost_capacity` slots per neighbour face
        (below first), absent slots holding NaN.  One grouped send/recv, no host synchronisation."""
        n = q_owned.shape[0]
        if self.world == 1:
            return q_owned, gid_owned, n
        dev = q_owned.device
        cap = self.ghost_capacity(n)
        peers = [p for p in (self.rank - 1, self.rank + 1) if 0 <= p < self.world]
        n_total = n + cap * len(peers)
        if (self._qall is None or self._qall.shape[0] != n_total or self._qall.device != dev
                or self._qall.dtype != q_owned.dtype):
            self._alloc(n_total, cap, peers, q_owned.dtype, dev)
        if q_owned.data_ptr() != self._qall.data_ptr():  # owned_view(): the caller already writes in place
            self._qall[:n].copy_(q_owned)
        if gid_owned.data_ptr() != self._gall.data_ptr():
            self._gall[:n].copy_(gid_owned)
        lo_p, hi_p = self.rank - 1, self.rank + 1
        if q_owned.is_cuda:
            # both faces in one kernel launch (nlb200_pack_faces)
            L = _lib.lib()
            if self._state is None or self._state.device != dev:
                self._state = torch.zeros(4, dtype=torch.int64, device=dev)  # zeroed once: the kernel leaves it zeroed
                self._cnt2 = torch.zeros(2, dtype=torch.int64, device=dev)
            dtype = _lib.F64 if q_owned.dtype == torch.float64 else _lib.F32
            has_lo, has_hi = lo_p in self._sq, hi_p in self._sq
            st = L.nlb200_pack_faces(
                q_owned.data_ptr(), gid_owned.data_ptr(), n, dtype, self.stride, self.axis,
                self.lo + self.sl if has_lo else -float("inf"), self.hi - self.sl if has_hi else float("inf"),
                self._sq[lo_p].data_ptr() if has_lo else None, self._sg[lo_p].data_ptr() if has_lo else None,
                self._sq[hi_p].data_ptr() if has_hi else None, self._sg[hi_p].data_ptr() if has_hi else None,
                cap, self._cnt2.data_ptr(), self._state.data_ptr(), torch.cuda.current_stream().cuda_stream)
            if st != _lib.OK:
                raise _lib.NlistError(st, "nlb200_pack_faces failed")
            if has_lo:
                self._cnt[lo_p] = self._cnt2[0:1]
            if has_hi:
                self._cnt[hi_p] = self._cnt2[1:2]
        else:
            for p in peers:
                if p < self.rank:
                    self._pack(q_owned, gid_owned, -float("inf"), self.lo + self.sl, self._sq[p], self._sg[p],
                               self._cnt[p])
                else:
                    self._pack(q_owned, gid_owned, self.hi - self.sl, float("inf"), self._sq[p], self._sg[p],
                               self._cnt[p])
        ops, at = [], n
        for p in peers:
            ops.append(dist.P2POp(dist.isend, self._sq[p], p, group=self.group))
            ops.append(dist.P2POp(dist.irecv, self._qall[at:at + cap], p, group=self.group))
            ops.append(dist.P2POp(dist.isend, self._sg[p], p, group=self.group))
            ops.append(dist.P2POp(dist.irecv, self._gall[at:at + cap], p, group=self.group))
            at += cap
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        self._last = (self._qall, self._gall, n)
        return self._last

    def refresh(self, nl, q_owned: torch.Tensor, stream=None):
        """Incremental halo refresh over the send/recv transport: between two builds, the records
        the last exchange sent — the same particles, in the same order, found again through the global ids that went
        with them (the owned global ids must ascend, as partition() and global_ids() give them) — are re-sent at
        their CURRENT positions into the same ghost slots of the neighbours.  No selection, ids and list untouched.
        Every rank calls it; returns the assembly buffers.  (`nl` is unused here: the peer-store transport keeps
        the recorded face set in the handle.)"""
        if self.world == 1:
            return q_owned, self._gall, q_owned.shape[0]
        if self._last is None:
            raise _lib.NlistError(_lib.ERR_STATE, "refresh() needs an exchange (a build) first")
        n = q_owned.shape[0]
        ctx = torch.cuda.stream(stream) if (stream is not None and q_owned.is_cuda) else _null()
        with ctx:
            if q_owned.data_ptr() != self._qall.data_ptr():
                self._qall[:n].copy_(q_owned)
            gid_owned = self._gall[:n]
            peers = [p for p in (self.rank - 1, self.rank + 1) if 0 <= p < self.world]
            cap = self._sq[peers[0]].shape[0]
            ops, at = [], n
            for p in peers:
                k = int(min(int(self._cnt[p][0]), cap))  # records the last exchange sent to p (host read: rare call)
                if k > 0:
                    idx = torch.searchsorted(gid_owned, self._sg[p][:k])
                    self._sq[p][:k] = q_owned[idx]
                ops.append(dist.P2POp(dist.isend, self._sq[p], p, group=self.group))
                ops.append(dist.P2POp(dist.irecv, self._qall[at:at + cap], p, group=self.group))
                at += cap
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        self._last = (self._qall, self._gall, n)
        return self._last

    def owned_view(self, n_owned: int, dtype=torch.float64, device=None):
        """(positions, global ids) views of the first n_owned slots of the assembly buffers: a caller that keeps its
        particles there saves the device-to-device copy of every exchange (the buffers are allocated here)."""
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        cap = self.ghost_capacity(n_owned)
        peers = [p for p in (self.rank - 1, self.rank + 1) if 0 <= p < self.world]
        n_total = n_owned + cap * len(peers)
        if self._qall is None or self._qall.shape[0] != n_total:
            self._alloc(n_total, cap, peers, dtype, dev)
        return self._qall[:n_owned], self._gall[:n_owned]

    def check(self) -> tuple:
        """After the build's stream has been synchronised: raises if a face had more ghosts than the capacity;
        returns the ghosts sent (below, above)."""
        sent = []
        for p in (self.rank - 1, self.rank + 1):
            c = int(self._cnt[p][0]) if (self._cnt is not None and p in self._cnt) else 0  # small D2H, synchronises
            if c > (self._cap or 0):
                raise _lib.NlistError(_lib.ERR_CAPACITY,
                                      f"{c} ghosts for rank {p} exceed the face capacity {self._cap}: raise `slack`")
            sent.append(c)
        return tuple(sent)

    def last_assembled(self):
        """(q_all, gid_all, n_owned) of the last exchange — lets one rank rebuild (e.g. under a profiler) without a new
        exchange, which would need every rank."""
        return self._last

    # -- build ----------------------------------------------------------------------------------------------------
    def global_ids(self, n_owned: int, device) -> torch.Tensor:
        """Global ids of equally sized slabs (local_fcc_slab): rank * n + local index."""
        if self._gid_default is None or self._gid_default.numel() != n_owned or self._gid_default.device != device:
            self._gid_default = torch.arange(n_owned, dtype=torch.int32, device=device) + self.rank * n_owned
        return self._gid_default

    def build(self, nl, q_owned: torch.Tensor, stream=None, gid_owned: torch.Tensor | None = None, build_fn=None):
        """Ghost exchange followed by the list build of the owned rows.  `nl` is a VerletListB200 created with the
        GLOBAL box and initialised for n_owned + max_ghosts(n_owned) particles; `build_fn(q_all, n_owned, gid_all)`
        replaces nl.build in the CPU logic tests."""
        ctx = torch.cuda.stream(stream) if (stream is not None and q_owned.is_cuda) else _null()
        with ctx:
            if gid_owned is None:
                gid_owned = self.global_ids(q_owned.shape[0], q_owned.device)
            q_all, gid_all, n_owned = self.exchange(q_owned, gid_owned)
            if build_fn is not None:
                return build_fn(q_all, n_owned, gid_all)
            nl.build(q_all, n_owned=n_owned, global_ids=gid_all if self.world > 1 else None, stream=stream)
        return None


class GraphedHaloBuild:
    """Exchange + build of identical steps (same buffers, same counts) replayed as ONE CUDA graph: the halo packing
    kernels, the grouped NCCL send/recv and the build's kernel chain are captured once, so a
    step costs one graph launch on the host instead of ~15 launches and a c10d group call.  For small systems the
    host side is what limits a multi-GPU step.  Falls back to eager steps if the capture is refused."""

    def __init__(self, halo: SlabDecomposition, nl, q_owned: torch.Tensor, gid_owned: torch.Tensor, stream):
        self.halo, self.nl, self.q, self.g, self.stream = halo, nl, q_owned, gid_owned, stream
        self.graph = None
        for _ in range(3):  # allocations, communicators, the library's graph: all before the capture
            halo.build(nl, q_owned, stream, gid_owned=gid_owned)
        stream.synchronize()
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream, capture_error_mode="thread_local"):
                halo.build(nl, q_owned, stream, gid_owned=gid_owned)
            self.graph = g
        except Exception as e:  # noqa: BLE001 — any capture problem means: stay eager
            self.error = repr(e)
            self.graph = None
            torch.cuda.synchronize()

    def step(self) -> None:
        if self.graph is not None:
            with torch.cuda.stream(self.stream):
                self.graph.replay()
            # the library must know that a build is in flight on this stream
            self.nl._mark_pending(self.stream)
        else:
            self.halo.build(self.nl, self.q, self.stream, gid_owned=self.g)


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _DevArray:
    """__cuda_array_interface__ view of raw device memory (library-owned or a neighbour's, opened by CUDA IPC)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


class PeerSlabDecomposition(SlabDecomposition):
    """The slab decomposition with the halo exchange done by PEER STORES instead of NCCL send/recv: the packing kernel
    of a rank writes its face particles straight into its neighbours' assembly buffers over NVLink (pointers from
    CUDA IPC, exchanged once), flags in a 64-byte control block per rank order the steps on the device
    (include/nlist_b200.h, "halo exchange by peer stores").  Per step: nlb200_pack_faces_p2p, nlb200_halo_wait, the
    build, nlb200_halo_done — four launches, no host synchronisation, no NCCL kernel.  Measured at 2 GPUs on the
    contract workload: exchange 49 us (packing + NCCL group of 4 send/recv pairs) -> see profiles/r02_halo_p2p.md.

    Every rank allocates [control | positions (n_cap + 2 cap records) | global ids] of the same size; a neighbour's ghost
    region sits at the same offset on every rank.  torch.distributed is used once, on the host, to agree on the sizes
    and to pass the IPC handles around.  If a rank cannot export or open a handle, all ranks fall back to NCCL."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self._p2p = None
        self._synced_handle = None

    # -- set-up (collective) ----------------------------------------------------------------------------------------
    def _setup(self, n_owned: int, dtype, dev) -> bool:
        import ctypes as C
        if self._p2p is not None:
            return self._p2p.get("ok", False)
        L = _lib.lib()
        cap = self.ghost_capacity(n_owned)
        t = torch.tensor([n_owned], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        n_cap = (int(t.item()) + 31) // 32 * 32
        esz = 8 if dtype == torch.float64 else 4
        n_total = n_cap + 2 * cap
        o_q = 256
        o_g = (o_q + n_total * self.stride * esz + 255) // 256 * 256
        nbytes = o_g + n_total * 4
        base = C.c_void_p()
        handle = (C.c_char * 64)()
        st = L.nlb200_p2p_alloc(nbytes, C.byref(base), C.cast(handle, C.c_void_p))
        mine = {"ok": st == _lib.OK, "handle": bytes(handle), "nbytes": nbytes}
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=self.group)
        ok = all(e["ok"] and e["nbytes"] == nbytes for e in everyone)
        peers = {}
        if ok:
            for p in (self.rank - 1, self.rank + 1):
                if 0 <= p < self.world:
                    ptr = C.c_void_p()
                    hb = (C.c_char * 64).from_buffer_copy(everyone[p]["handle"])
                    if L.nlb200_p2p_open(C.cast(hb, C.c_void_p), C.byref(ptr)) != _lib.OK:
                        ok = False
                    else:
                        peers[p] = ptr.value
        flag = torch.tensor([1 if ok else 0], dtype=torch.int64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        ok = bool(int(flag.item()))
        self._p2p = {"ok": ok, "base": base.value, "peers": peers, "cap": cap, "n_cap": n_cap, "n_total": n_total,
                     "o_q": o_q, "o_g": o_g, "esz": esz, "nbytes": nbytes}
        if not ok:
            return False
        ts = "<f8" if dtype == torch.float64 else "<f4"
        self._qall = torch.as_tensor(_DevArray(base.value + o_q, (n_total, self.stride), ts), device=dev)
        self._gall = torch.as_tensor(_DevArray(base.value + o_g, (n_total,), "<i4"), device=dev)
        self._qall.fill_(float("nan"))  # every slot starts as an absent record (also the regions of missing faces)
        self._gall.zero_()
        self._ctrl = torch.as_tensor(_DevArray(base.value, (8,), "<i8"), device=dev)
        self._state = torch.zeros(8, dtype=torch.int64, device=dev)  # cursors, ticket, previous counts
        self._cnt2 = torch.zeros(2, dtype=torch.int64, device=dev)
        self._send_idx = torch.zeros((2, max(cap, 1)), dtype=torch.int32, device=dev)  # the recorded face set (refresh)
        torch.cuda.synchronize()
        dist.barrier(group=self.group)  # nobody writes into a neighbour before it has initialised its buffer
        return True

    def max_ghosts(self, n_owned: int) -> int:
        """Slots behind the owned records in the peer layout: both ghost regions (an end slab's missing face stays
        absent) plus the padding of the owned region up to the common size."""
        if self.world == 1:
            return 0
        if self._p2p is not None and self._p2p.get("ok"):
            return self._p2p["n_total"] - n_owned
        return 2 * self.ghost_capacity(n_owned) + 64

    def owned_view(self, n_owned: int, dtype=torch.float64, device=None):
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        if not self._setup(n_owned, dtype, dev):
            return super().owned_view(n_owned, dtype, dev)
        return self._qall[:n_owned], self._gall[:n_owned]

    def _addr(self, base: int, what: str, slot: int = 0) -> int:
        p = self._p2p
        if what == "q":
            return base + p["o_q"] + slot * self.stride * p["esz"]
        if what == "g":
            return base + p["o_g"] + slot * 4
        return base + {"ready0": 8, "ready1": 16, "free0": 24, "free1": 32}[what]

    def _pack_args(self):
        """(cut_lo, cut_hi, peer regions, capacity, counts, state, ready flags) — what nlb200_pack_faces_p2p and
        nlb200_set_halo_pack are told: my lower-face particles go into the lower neighbour's "ghosts from above"
        region (second region), my upper-face particles into the upper neighbour's "ghosts from below" region."""
        p = self._p2p
        lo_p, hi_p = p["peers"].get(self.rank - 1), p["peers"].get(self.rank + 1)
        cap, n_cap = p["cap"], p["n_cap"]
        return (self.lo + self.sl if lo_p else -float("inf"), self.hi - self.sl if hi_p else float("inf"),
                self._addr(lo_p, "q", n_cap + cap) if lo_p else None, self._addr(lo_p, "g", n_cap + cap) if lo_p else None,
                self._addr(hi_p, "q", n_cap) if hi_p else None, self._addr(hi_p, "g", n_cap) if hi_p else None,
                cap, self._cnt2.data_ptr(), self._state.data_ptr(),
                self._addr(lo_p, "ready1") if lo_p else None, self._addr(hi_p, "ready0") if hi_p else None)

    def exchange(self, q_owned: torch.Tensor, gid_owned: torch.Tensor, launch: bool = True):
        """launch=False: only place the owned records and return the assembly buffers — the build of a handle with
        nlb200_set_halo_pack packs, sends and waits itself (build())."""
        if self.world == 1 or not q_owned.is_cuda:
            return super().exchange(q_owned, gid_owned)
        n = q_owned.shape[0]
        if not self._setup(n, q_owned.dtype, q_owned.device):
            return super().exchange(q_owned, gid_owned)
        p = self._p2p
        if q_owned.data_ptr() != self._qall.data_ptr():
            self._qall[:n].copy_(q_owned)
        if gid_owned.data_ptr() != self._gall.data_ptr():
            self._gall[:n].copy_(gid_owned)
        L = _lib.lib()
        lo_p, hi_p = p["peers"].get(self.rank - 1), p["peers"].get(self.rank + 1)
        dtype = _lib.F64 if q_owned.dtype == torch.float64 else _lib.F32
        s = torch.cuda.current_stream().cuda_stream
        if launch:
            a = self._pack_args()
            st = L.nlb200_pack_faces_p2p(self._qall.data_ptr(), self._gall.data_ptr(), n, dtype, self.stride, self.axis,
                                         a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], p["base"], a[9], a[10], s)
            if st != _lib.OK:
                raise _lib.NlistError(st, "nlb200_pack_faces_p2p failed")
        # (the packing kernel ends by waiting for this rank's own ghosts: no separate nlb200_halo_wait launch)
        self._cnt = {}
        if lo_p:
            self._cnt[self.rank - 1] = self._cnt2[0:1]
        if hi_p:
            self._cnt[self.rank + 1] = self._cnt2[1:2]
        self._last = (self._qall, self._gall, n)
        return self._last

    def done(self) -> None:
        """After the build has been enqueued: the neighbours may overwrite this rank's ghosts (stream-ordered)."""
        p = self._p2p
        if not p or not p.get("ok"):
            return
        lo_p, hi_p = p["peers"].get(self.rank - 1), p["peers"].get(self.rank + 1)
        st = _lib.lib().nlb200_halo_done(p["base"], self._addr(lo_p, "free1") if lo_p else None,
                                         self._addr(hi_p, "free0") if hi_p else None,
                                         torch.cuda.current_stream().cuda_stream)
        if st != _lib.OK:
            raise _lib.NlistError(st, "nlb200_halo_done failed")

    def build(self, nl, q_owned: torch.Tensor, stream=None, gid_owned: torch.Tensor | None = None, build_fn=None):
        ctx = torch.cuda.stream(stream) if (stream is not None and q_owned.is_cuda) else _null()
        with ctx:
            if gid_owned is None:
                gid_owned = self.global_ids(q_owned.shape[0], q_owned.device)
            # the library's own build of a peer-store decomposition packs, sends and waits inside its binning kernels
            # (nlb200_set_halo_pack; NLB_HALO_FUSED=0: the separate packing kernel)
            fused = (build_fn is None and self.world > 1 and q_owned.is_cuda
                     and os.environ.get("NLB_HALO_FUSED", "1") != "0"
                     and self._setup(q_owned.shape[0], q_owned.dtype, q_owned.device))
            if fused:
                q_all, gid_all, n_owned = self.exchange(q_owned, gid_owned, launch=False)
            else:
                q_all, gid_all, n_owned = self.exchange(q_owned, gid_owned)
            if build_fn is not None:
                out = build_fn(q_all, n_owned, gid_all)
            else:
                if self.uses_peer_stores() and self._synced_handle is not nl:
                    # nlb200_halo_done becomes part of the build's last kernel
                    p = self._p2p
                    lo_p, hi_p = p["peers"].get(self.rank - 1), p["peers"].get(self.rank + 1)
                    _lib.check(nl._h, _lib.lib().nlb200_set_halo_sync(
                        nl._h, p["base"], self._addr(lo_p, "free1") if lo_p else None,
                        self._addr(hi_p, "free0") if hi_p else None))
                    if fused:
                        a = self._pack_args()
                        _lib.check(nl._h, _lib.lib().nlb200_set_halo_pack(
                            nl._h, self.axis, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10],
                            self._send_idx[0].data_ptr(), self._send_idx[1].data_ptr()))
                    self._synced_handle = nl
                nl.build(q_all, n_owned=n_owned, global_ids=gid_all if self.world > 1 else None, stream=stream)
                out = None
            if q_owned.is_cuda and self.world > 1 and (build_fn is not None or not self.uses_peer_stores()):
                self.done()
        return out

    def refresh(self, nl, q_owned: torch.Tensor, stream=None):
        """Incremental halo refresh: between two builds, re-send the CURRENT positions of the face set
        the last build of `nl` recorded into the same ghost slots of the neighbours and wait for this rank's own ghosts
        (nlb200_halo_refresh).  Every rank calls it; returns the assembly buffers (owned records + refreshed ghosts).
        After consuming the ghosts call done() — the neighbours may then overwrite them."""
        if not self.uses_peer_stores():
            return super().refresh(nl, q_owned, stream)  # the send/recv fallback
        if self._synced_handle is not nl:
            raise _lib.NlistError(_lib.ERR_STATE, "refresh() needs a build of this handle with the folded exchange first")
        n = q_owned.shape[0]
        ctx = torch.cuda.stream(stream) if stream is not None else _null()
        with ctx:
            if q_owned.data_ptr() != self._qall.data_ptr():
                self._qall[:n].copy_(q_owned)
            _lib.check(nl._h, _lib.lib().nlb200_halo_refresh(nl._h, self._qall.data_ptr(),
                                                             torch.cuda.current_stream().cuda_stream))
        return self._qall, self._gall, n

    def uses_peer_stores(self) -> bool:
        return bool(self._p2p and self._p2p.get("ok"))

    def check(self) -> tuple:
        if self.uses_peer_stores() and int(self._ctrl[5]) != 0:
            raise _lib.NlistError(_lib.ERR_STATE, "a halo flag did not arrive within the spin limit (lost neighbour?)")
        return super().check()

    def close(self) -> None:
        p = self._p2p
        if p and p.get("ok"):
            torch.cuda.synchronize()
            dist.barrier(group=self.group)  # nobody frees memory a neighbour may still write into
            L = _lib.lib()
            if self._synced_handle is not None and self._synced_handle._h:
                L.nlb200_set_halo_sync(self._synced_handle._h, None, None, None)
            self._synced_handle = None
            self._qall = self._gall = self._ctrl = None
            for ptr in p["peers"].values():
                L.nlb200_p2p_close(ptr)
            L.nlb200_p2p_free(p["base"])
        self._p2p = None
