"""Periodic boundaries (minimum image) on top of the open-boundary build — SURVEY.md §8f f3.

The reference wraps only CELL INDICES (neighlist_cpu.hpp:61-66); its distances are plain Euclidean
(neighlist_cpu.hpp:219-223), so its list is an open-boundary list and that is what libnlist_b200 reproduces bit for
bit.  Real MD callers need the minimum image.  It is provided here the way domain-decomposed MD codes do it, with the
machinery of the multi-GPU halo: every particle within the search length of a face gets a periodic IMAGE beyond the
opposite face (axis by axis, so edge and corner images follow from images of images), the images travel as ghost
records behind the owned particles with the ORIGINAL particle's id as their global id, and one open-boundary
nlb200_build_subset over the extended box yields rows whose partners are the minimum-image neighbours:

    q_all = [ particles | x images | y images (of particles and x images) | z images (of all of those) ] + SL
    box   = L + 2 SL per axis,   n_owned = n,   global_ids = [ 0..n-1 | id of the imaged particle ... ]

Image buffers have a fixed capacity (unused slots are NaN records = absent, include/nlist_b200.h), so a build needs no
host synchronisation and replays the library's CUDA graph.  Preconditions: positions in [0, L) on every axis and
L >= 2 * search_length + (so that a particle and its image are never both partners, and no particle meets its own
image).  HALF lists: the pair (i, j) is kept in the row of the smaller id, whichever of the two is the image.
"""
from __future__ import annotations

import torch

from . import _lib
from .neighlist import VerletListB200


class PeriodicVerletList:
    def __init__(self, search_length: float, Lx: float, Ly: float, Lz: float, dtype="f64", mode="full_csr",
                 slack: float = 1.5, **opts):
        self.sl = float(search_length)
        self.L = (float(Lx), float(Ly), float(Lz))
        if min(self.L) < 2.0 * self.sl * (1.0 + 1e-9):
            raise ValueError("periodic box must be at least twice the search length on every axis")
        self.slack = float(slack)
        self.nl = VerletListB200(search_length, *(l + 2.0 * self.sl for l in self.L), dtype=dtype, mode=mode,
                                 position_stride=4, **opts)
        self.n = 0
        self._caps = None
        self._buf = None

    def initialize(self, n: int, max_entries: int = 0) -> None:
        self.n = int(n)
        caps, cur = [], self.n
        for a in range(3):
            cap = (int(cur * self.sl / self.L[a] * self.slack) + 256 + 31) // 32 * 32
            caps.append(cap)
            cur += 2 * cap
        self._caps, self.n_total = caps, cur
        if max_entries == 0:
            dens = self.n / (self.L[0] * self.L[1] * self.L[2])
            per = dens * 4.18879 * self.sl ** 3 * (0.5 if self.nl.mode == _lib.HALF_CSR else 1.0)
            max_entries = int(self.n * per * 1.3) + 16 * self.n + 1024
        self.nl.initialize(self.n_total, max_entries)

    def _alloc(self, dtype, dev):
        L = _lib.lib()
        self._q = torch.empty((self.n_total, 4), dtype=dtype, device=dev)
        self._g = torch.zeros(self.n_total, dtype=torch.int32, device=dev)
        self._g[:self.n] = torch.arange(self.n, dtype=torch.int32, device=dev)
        self._cnt = torch.zeros(6, dtype=torch.int64, device=dev)
        self._ws = torch.empty(2 * L.nlb200_select_slab_workspace(self.n_total) + 512, dtype=torch.uint8, device=dev)
        self._buf = True

    def build(self, q: torch.Tensor, stream: torch.cuda.Stream | None = None) -> None:
        """q: (n, 4) CUDA positions in [0, L).  Asynchronous; no host synchronisation."""
        if q.shape[0] != self.n or q.shape[1] != 4 or not q.is_cuda:
            raise ValueError("q must be a CUDA tensor of shape (n, 4) with the n given to initialize")
        s = stream if stream is not None else torch.cuda.current_stream()
        with torch.cuda.stream(s):
            if self._buf is None or self._q.dtype != q.dtype or self._q.device != q.device:
                self._alloc(q.dtype, q.device)
            L = _lib.lib()
            dtype = _lib.F64 if q.dtype == torch.float64 else _lib.F32
            self._q[:self.n].copy_(q)
            cur = self.n
            for a in range(3):
                cap = self._caps[a]
                lo = self._q[cur:cur + cap]            # images of the particles near the lower face: + L
                hi = self._q[cur + cap:cur + 2 * cap]  # near the upper face: - L
                st = L.nlb200_pack_slab2(self._q.data_ptr(), self._g.data_ptr(), cur, dtype, 4, a, self.sl,
                                         self.L[a] - self.sl, lo.data_ptr(), self._g[cur:].data_ptr(), hi.data_ptr(),
                                         self._g[cur + cap:].data_ptr(), cap, self._cnt[2 * a:].data_ptr(),
                                         self._ws.data_ptr(), self._ws.numel(), s.cuda_stream)
                if st != _lib.OK:
                    raise _lib.NlistError(st, "nlb200_pack_slab2 failed")
                lo[:, a] += self.L[a]
                hi[:, a] -= self.L[a]
                cur += 2 * cap
            self._q[:, :3] += self.sl  # origin of the extended box
            self.nl.build(self._q, n_owned=self.n, global_ids=self._g, stream=s)

    def synchronize(self):
        st = self.nl.synchronize()
        cnt = self._cnt.cpu().tolist()
        for a in range(3):
            if max(cnt[2 * a], cnt[2 * a + 1]) > self._caps[a]:
                raise _lib.NlistError(_lib.ERR_CAPACITY, f"axis {a}: {max(cnt[2*a], cnt[2*a+1])} periodic images exceed "
                                                         f"the capacity {self._caps[a]}: raise `slack`")
        return st

    def number_of_partners(self):
        return self.nl.number_of_partners()

    def offsets(self):
        return self.nl.offsets()

    def partners(self):
        return self.nl.partners()

    def number_of_pairs(self):
        return self.nl.number_of_pairs()

    def close(self):
        self.nl.close()


class PeriodicSlabDecomposition:
    """Minimum-image lists over the GPUs of one box: the halo of the slab decomposition (parallel.py) closed into a
    RING — SURVEY.md §8f f3, "periodic boundaries turn the multi-GPU halo into a ring".

    One process per GPU, `world` slabs of equal thickness along `axis`; rank r owns the particles with
    r * t <= q[axis] < (r + 1) * t.  Per build:

      1. ring exchange along `axis` (ONE grouped send/recv): the owned particles within the search length of the lower
         face go to rank r - 1, those near the upper face to rank r + 1, indices modulo `world`; what arrives across the
         periodic seam is shifted by -/+ L[axis], so every rank sees a contiguous slab with a ghost layer on both sides;
      2. periodic images along the two other axes, locally, axis by axis over everything assembled so far (owned
         particles, ring ghosts, earlier images), exactly as PeriodicVerletList does on one GPU;
      3. one open-boundary nlb200_build_subset over the box extended by the search length on every side:

         q_all = [ owned | from below | from above | images axis a (lo, hi) | images axis b (lo, hi) ] + SL
         n_owned = owned,  global_ids = the id of the particle a record is (an image of)

    Every buffer has a fixed capacity (absent slots are NaN records), so exchange + build need no host synchronisation.
    All ranks bin on the same extended GLOBAL grid, hence a rank's rows equal the rows PeriodicVerletList gives those
    particles on one GPU.  Preconditions: positions in [0, L) on every axis, L >= 2 * search_length on every axis,
    slab thickness >= search_length, world >= 2 (one GPU: PeriodicVerletList).  CPU tensors + the gloo backend are
    accepted for the logic tests; the list build itself is the CUDA library (`build_fn` injects the oracle in tests).
    """

    def __init__(self, world: int, rank: int, box, search_length: float, axis: int = 2, group=None,
                 slack: float = 1.5):
        import torch.distributed as dist
        self._dist = dist
        if world < 2 or not (0 <= rank < world):
            raise ValueError("PeriodicSlabDecomposition needs world >= 2 and 0 <= rank < world")
        self.world, self.rank, self.axis = int(world), int(rank), int(axis)
        self.L = tuple(float(b) for b in box)
        self.sl = float(search_length)
        if min(self.L) < 2.0 * self.sl * (1.0 + 1e-9):
            raise ValueError("periodic box must be at least twice the search length on every axis")
        self.thickness = self.L[self.axis] / world
        if self.thickness < self.sl:
            raise ValueError("slab thinner than the search length: ghosts would come from second neighbours")
        self.lo, self.hi = rank * self.thickness, (rank + 1) * self.thickness
        self.image_axes = tuple(a for a in range(3) if a != self.axis)
        self.group, self.slack = group, float(slack)
        self._cap_z = None
        self._bufs = None

    # -- geometry / capacities ----------------------------------------------------------------------------------
    def extended_box(self):
        """Box of the handle: every axis grown by the search length on both sides."""
        return tuple(l + 2.0 * self.sl for l in self.L)

    def owns(self, q):
        z = q[:, self.axis]
        hi = self.hi if self.rank + 1 < self.world else float("inf")  # guard against L * (1 - eps) rounding
        return (z >= self.lo) & (z < hi)

    def partition(self, q_global):
        import numpy as np
        m = self.owns(q_global)
        return np.ascontiguousarray(q_global[m]), np.nonzero(m)[0].astype(np.int32)

    def ring_capacity(self, n_owned: int) -> int:
        """Records per ring message; both ends of a message must agree, so the largest estimate is agreed once
        (a collective: every rank makes its first call together)."""
        if self._cap_z is None:
            cap = (int(n_owned * min(1.0, self.sl / self.thickness) * self.slack) + 1024 + 31) // 32 * 32
            dev = "cuda" if self._dist.get_backend(self.group) == "nccl" else "cpu"
            t = torch.tensor([cap], dtype=torch.int64, device=dev)
            self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX, group=self.group)
            self._cap_z = int(t.item())
        return self._cap_z

    def layout(self, n_owned: int):
        """(n_total, [(axis, capacity), ...] for the two image axes)."""
        cur = n_owned + 2 * self.ring_capacity(n_owned)
        caps = []
        for a in self.image_axes:
            cap = (int(cur * self.sl / self.L[a] * self.slack) + 256 + 31) // 32 * 32
            caps.append((a, cap))
            cur += 2 * cap
        return cur, caps

    def n_total(self, n_owned: int) -> int:
        """Slots of the assembled array: initialise the handle for this many particles."""
        return self.layout(n_owned)[0]

    # -- selection of the records near the two faces of an axis -------------------------------------------------
    def _faces(self, q, g, n, axis, cut_lo, cut_hi, lo_q, lo_g, hi_q, hi_g, cnt):
        """lo_q/lo_g <- the records of q[:n] with q[axis] < cut_lo, hi_q/hi_g <- those with q[axis] >= cut_hi (ascending,
        NaN records behind); cnt[0:2] = how many there were (may exceed the capacity: reported by check())."""
        cap = lo_q.shape[0]
        if not q.is_cuda:  # logic tests on CPU tensors (gloo): same contract as nlb200_pack_slab2
            z = q[:n, axis]
            for k, (idx, oq, og) in enumerate(((torch.nonzero(z < cut_lo).flatten(), lo_q, lo_g),
                                               (torch.nonzero(z >= cut_hi).flatten(), hi_q, hi_g))):
                cnt[k] = idx.numel()
                idx = idx[:cap]
                oq.fill_(float("nan"))
                oq[:idx.numel()] = q[idx]
                og[:idx.numel()] = g[idx]
            return
        lib = _lib.lib()
        dtype = _lib.F64 if q.dtype == torch.float64 else _lib.F32
        st = lib.nlb200_pack_slab2(q.data_ptr(), g.data_ptr(), n, dtype, 4, axis, cut_lo, cut_hi, lo_q.data_ptr(),
                                   lo_g.data_ptr(), hi_q.data_ptr(), hi_g.data_ptr(), cap, cnt.data_ptr(),
                                   self._ws.data_ptr(), self._ws.numel(), torch.cuda.current_stream().cuda_stream)
        if st != _lib.OK:
            raise _lib.NlistError(st, "nlb200_pack_slab2 failed")

    def _alloc(self, n_owned, dtype, dev):
        n_total, caps = self.layout(n_owned)
        cz = self._cap_z
        b = {"q": torch.empty((n_total, 4), dtype=dtype, device=dev),
             "g": torch.zeros(n_total, dtype=torch.int32, device=dev),
             "sq": [torch.empty((cz, 4), dtype=dtype, device=dev) for _ in range(2)],   # to rank-1, to rank+1
             "sg": [torch.zeros(cz, dtype=torch.int32, device=dev) for _ in range(2)],
             "cnt": torch.zeros(2 + 2 * len(caps), dtype=torch.int64, device=dev),
             "n_owned": n_owned, "n_total": n_total, "caps": caps}
        if dev.type == "cuda":
            self._ws = torch.empty(2 * _lib.lib().nlb200_select_slab_workspace(n_total) + 512, dtype=torch.uint8,
                                   device=dev)
        self._bufs = b

    # -- exchange + images ----------------------------------------------------------------------------------------
    def exchange(self, q_owned: torch.Tensor, gid_owned: torch.Tensor):
        """(q_all, gid_all, n_owned) in the coordinates of the extended box; no host synchronisation."""
        dist = self._dist
        n = q_owned.shape[0]
        if q_owned.shape[1] != 4:
            raise ValueError("positions must be {x, y, z, w} records")
        cz = self.ring_capacity(n)
        b = self._bufs
        if b is None or b["n_owned"] != n or b["q"].dtype != q_owned.dtype or b["q"].device != q_owned.device:
            self._alloc(n, q_owned.dtype, q_owned.device)
            b = self._bufs
        q, g, cnt = b["q"], b["g"], b["cnt"]
        q[:n].copy_(q_owned)
        g[:n].copy_(gid_owned)
        # 1. the ring: lower-face records go down, upper-face records go up
        self._faces(q, g, n, self.axis, self.lo + self.sl, self.hi - self.sl, b["sq"][0], b["sg"][0], b["sq"][1],
                    b["sg"][1], cnt[0:2])
        down, up = (self.rank - 1) % self.world, (self.rank + 1) % self.world
        from_down, from_up = slice(n, n + cz), slice(n + cz, n + 2 * cz)
        # posting order matters when both neighbours are the same rank (world == 2): what I send DOWN is what the peer
        # receives FROM ABOVE, so sends are posted (down, up) and receives (from above, from below)
        ops = [dist.P2POp(dist.isend, b["sq"][0], down, group=self.group),
               dist.P2POp(dist.isend, b["sg"][0], down, group=self.group),
               dist.P2POp(dist.isend, b["sq"][1], up, group=self.group),
               dist.P2POp(dist.isend, b["sg"][1], up, group=self.group),
               dist.P2POp(dist.irecv, q[from_up], up, group=self.group),
               dist.P2POp(dist.irecv, g[from_up], up, group=self.group),
               dist.P2POp(dist.irecv, q[from_down], down, group=self.group),
               dist.P2POp(dist.irecv, g[from_down], down, group=self.group)]
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        # across the periodic seam the neighbour's coordinates are one period away
        if self.rank == 0:
            q[from_down, self.axis] -= self.L[self.axis]
        if self.rank == self.world - 1:
            q[from_up, self.axis] += self.L[self.axis]
        # 2. periodic images along the other two axes, of everything assembled so far
        cur = n + 2 * cz
        for k, (a, cap) in enumerate(b["caps"]):
            lo, hi = slice(cur, cur + cap), slice(cur + cap, cur + 2 * cap)
            self._faces(q, g, cur, a, self.sl, self.L[a] - self.sl, q[lo], g[lo], q[hi], g[hi],
                        cnt[2 + 2 * k:4 + 2 * k])
            q[lo, a] += self.L[a]   # near the lower face: image beyond the upper face
            q[hi, a] -= self.L[a]
            cur += 2 * cap
        # 3. origin of the extended box (the slab axis too: ring ghosts of the end ranks lie outside [0, L))
        q[:, :3] += self.sl
        return q, g, n

    def build(self, nl, q_owned: torch.Tensor, stream=None, gid_owned: torch.Tensor | None = None, build_fn=None):
        """Ring exchange, images and the list build of the owned rows.  `nl`: a VerletListB200 created with
        extended_box() and initialised for n_total(n_owned) particles."""
        ctx = torch.cuda.stream(stream) if (stream is not None and q_owned.is_cuda) else _Null()
        with ctx:
            q_all, gid_all, n_owned = self.exchange(q_owned, gid_owned)
            if build_fn is not None:
                return build_fn(q_all, n_owned, gid_all)
            nl.build(q_all, n_owned=n_owned, global_ids=gid_all, stream=stream)
        return None

    def check(self) -> tuple:
        """After the build's stream has been synchronised: raises if a message or an image buffer overflowed; returns
        the counts (down, up, then lo / hi per image axis)."""
        b = self._bufs
        cnt = b["cnt"].cpu().tolist()
        caps = [self._cap_z, self._cap_z] + [c for _, c in b["caps"] for _ in range(2)]
        for c, cap in zip(cnt, caps):
            if c > cap:
                raise _lib.NlistError(_lib.ERR_CAPACITY, f"{c} records exceed a halo capacity of {cap}: raise `slack`")
        return tuple(cnt)


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
