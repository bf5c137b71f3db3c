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
