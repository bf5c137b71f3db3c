"""Multi-process logic of the slab decomposition (md_neighbor_list_b200/parallel.py) on CPU: world_size 2 and 3 over
the gloo backend.  Partition, ghost selection, exchange and the ownership rule are the product code; the list build
itself is replaced by the oracle (tests may call it), so that no GPU is needed.  The union of the ranks' rows must be
exactly the single-process list of the global system, FULL and HALF."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

SL = 3.3


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _global_system(world: int):
    from oracle import oracle as O
    L = 16.0
    s = (0.25 * 1.0) ** (-1.0 / 3.0)
    sx = int(L / s)
    q = O.gen_fcc(1.0, L, sx, sx, sx * world)
    return q, (L, L, L * world)


def _rows(csr, rows_of, mapping):
    out = []
    for i in rows_of:
        b, e = csr.offsets[i], csr.offsets[i + 1]
        out.append(np.sort(mapping[csr.partners[b:e]]))
    return out


def _worker(rank: int, world: int, port: int):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from md_neighbor_list_b200.parallel import SlabDecomposition
        from oracle import oracle as O
        q, box = _global_system(world)
        ref_full = O.build_full(q, SL, box)
        dec = SlabDecomposition(world, rank, box, SL, axis=2)
        q_own, gid_own = dec.partition(q)
        assert q_own.shape[0] > 0

        def build_fn(q_all, n_owned, gid_all):
            qa = q_all.numpy()
            ga = gid_all.numpy()
            # fixed-capacity ghost slots: NaN records are absent (the library bins them nowhere); drop them here
            assert qa.shape[0] == n_owned + dec.max_ghosts(n_owned)
            present = ~np.isnan(qa[:, 0])
            assert present[:n_owned].all()
            qa, ga = qa[present], ga[present]
            # ghosts must be exactly the foreign particles within SL of this slab's faces
            lo, hi = rank * dec.thickness, (rank + 1) * dec.thickness
            z = q[:, 2]
            want = set()
            if rank > 0:
                want |= set(np.nonzero((z < lo) & (z >= lo - SL))[0].tolist())
            if rank + 1 < world:
                want |= set(np.nonzero((z >= hi) & (z < hi + SL))[0].tolist())
            assert set(ga[n_owned:].tolist()) == want
            assert np.array_equal(qa[n_owned:], q[ga[n_owned:]])
            local = O.build_full(qa, SL, box)
            return _rows(local, range(n_owned), ga), ga[:n_owned]

        rows, gids = dec.build(None, torch.from_numpy(q_own), gid_owned=torch.from_numpy(gid_own), build_fn=build_fn)
        ident = np.arange(q.shape[0])
        want_rows = _rows(ref_full, gids, ident)
        assert len(rows) == len(want_rows)
        for a, b in zip(rows, want_rows):
            assert np.array_equal(a, b)
        # HALF ownership rule: the row of the smaller global id keeps the pair -> every pair exactly once
        half_local = sum(int((r > g).sum()) for r, g in zip(rows, gids))
        tot = torch.tensor([half_local], dtype=torch.int64)
        dist.all_reduce(tot)
        assert int(tot) == O.build_half(q, SL, box).number_of_pairs
        sent = dec.check()  # no face overflowed its capacity
        assert len(sent) == 2 and sum(sent) > 0 and all(c <= dec.ghost_capacity(q_own.shape[0]) for c in sent)
        # every particle is owned by exactly one rank
        cnt = torch.tensor([q_own.shape[0]], dtype=torch.int64)
        dist.all_reduce(cnt)
        assert int(cnt) == q.shape[0]
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_decomposition_matches_single_process(world):
    from oracle import oracle as O
    O.lib()
    mp.spawn(_worker, args=(world, _free_port()), nprocs=world, join=True)


def _refresh_worker(rank: int, world: int, port: int):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from md_neighbor_list_b200.parallel import SlabDecomposition
        q, box = _global_system(world)
        dec = SlabDecomposition(world, rank, box, SL, axis=2)
        q_own, gid_own = dec.partition(q)
        qa0, ga0, n0 = dec.exchange(torch.from_numpy(q_own), torch.from_numpy(gid_own))
        present0 = ~torch.isnan(qa0[:, 0])
        ga0 = ga0.clone()
        # the particles move a little in x and y (the list would stay valid): the same face set, new positions
        rng = np.random.default_rng(1)
        q_new = q.copy()
        q_new[:, :2] += (rng.random((q.shape[0], 2)) - 0.5) * 0.1
        qa1, ga1, n1 = dec.refresh(None, torch.from_numpy(np.ascontiguousarray(q_new[gid_own])))
        assert n1 == n0 and torch.equal(ga1, ga0), "ids and slots must not change"
        present1 = ~torch.isnan(qa1[:, 0])
        assert torch.equal(present1, present0)
        ghosts = torch.nonzero(present1[n0:]).flatten() + n0
        assert ghosts.numel() > 0
        want = torch.from_numpy(q_new)[ga1[ghosts].long()]
        assert torch.equal(qa1[ghosts], want), "a ghost slot must hold its particle's current position"
        assert torch.equal(qa1[:n0], torch.from_numpy(np.ascontiguousarray(q_new[gid_own])))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_refresh_resends_the_recorded_face_set(world):
    """SURVEY.md §8f f2 on the send/recv transport (gloo here, NCCL on the device): between two builds the ghosts are
    refreshed in place — same slots, same ids, current positions."""
    mp.spawn(_refresh_worker, args=(world, _free_port()), nprocs=world, join=True)


def test_slab_thinner_than_search_length_is_rejected():
    from md_neighbor_list_b200.parallel import SlabDecomposition
    with pytest.raises(ValueError):
        SlabDecomposition(4, 0, (50.0, 50.0, 10.0), SL)


def _overflow_worker(rank: int, world: int, port: int):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from md_neighbor_list_b200 import NlistError, _lib
        from md_neighbor_list_b200.parallel import SlabDecomposition
        q, box = _global_system(world)
        dec = SlabDecomposition(world, rank, box, SL, axis=2, slack=0.0)
        dec._cap = 32  # far too small
        q_own, gid_own = dec.partition(q)
        dec.build(None, torch.from_numpy(q_own), gid_owned=torch.from_numpy(gid_own), build_fn=lambda *a: None)
        try:
            dec.check()
            raise AssertionError("overflow not detected")
        except NlistError as e:
            assert e.status == _lib.ERR_CAPACITY
    finally:
        dist.destroy_process_group()


def test_ghost_capacity_overflow_is_detected():
    mp.spawn(_overflow_worker, args=(2, _free_port()), nprocs=2, join=True)


def test_single_rank_is_a_no_op():
    from md_neighbor_list_b200.parallel import SlabDecomposition
    dec = SlabDecomposition(1, 0, (20.0, 20.0, 20.0), SL)
    q = torch.rand(100, 4, dtype=torch.float64) * 20
    g = torch.arange(100, dtype=torch.int32)
    qa, ga, n = dec.exchange(q, g)
    assert n == 100 and qa is q and ga is g and dec.max_ghosts(100) == 0


# ---------------------------------------------------------------------------------------------------------------
# periodic ring: minimum-image rows over world slabs
# ---------------------------------------------------------------------------------------------------------------
def _periodic_system(world: int):
    rng = np.random.default_rng(7)
    box = (11.0, 12.5, 7.5 * world)
    n = 700 * world
    q = np.zeros((n, 4))
    q[:, :3] = rng.random((n, 3)) * np.array(box)
    # particles on the faces and just inside them, on every axis
    q[0, :3] = (0.0, 0.0, 0.0)
    q[1, :3] = (box[0] - 1e-9, box[1] - 1e-9, box[2] - 1e-9)
    q[2, :3] = (5.0, 6.0, 7.5)  # exactly on the seam between slab 0 and slab 1
    return q, box


def _minimum_image_rows(q, box, rows_of):
    L = np.array(box)
    out = []
    for i in rows_of:
        d = q[:, :3] - q[i, :3]
        d -= L * np.round(d / L)
        r2 = (d * d).sum(axis=1)
        m = r2 <= SL * SL
        m[i] = False
        out.append(np.nonzero(m)[0])
    return out


def _periodic_worker(rank: int, world: int, port: int):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from md_neighbor_list_b200.periodic import PeriodicSlabDecomposition
        from oracle import oracle as O
        q, box = _periodic_system(world)
        dec = PeriodicSlabDecomposition(world, rank, box, SL, axis=2)
        q_own, gid_own = dec.partition(q)
        assert q_own.shape[0] > 0
        ext = dec.extended_box()

        def build_fn(q_all, n_owned, gid_all):
            qa, ga = q_all.numpy(), gid_all.numpy()
            assert qa.shape[0] == dec.n_total(n_owned)
            present = ~np.isnan(qa[:, 0])
            assert present[:n_owned].all()
            qa, ga = qa[present], ga[present]
            # every record lies inside the extended box and is a periodic copy of the particle whose id it carries
            assert (qa[:, :3] >= 0).all() and (qa[:, :3] < np.array(ext)).all()
            d = qa[:, :3] - SL - q[ga, :3]
            assert np.abs(d - np.array(box) * np.round(d / np.array(box))).max() < 1e-9
            assert np.abs(qa[:n_owned, :3] - SL - q[ga[:n_owned], :3]).max() < 1e-12  # owned: shifted only
            local = O.build_full(qa, SL, ext)  # open boundary over the extended box
            return _rows(local, range(n_owned), ga), ga[:n_owned]

        rows, gids = dec.build(None, torch.from_numpy(q_own), gid_owned=torch.from_numpy(gid_own), build_fn=build_fn)
        want = _minimum_image_rows(q, box, gids)
        assert len(rows) == len(want)
        for a, b2 in zip(rows, want):
            assert np.array_equal(a, b2)
        cnt = dec.check()
        assert len(cnt) == 6 and all(c > 0 for c in cnt)
        tot = torch.tensor([q_own.shape[0]], dtype=torch.int64)
        dist.all_reduce(tot)
        assert int(tot) == q.shape[0]
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_periodic_ring_gives_minimum_image_rows(world):
    from oracle import oracle as O
    O.lib()
    mp.spawn(_periodic_worker, args=(world, _free_port()), nprocs=world, join=True)


def test_periodic_ring_preconditions():
    from md_neighbor_list_b200.periodic import PeriodicSlabDecomposition
    with pytest.raises(ValueError):
        PeriodicSlabDecomposition(1, 0, (20.0, 20.0, 20.0), SL)          # one GPU: PeriodicVerletList
    with pytest.raises(ValueError):
        PeriodicSlabDecomposition(2, 0, (20.0, 6.0, 20.0), SL)           # an axis shorter than 2 SL
    with pytest.raises(ValueError):
        PeriodicSlabDecomposition(4, 0, (20.0, 20.0, 12.0), SL)          # slabs thinner than SL
