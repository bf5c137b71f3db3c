"""Device check of the slab decomposition with peer-store halos (md_neighbor_list_b200.parallel.PeerSlabDecomposition):
  torchrun --nproc-per-node G --master-addr 127.0.0.1 tools/halo_rows_check.py
EVERY rank builds the rows of its slab with the CUDA library — exchange folded into the build (nlb200_set_halo_pack)
and with the separate packing kernel (NLB_HALO_FUSED=0) — for several steps between which the particles MOVE (so the
face populations and the ghost counts change and stale ghost slots must be cleared), and compares counts, offsets and
row-sorted partners with the oracle run on the global system; then the incremental halo refresh (nlb200_halo_refresh):
the recorded face set re-sent at new positions, every ghost slot against the neighbour's current record.  Rank 0 prints
HALO ROWS OK."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from md_neighbor_list_b200 import VerletListB200, parallel  # noqa: E402
from oracle import oracle as O  # noqa: E402  (the checker)

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
dev = torch.device("cuda", torch.cuda.current_device())
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=dev)
SL = 3.3
box = (20.0, 16.7, 13.4 * world)  # slab faces do NOT coincide with cell faces
n_per = 4000
ok = True
for fused in ("1", "0"):
    os.environ["NLB_HALO_FUSED"] = fused
    rng = np.random.default_rng(5)
    halo = parallel.PeerSlabDecomposition(world, rank, box, SL, axis=2)
    # global system: every rank draws the same particles; a rank owns those of its slab, in global-id order
    qg = np.zeros((n_per * world, 4))
    qg[:, :3] = rng.random((n_per * world, 3)) * np.array(box)
    vel = (rng.random((n_per * world, 3)) - 0.5) * 0.8
    # ownership is fixed by the initial positions
    own = np.nonzero(halo.owns(qg))[0]
    n_own = own.shape[0]
    q_dev, gid_dev = halo.owned_view(n_own, torch.float64, dev)
    gid_dev.copy_(torch.from_numpy(own.astype(np.int32)))
    nl = VerletListB200(SL, *box, dtype="f64", mode="full_csr", cell_window=halo.cell_window())
    nl.initialize(n_own + halo.max_ghosts(n_own), int(n_own * 4.18879 * SL ** 3 * n_per * world / np.prod(box) * 1.6) + 4096)
    s = torch.cuda.Stream()
    for step in range(4):
        q_now = qg.copy()
        q_now[:, :3] += vel * (0.6 * step)
        q_now[:, :3] = np.clip(q_now[:, :3], 0.0, np.array(box) - 1e-9)
        # a particle stays inside its owner's slab (it may enter or leave the band within SL of a face: the face
        # populations and ghost counts change from step to step, the ownership does not)
        th = box[2] / world
        slab = np.floor(qg[:, 2] / th)
        q_now[:, 2] = np.clip(q_now[:, 2], slab * th, (slab + 1) * th - 1e-9)
        q_dev.copy_(torch.from_numpy(np.ascontiguousarray(q_now[own])))
        torch.cuda.synchronize()
        dist.barrier()
        for _ in range(2):  # the second build replays the library's graph on the same buffers
            halo.build(nl, q_dev, s, gid_owned=gid_dev)
        st = nl.synchronize()
        halo.check()
        # every particle within SL of this slab must have arrived: compare with the oracle on the GLOBAL system
        ref = O.build_full(np.ascontiguousarray(q_now), SL, box).sorted_rows()
        cnt = nl.number_of_partners().cpu().numpy()
        off = nl.offsets().cpu().numpy()
        lst = nl.partners().cpu().numpy().copy()
        O.lib().orc_sort_rows(lst.ctypes.data, off.ctypes.data, n_own)
        good = np.array_equal(cnt, ref.number_of_partners[own])
        if good:
            for k in (list(range(0, n_own, 7)) + [n_own - 1]):
                i = own[k]
                if not np.array_equal(lst[off[k]:off[k + 1]], ref.partners[ref.offsets[i]:ref.offsets[i + 1]]):
                    good = False
                    break
        if not good:
            ok = False
            print(f"rank {rank} fused={fused} step {step}: rows differ from the oracle", flush=True)
            break
    if ok and fused == "1":
        # incremental halo refresh: the particles move a little (the list would stay valid), the face
        # set recorded by the last build is re-sent: every ghost slot must hold the CURRENT position of its global id,
        # unused slots stay absent, global ids are untouched; a build afterwards continues the flag protocol
        qa0, _, no0 = halo.last_assembled()
        present0 = ~np.isnan(qa0.cpu().numpy()[:, 0])
        present0[:no0] = False
        for rep in range(2):
            q_new = q_now.copy()
            q_new[:, :2] = np.clip(q_new[:, :2] + (rep + 1) * 0.05 * vel[:, :2], 0.0, np.array(box[:2]) - 1e-9)
            q_dev.copy_(torch.from_numpy(np.ascontiguousarray(q_new[own])))
            torch.cuda.synchronize()
            dist.barrier()
            qa, ga, no = halo.refresh(nl, q_dev, s)
            s.synchronize()
            halo.check()
            qh, gh = qa.cpu().numpy(), ga.cpu().numpy()
            halo.done()
            present = ~np.isnan(qh[:, 0])
            present[:no] = False
            if not np.array_equal(present, present0) or not np.array_equal(qh[present], q_new[gh[present]]):
                ok = False
                print(f"rank {rank}: refreshed ghosts differ from the neighbours' current positions", flush=True)
                break
        if ok:
            q_dev.copy_(torch.from_numpy(np.ascontiguousarray(q_new[own])))
            torch.cuda.synchronize()
            dist.barrier()
            halo.build(nl, q_dev, s, gid_owned=gid_dev)
            nl.synchronize()
            halo.check()
            ref = O.build_full(np.ascontiguousarray(q_new), SL, box).sorted_rows()
            if not np.array_equal(nl.number_of_partners().cpu().numpy(), ref.number_of_partners[own]):
                ok = False
                print(f"rank {rank}: build after the refresh differs from the oracle", flush=True)
    nl.close()
    halo.close()
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("HALO ROWS OK" if int(t) == 1 else "HALO ROWS FAILED", flush=True)
dist.destroy_process_group()
