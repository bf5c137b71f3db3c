"""Where a slab rank's step goes (contract workload, 119 164 particles per GPU): packing, NCCL exchange, build.
  torchrun --nproc-per-node N tools/halo_breakdown.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from md_neighbor_list_b200 import VerletListB200, parallel  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=dev)
SL, L = 3.3, 50.0
Slab = parallel.SlabDecomposition if os.environ.get('NLB_HALO', 'p2p') == 'nccl' else parallel.PeerSlabDecomposition
halo = Slab(world, rank, (L, L, L * world), SL, axis=2)
q = halo.local_fcc_slab(1.0, L)
n = q.shape[0]
q_dev, gid_dev = halo.owned_view(n, torch.float64, dev)
q_dev.copy_(torch.from_numpy(q))
gid_dev.copy_(torch.arange(n, dtype=torch.int32, device=dev) + rank * n)
nl = VerletListB200(SL, L, L, L * world, cell_window=halo.cell_window())
nl.initialize(n + halo.max_ghosts(n), int(n * 4.18879 * SL ** 3 * 1.3))
s = torch.cuda.Stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=30):
    out = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        with torch.cuda.stream(s):
            flush.fill_(1)
            a.record(s)
            fn()
            b.record(s)
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b))
    out.sort()
    return out[len(out) // 2]


def full():
    halo.build(nl, q_dev, s, gid_owned=gid_dev)


def exch():
    halo.exchange(q_dev, gid_dev)
    if hasattr(halo, "done"):
        halo.done()  # (inside a full step the build's last kernel does this)


for _ in range(3):
    with torch.cuda.stream(s):
        full()
    nl.synchronize()
qa, ga, no = halo.last_assembled()


nl2 = VerletListB200(SL, L, L, L * world, cell_window=halo.cell_window())  # no halo flags on this handle
nl2.initialize(n + halo.max_ghosts(n), int(n * 4.18879 * SL ** 3 * 1.3))


def build_only():
    nl2.build(qa, n_owned=no, global_ids=ga, stream=s)


res = {"rank": rank, "full_ms": timed(full), "exchange_ms": timed(exch), "build_only_ms": timed(build_only)}
nl.synchronize()
nl2.synchronize()
res["peer_stores"] = bool(getattr(halo, "uses_peer_stores", lambda: False)())
res["ghost_slots"] = int(qa.shape[0] - no)
res["present"] = int((~torch.isnan(qa[:, 0])).sum())
allr = [None] * world
dist.all_gather_object(allr, res)
if rank == 0:
    for r in allr:
        print(json.dumps(r))
halo.check()
if hasattr(halo, 'close'):
    halo.close()
dist.destroy_process_group()
