"""Device check of the periodic ring (md_neighbor_list_b200.periodic.PeriodicSlabDecomposition):
  torchrun --nproc-per-node G --master-addr 127.0.0.1 tools/periodic_ring_check.py
Every rank builds the minimum-image rows of its slab with the CUDA library (FULL and HALF) and compares them with a
numpy minimum-image brute force of the global system; rank 0 prints PERIODIC RING OK."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from md_neighbor_list_b200 import PeriodicSlabDecomposition, VerletListB200  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
dev = torch.device("cuda", torch.cuda.current_device())
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=dev)
SL = 3.3
rng = np.random.default_rng(11)
box = (21.0, 17.5, 9.0 * world)
n = 3000 * world
q = np.zeros((n, 4))
q[:, :3] = rng.random((n, 3)) * np.array(box)
Lb = np.array(box)
ok = True
for mode in ("full_csr", "half_csr"):
    dec = PeriodicSlabDecomposition(world, rank, box, SL, axis=2)
    q_own, gid_own = dec.partition(q)
    nl = VerletListB200(SL, *dec.extended_box(), dtype="f64", mode=mode)
    nl.initialize(dec.n_total(q_own.shape[0]), int(q_own.shape[0] * 4.18879 * SL ** 3 * n / np.prod(Lb) * 1.5) + 4096)
    s = torch.cuda.Stream()
    qd, gd = torch.from_numpy(q_own).to(dev), torch.from_numpy(gid_own).to(dev)
    for _ in range(2):  # the second build replays the library's graph on the same buffers
        dec.build(nl, qd, s, gid_owned=gd)
    st = nl.synchronize()
    dec.check()
    off = nl.offsets().cpu().numpy()
    lst = nl.partners().cpu().numpy()
    for k, i in enumerate(gid_own):
        d = q[:, :3] - q[i, :3]
        d -= Lb * np.round(d / Lb)
        m = (d * d).sum(axis=1) <= SL * SL
        m[i] = False
        want = np.nonzero(m)[0]
        if mode == "half_csr":
            want = want[want > i]
        got = np.sort(lst[off[k]:off[k + 1]])
        if not np.array_equal(got, want):
            ok = False
            print(f"rank {rank} {mode}: row of particle {i} differs: {got[:8]} vs {want[:8]}", flush=True)
            break
    nl.close()
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("PERIODIC RING OK" if int(t) == 1 else "PERIODIC RING FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(t) == 1 else 1)
