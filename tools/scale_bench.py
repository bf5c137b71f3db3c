"""Weak-scaling study of BASELINE.json configs[3]: every rank owns one slab of the reference's
jittered FCC lattice, sx x sy x sz lattice cells (4 atoms each), slabs stacked along z, ghost exchange over NCCL.
  torchrun --nproc-per-node G tools/scale_bench.py --sx 320 --sy 320 --sz 40 --steps 5     (16 384 000 particles/GPU)
Prints one JSON line on rank 0: ms per build (max over ranks, CUDA events, L2 cold by size), entries/s, ghosts."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from md_neighbor_list_b200 import VerletListB200, _lib, parallel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sx", type=int, default=320)
ap.add_argument("--sy", type=int, default=320)
ap.add_argument("--sz", type=int, default=40)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--mode", default="full_csr")
args = ap.parse_args()

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)

SL, DENS = 3.3, 1.0
s = (0.25 * DENS) ** (-1.0 / 3.0)  # lattice constant, make_list.cpp:54
Lx, Ly, Lz_slab = args.sx * s, args.sy * s, args.sz * s
box = (Lx, Ly, Lz_slab * world)
L = _lib.lib()
n = L.nlb200_workload_fcc(DENS, 1.0, args.sx, args.sy, args.sz, 2 + rank, None, 4, 0)
q = np.zeros((n, 4), dtype=np.float64)
assert L.nlb200_workload_fcc(DENS, 1.0, args.sx, args.sy, args.sz, 2 + rank, q.ctypes.data, 4, n) == n
q[:, 2] += rank * Lz_slab
halo = parallel.SlabDecomposition(world, rank, box, SL, axis=2) if world > 1 else None
gid_dev = None
if halo is None:
    q_dev = torch.from_numpy(q).to(dev)
else:
    # particles live in the head of the halo assembly buffer: no device-to-device copy per exchange
    q_dev, gid_dev = halo.owned_view(n, torch.float64, dev)
    q_dev.copy_(torch.from_numpy(q))
    gid_dev.copy_(torch.arange(n, dtype=torch.int32, device=dev) + rank * n)
del q
n_total = n + (halo.max_ghosts(n) if halo else 0)
nl = VerletListB200(SL, *box, dtype="f64", mode=args.mode, kernel_variant=int(os.environ.get("NLB_VARIANT", "0")))
per_row = 4.18879 * SL ** 3 * DENS * (0.5 if args.mode == "half_csr" else 1.0)
nl.initialize(n_total, int(n * per_row * 1.05) + 1024)
stream = torch.cuda.Stream()


def one():
    if halo is None:
        nl.build(q_dev, stream=stream)
    else:
        halo.build(nl, q_dev, stream, gid_owned=gid_dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def sync_growing():
    from md_neighbor_list_b200 import NlistError
    for _ in range(4):
        try:
            return nl.synchronize()
        except NlistError as e:
            if e.status == _lib.ERR_CAPACITY:
                nl.reserve(nl.stats().required_entries)
            elif e.status == _lib.ERR_CELL_CAPACITY:
                nl.reserve_cell_capacity(nl.stats().max_in_cell)
            else:
                raise
            with torch.cuda.stream(stream):
                one()
    raise RuntimeError("capacity retries exhausted")


with torch.cuda.stream(stream):
    one()
sync_growing()
with torch.cuda.stream(stream):
    for _ in range(args.warmup):
        one()
st = nl.synchronize()
sent = halo.check() if halo else (0, 0)
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
barrier()
with torch.cuda.stream(stream):
    for a, b in ev:
        a.record(stream)
        one()
        b.record(stream)
barrier()
ms = sorted(a.elapsed_time(b) for a, b in ev)
# per-stage device times on rank 0: a profiled handle rebuilds from the assembled records (no exchange: that needs all ranks)
stage_ms = {}
if rank == 0:
    nl.close()
    torch.cuda.empty_cache()
    nlp = VerletListB200(SL, *box, dtype="f64", mode=args.mode, profile=True)
    nlp.initialize(n_total, int(n * per_row * 1.05) + 1024)
    for r in range(3):
        with torch.cuda.stream(stream):
            if halo is None:
                nlp.build(q_dev, stream=stream)
            else:
                qa, ga, no = halo.last_assembled()
                nlp.build(qa, n_owned=no, global_ids=ga, stream=stream)
        nlp.synchronize()
    stage_ms = {k: round(v, 4) for k, v in nlp.stage_times().items()}
    nlp.close()
v = torch.tensor([ms[len(ms) // 2]], dtype=torch.float64, device=dev)
tot = torch.tensor([float(st.number_of_pairs), float(n)], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(v, op=dist.ReduceOp.MAX)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
if rank == 0:
    t = float(v.item())
    print(json.dumps({"workload": f"FCC {args.sx}x{args.sy}x{args.sz} lattice cells per GPU, density 1.0, SL 3.3",
                      "mode": args.mode, "n_gpus": world, "particles_per_gpu": n, "particles": int(tot[1].item()),
                      "entries": int(tot[0].item()), "ms_per_build": t,
                      "entries_per_s": float(tot[0].item()) / (t * 1e-3),
                      "particles_per_s": float(tot[1].item()) / (t * 1e-3),
                      "ghost_capacity_per_face": halo.ghost_capacity(n) if halo else 0,
                      "ghosts_sent_rank0": list(sent), "max_in_cell": st.max_in_cell,
                      "max_partners": st.max_partners, "stage_ms_rank0": stage_ms}))
if world > 1:
    dist.destroy_process_group()
