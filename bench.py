#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 Verlet-list builder.

  python bench.py --gpus 1 --steps K --warmup W            # our arm (1 GPU)
  torchrun ... bench.py --gpus N --steps K --warmup W       # our arm, N ranks (slab decomposition + NCCL halo)
  python bench.py --impl reference --steps K --warmup W     # the reference's own CPU classes (oracle/_ref)

A "step" is one complete list build (cell binning, pair search, CSR emission) of one batch of synthetic particles.
N=1 workload = BASELINE.json configs[1]: the reference default system at density 1.0 (make_list.cpp:17-24: L=50,
search length 3.3, jittered FCC from mt19937(2), N=119164), double precision, FULL list (GPU semantics,
kernel_impl.cuh:3-35) in CSR.  N>1: the same system replicated along z (one 50^3 slab per GPU) with ghost exchange.

metric  = unordered neighbour pairs listed per second (whole job); the CPU reference lists the same pairs (half list).
value   = positions already resident in HBM when the timed region starts; L2 is flushed between timed builds.
e2e     = same build through the public API from pinned HOST buffers: H2D of positions, build, D2H of counts,
          offsets and the partner list inside the timed region.
"""
from __future__ import annotations

import argparse
import faulthandler
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SL = 3.3
L_DEFAULT = 50.0
METRIC = "neighbor_pairs_listed_per_s"
UNIT = "pairs/s"


# ---------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int = 0):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index
        self._t = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self._t = threading.Thread(target=self._read, daemon=True)
        self._t.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.rows:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                 f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# CPU baseline (oracle/_ref = the reference's own classes; falls back to the oracle port)
# ---------------------------------------------------------------------------------------------------------------
def cpu_reference_run(q, loops: int, variants=None, warmup: int = 2):
    """Times the reference's CPU classes on q — `warmup` untimed builds, then `loops` timed builds on ONE instance,
    the reference's own LOOP protocol (make_list.cpp:152-157); returns {variant: ms_per_build}, pairs, kind."""
    from oracle import oracle as O
    out = {}
    pairs = None
    kind = "reference"
    variants = variants or ["avx512_8x1", "avx2_4x1", "scalar"]
    for v in variants:
        if not O.ref_available(v):
            continue
        r, ms = O.ref_build(v, q, SL, (L_DEFAULT, L_DEFAULT, L_DEFAULT), loops=loops, warmup=warmup)
        out[v] = ms
        pairs = r.number_of_pairs
    if not out:  # oracle/_ref absent (should not happen on the GPU box: the .so files travel)
        kind = "port"
        t0 = time.perf_counter()
        for _ in range(loops):
            r = O.build_half(q, SL, (L_DEFAULT, L_DEFAULT, L_DEFAULT))
        out["oracle_port_scalar"] = (time.perf_counter() - t0) * 1e3 / loops
        pairs = r.number_of_pairs
    return out, pairs, kind


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    q = O.gen_fcc(1.0)
    # pick the fastest variant with a short probe, then ONE instance of it: `warmup` untimed builds followed by
    # `steps` timed builds back to back — the reference's own LOOP protocol (make_list.cpp:152-157).  A fresh
    # instance per timed build would charge every build the first-touch page faults of the 143 MB pair buffers
    # (79.9 ms/build measured that way vs 57.2 ms in steady state).
    probe, pairs, kind = cpu_reference_run(q, 2, warmup=1)
    best_name = min(probe, key=probe.get)
    timed, pairs, kind = cpu_reference_run(q, max(args.steps, 1), [best_name], warmup=max(args.warmup, 1))
    ms = timed[best_name]
    val = pairs / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": kind,
                         "sample": f"{args.steps} full builds (after {max(args.warmup, 1)} untimed ones on the same "
                                   f"instance) of the density-1.0 default system with the reference's "
                                   f"{best_name} class (single-threaded by construction); probe ms/build: "
                                   + ", ".join(f"{k}={v:.1f}" for k, v in probe.items())},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "host": {"nproc": os.cpu_count()},
    }
    print(json.dumps(line))


def workload_config(n_gpus: int) -> dict:
    return {"workload": "reference default system (make_list.cpp:17-24): jittered FCC, density 1.0, L=50, "
                        "search length 3.3 (rc 3.0 + margin 0.3), N=119164 per GPU"
                        + ("" if n_gpus == 1 else f", {n_gpus} slabs stacked along z with ghost exchange"),
            "list": "full (both directions), CSR, int64 offsets", "particles_per_gpu": 119164,
            "l2": "flushed between timed builds (256 MiB write)", "parallelism": f"slab{n_gpus}"}



# ---------------------------------------------------------------------------------------------------------------
# helpers of our arm
# ---------------------------------------------------------------------------------------------------------------
def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def _fp32_issue_rate(dev) -> float:
    """lane-FMA per second of the whole chip: the measured 126 lane-FMA/clk/SM (profiles/r01_microbench_issue_rates.txt,
    re-read here) x the SM count and the SM clock the device reports."""
    import torch
    rate = 126.0
    try:
        with open(os.path.join(ROOT, "profiles", "r01_microbench_issue_rates.txt")) as f:
            for line in f:
                if line.startswith("FFMA ") and "lane-FMA/clk/SM" in line:
                    rate = float(line.split("(")[1].split()[0])
    except (OSError, ValueError, IndexError):
        pass
    props = torch.cuda.get_device_properties(dev)
    clk = getattr(props, "clock_rate", None)
    ghz = (clk * 1e3) if clk else 1.965e9
    return rate * props.multi_processor_count * ghz


def time_builds(torch, nl, q_dev, stream, flush, reps, **bk):
    """median ms of `reps` graph-replayed builds, L2 flushed before each (CUDA events on the build's stream)"""
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            if flush is not None:
                flush.fill_(1)
            e0.record(stream)
            nl.build(q_dev, stream=stream, **bk)
            e1.record(stream)
        nl.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[len(ms) // 2]


def build_growing(torch, nl, q_dev, stream, **bk):
    """first build of a handle: capacities the estimate missed are grown and the build repeated"""
    from md_neighbor_list_b200 import NlistError, _lib
    for _ in range(6):
        with torch.cuda.stream(stream):
            nl.build(q_dev, stream=stream, **bk)
        try:
            return nl.synchronize()
        except NlistError as e:
            if e.status == _lib.ERR_CAPACITY:
                nl.reserve(nl.stats().required_entries)
            elif e.status == _lib.ERR_CELL_CAPACITY:
                nl.reserve_cell_capacity(nl.stats().max_in_cell)
            else:
                raise
    raise RuntimeError("capacity retries exhausted")


def side_workloads(torch, dev, stream, flush, peak):
    """The other inputs BASELINE.json's metric names, each timed like the headline (graph replay, L2 flushed):
    density 0.5 (FULL and the CPU classes' HALF list), density 1.0 HALF, uniform-random particles at both densities."""
    from md_neighbor_list_b200 import VerletListB200, workloads
    out = []
    n_uni = 1 << 20
    cases = [("reference default system, density 0.5", lambda: workloads.fcc(0.5, L_DEFAULT), L_DEFAULT, "full_csr"),
             ("reference default system, density 0.5", lambda: workloads.fcc(0.5, L_DEFAULT), L_DEFAULT, "half_csr"),
             ("reference default system, density 1.0", lambda: workloads.fcc(1.0, L_DEFAULT), L_DEFAULT, "half_csr"),
             ("uniform random, density 1.0, N=2^20", lambda: workloads.uniform(n_uni, round(n_uni ** (1 / 3))),
              float(round(n_uni ** (1 / 3))), "full_csr"),
             ("uniform random, density 0.5, N=2^20", lambda: workloads.uniform(n_uni, round((2 * n_uni) ** (1 / 3))),
              float(round((2 * n_uni) ** (1 / 3))), "full_csr")]
    for name, gen, L, mode in cases:
        q = gen()
        n = q.shape[0]
        qd = torch.from_numpy(q).to(dev)
        nl = VerletListB200(SL, L, L, L, dtype="f64", mode=mode)
        nl.initialize(n)
        st = build_growing(torch, nl, qd, stream)
        for _ in range(3):
            time_builds(torch, nl, qd, stream, flush, 1)
        ms = time_builds(torch, nl, qd, stream, flush, 9)
        entries = st.number_of_pairs
        pairs = entries if mode == "half_csr" else entries // 2
        b_alg = n * 40 + 4 * entries
        out.append({"workload": name, "list": mode, "particles": n, "ms_per_build": ms, "entries": entries,
                    "pairs_per_s": pairs / (ms * 1e-3), "entries_per_s": entries / (ms * 1e-3),
                    "candidate_tests_per_s": st.candidates_tested / (ms * 1e-3),
                    "algorithmic_gbs": b_alg / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": b_alg / (ms * 1e-3) / 1e9 / peak})
        nl.close()
        del qd
        torch.cuda.empty_cache()
    return out


def c2_block(torch, dev, stream, peak):
    """BASELINE.json configs[2]: 2^24 uniform-random particles, density 1.0, SL 3.3 — the HBM-roofline study.  Inputs
    larger than L2 (no flush needed); per-stage times from a profiled handle."""
    from md_neighbor_list_b200 import VerletListB200, workloads
    n = 1 << 24
    L = float(round(n ** (1.0 / 3.0)))
    q = workloads.uniform(n, L)
    qd = torch.from_numpy(q).to(dev)
    del q
    out = {"workload": "uniform random, N=2^24, density 1.0, L=256, search length 3.3; FULL list, CSR", "particles": n}
    for profile in (False, True):
        nl = VerletListB200(SL, L, L, L, dtype="f64", mode="full_csr", profile=profile)
        nl.initialize(n)
        st = build_growing(torch, nl, qd, stream)
        time_builds(torch, nl, qd, stream, None, 1)
        if not profile:
            ms = time_builds(torch, nl, qd, stream, None, 5)
            entries = st.number_of_pairs
            b_alg = n * 40 + 4 * entries
            out.update({"ms_per_build": ms, "entries": entries, "pairs_per_s": entries / 2 / (ms * 1e-3),
                        "candidate_tests_per_s": st.candidates_tested / (ms * 1e-3), "band_retests": st.band_tests,
                        "algorithmic_bytes": b_alg, "algorithmic_gbs": b_alg / (ms * 1e-3) / 1e9,
                        "frac_of_hbm_peak": b_alg / (ms * 1e-3) / 1e9 / peak, "max_partners": st.max_partners,
                        "l2": "inputs and outputs (10.6 GB) exceed L2: no flush"})
        else:
            time_builds(torch, nl, qd, stream, None, 2)
            out["stage_ms"] = {k: round(v, 4) for k, v in nl.stage_times().items()}
        nl.close()
        torch.cuda.empty_cache()
    return out


def gpu_reference_run():
    """The reference's own GPU path on this box (BASELINE.json configs[1]): make_list.cu + neighlist_gpu.hpp +
    kernel_impl.cuh, unmodified, compiled for sm_100 by oracle/Makefile (target refgpu) with -DUSE_WARP_UNROLL_SMEM,
    the README's headline variant.  Its own driver protocol: LOOP = 100 builds, then its O(N^2) self-test."""
    exe = os.path.join(ROOT, "oracle", "_ref", "gpu", "make_list_gpu_warp_unroll_smem.out")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/gpu not built (make -C oracle refgpu needs /root/reference)"}
    try:
        r = subprocess.run([exe, "128", "7"], capture_output=True, text=True, timeout=240)
    except (OSError, subprocess.TimeoutExpired) as e:
        return {"unavailable": repr(e)[:200]}
    out = {"variant": "make_neighlist_warp_unroll_smem (kernel_impl.cuh:363-436), tblock 128",
           "self_test": "TEST is passed." if "TEST is passed." in r.stderr else "FAILED: " + r.stderr[-200:]}
    for line in r.stdout.splitlines():
        if line.startswith("# of particles"):
            f = line.replace("[ms]", "").split()
            out["particles"] = int(f[3])
            out["ms_per_build"] = float(f[4]) / 100.0
            out["pairs_per_s"] = 7839886 / (out["ms_per_build"] * 1e-3) if out["particles"] == 119164 else None
    return out


def c3_block(torch, dist, dev, stream, rank, world, local_rank):
    """BASELINE.json configs[3]: jittered FCC, 320 x 320 x 40 lattice cells (16 384 000 particles) per GPU, slabs
    stacked along z, ghost exchange over NCCL; weak-scaling efficiency against the 1-GPU time of the same slab measured
    in the same process group (rank 0 builds its slab alone, without ghosts)."""
    from md_neighbor_list_b200 import VerletListB200, _lib, parallel
    sx, sy, sz = 320, 320, 40
    s_lat = (0.25 * 1.0) ** (-1.0 / 3.0)
    Lx, Ly, Lz = sx * s_lat, sy * s_lat, sz * s_lat
    box = (Lx, Ly, Lz * world)
    Lb = _lib.lib()
    import numpy as np
    n = Lb.nlb200_workload_fcc(1.0, 1.0, sx, sy, sz, 2 + rank, None, 4, 0)
    q = np.zeros((n, 4), dtype=np.float64)
    assert Lb.nlb200_workload_fcc(1.0, 1.0, sx, sy, sz, 2 + rank, q.ctypes.data, 4, n) == n
    q[:, 2] += rank * Lz
    Slab = parallel.SlabDecomposition if os.environ.get("NLB_HALO", "p2p") == "nccl" else parallel.PeerSlabDecomposition
    halo = Slab(world, rank, box, SL, axis=2)
    q_dev, gid_dev = halo.owned_view(n, torch.float64, dev)
    q_dev.copy_(torch.from_numpy(q))
    gid_dev.copy_(torch.arange(n, dtype=torch.int32, device=dev) + rank * n)
    del q
    n_total = n + halo.max_ghosts(n)
    per_row = 4.18879 * SL ** 3
    nl = VerletListB200(SL, *box, dtype="f64", mode="full_csr", cell_window=halo.cell_window())
    nl.initialize(n_total, int(n * per_row * 1.05) + 1024)

    def one():
        halo.build(nl, q_dev, stream, gid_owned=gid_dev)

    from md_neighbor_list_b200 import NlistError
    for _ in range(6):
        with torch.cuda.stream(stream):
            one()
        try:
            st = nl.synchronize()
            break
        except NlistError as e:
            if e.status == _lib.ERR_CAPACITY:
                nl.reserve(nl.stats().required_entries)
            elif e.status == _lib.ERR_CELL_CAPACITY:
                nl.reserve_cell_capacity(nl.stats().max_in_cell)
            else:
                raise
    with torch.cuda.stream(stream):
        one()
    st = nl.synchronize()
    halo.check()
    steps = 5
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    dist.barrier()
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        for a, b in ev:
            a.record(stream)
            one()
            b.record(stream)
    dist.barrier()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)[steps // 2]
    entries = float(nl.synchronize().number_of_pairs)
    nl.close()
    torch.cuda.empty_cache()
    # the same slab alone on rank 0 (no ghosts, no exchange): the 1-GPU time of the weak-scaling ratio
    ms1 = 0.0
    if rank == 0:
        nl1 = VerletListB200(SL, Lx, Ly, Lz, dtype="f64", mode="full_csr")
        nl1.initialize(n, int(n * per_row * 1.05) + 1024)
        build_growing(torch, nl1, q_dev, stream)
        time_builds(torch, nl1, q_dev, stream, None, 1)
        ms1 = time_builds(torch, nl1, q_dev, stream, None, 3)
        nl1.close()
    v = torch.tensor([ms, ms1], dtype=torch.float64, device=dev)
    tot = torch.tensor([entries, float(n)], dtype=torch.float64, device=dev)
    dist.all_reduce(v, op=dist.ReduceOp.MAX)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    transport = "peer stores (CUDA IPC over NVLink)" if getattr(halo, "uses_peer_stores", lambda: False)() else "NCCL send/recv"
    del q_dev, gid_dev
    if hasattr(halo, "close"):
        halo.close()
    del halo
    torch.cuda.empty_cache()
    ms, ms1 = float(v[0]), float(v[1])
    return {"workload": f"jittered FCC, {sx}x{sy}x{sz} lattice cells per GPU (16 384 000 particles), density 1.0, "
                        f"slabs along z, ghost exchange over NCCL; FULL list",
            "n_gpus": world, "particles": int(tot[1]), "entries": int(tot[0]), "ms_per_build": ms,
            "ms_per_build_one_slab_alone": ms1, "weak_scaling_efficiency": ms1 / ms if ms > 0 else None,
            "pairs_per_s": float(tot[0]) / 2 / (ms * 1e-3), "l2": "inputs and outputs exceed L2: no flush",
            "halo": transport}


def check_rank_rows(torch, nl, halo, n_owned):
    """Outside the timed region: the rows this rank just built (device path, NCCL ghosts) against the oracle run on the
    records the exchange assembled — counts, offsets and row-sorted partners, element by element."""
    import numpy as np
    from oracle import oracle as O
    nl.synchronize()  # the e2e loop left builds in flight: the accessors need a synchronized result
    qa, ga, no = halo.last_assembled()
    qh = qa.cpu().numpy()
    gh = ga.cpu().numpy()
    present = ~np.isnan(qh[:, 0])
    present[:no] = True
    idx = np.nonzero(present)[0]
    ref = O.build_full(np.ascontiguousarray(qh[idx]), SL, halo.box).sorted_rows()
    cnt = nl.number_of_partners().cpu().numpy()
    off = nl.offsets().cpu().numpy()
    lst = nl.partners().cpu().numpy().copy()
    O.lib().orc_sort_rows(lst.ctypes.data, off.ctypes.data, no)
    ok = np.array_equal(cnt, ref.number_of_partners[:no]) and int(off[-1]) == int(ref.offsets[no])
    if ok:
        want = gh[idx][ref.partners[:ref.offsets[no]]]  # oracle partners are positions in the compacted array
        o2 = ref.offsets[:no + 1].copy()
        O.lib().orc_sort_rows(want.ctypes.data, o2.ctypes.data, no)
        ok = np.array_equal(lst, want)
    return bool(ok)


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from md_neighbor_list_b200 import VerletListB200, workloads

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)

    L = L_DEFAULT
    if world == 1:
        q = workloads.fcc(1.0, L)
        n_owned = q.shape[0]
        box = (L, L, L)
        halo = None
    else:
        from md_neighbor_list_b200 import parallel
        # ghosts travel as peer stores of the packing kernel (CUDA IPC over NVLink); NLB_HALO=nccl: grouped send/recv
        Slab = parallel.SlabDecomposition if os.environ.get("NLB_HALO", "p2p") == "nccl" else parallel.PeerSlabDecomposition
        halo = Slab(world, rank, box=(L, L, L * world), search_length=SL, axis=2)
        q = halo.local_fcc_slab(1.0, L)  # this rank's owned particles (global order), host
        n_owned = q.shape[0]
        box = (L, L, L * world)

    stream = torch.cuda.Stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def flush_l2():
        flush.fill_(1)

    q_pinned = torch.from_numpy(q).pin_memory()
    gid_dev = None
    if halo is None:
        q_dev = q_pinned.to(dev, non_blocking=False)
    else:
        # particles live in the head of the halo assembly buffer: no device-to-device copy per exchange
        q_dev, gid_dev = halo.owned_view(n_owned, torch.float64, dev)
        q_dev.copy_(q_pinned)
        gid_dev.copy_(torch.arange(n_owned, dtype=torch.int32, device=dev) + rank * n_owned)

    window = halo.cell_window() if halo is not None else None  # a slab rank bins only its window of the global grid
    nl = VerletListB200(SL, *box, dtype="f64", mode="full_csr", cell_window=window)
    cap_particles = n_owned if halo is None else n_owned + halo.max_ghosts(n_owned)
    # the library estimates the list size from particles / box volume; a rank of a slab decomposition holds 1/world of
    # the global box, so it is told: density 1.0 * (4/3) pi SL^3 entries per owned row, +30 %
    max_entries = 0 if halo is None else int(n_owned * 4.18879 * SL ** 3 * 1.3)
    nl.initialize(cap_particles, max_entries)

    graphed = None  # set after the capacities have settled (below)

    def one_build(qd_local=None):
        nonlocal_graphed = graphed
        if halo is None:
            nl.build(q_dev, stream=stream)
        elif nonlocal_graphed is not None:
            nonlocal_graphed.step()
        else:
            halo.build(nl, q_dev, stream, gid_owned=gid_dev)

    def sync_growing():
        """synchronize; a capacity the estimate missed (list entries, particles per cell) is grown and the build
        repeated — outside the timed region"""
        from md_neighbor_list_b200 import NlistError, _lib
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
                    one_build()
        raise RuntimeError("capacity retries exhausted")

    # ---- warm-up ----
    with torch.cuda.stream(stream):
        one_build()
    sync_growing()
    if halo is not None and os.environ.get("NLB_HALO_GRAPH", "0") == "1":
        # opt-in experiment (measured at 2 GPUs: hot 0.234 ms vs 0.238 eager, cold 0.31 vs 0.25, and the process group
        # was slow to shut down after NCCL ops had been captured): exchange + build of identical steps as ONE CUDA graph (packing kernels, NCCL group, the build's kernel chain);
        # captured only now that no capacity will be reallocated
        g = parallel.GraphedHaloBuild(halo, nl, q_dev, gid_dev, stream)
        if g.graph is not None:
            graphed = g
        elif rank == 0:
            print("halo graph capture refused, eager steps: " + getattr(g, "error", "?")[:300], file=sys.stderr)
    # the clock sampler starts before the warm-up steps so that nothing idles the GPU between them and the timed region;
    # warm-up steps are shaped like timed ones (L2 flush + build)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            flush_l2()
            one_build()
    st = nl.synchronize()
    if halo is not None:
        halo.check()  # ghost-capacity overflow would have dropped ghosts
    pairs_local = st.number_of_pairs // 2
    entries_local = st.number_of_pairs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region: K builds, device-timed, L2 flushed before each ----
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t0 = time.time()
    with torch.cuda.stream(stream):
        for k in range(args.steps):
            flush_l2()
            ev[k][0].record(stream)
            one_build()
            ev[k][1].record(stream)
    barrier()
    t1 = time.time()
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    ms_total = float(sum(ms_steps))
    # hot-L2 protocol of the reference drivers (LOOP identical builds back to back, make_list.cu:124-130)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.steps):
            one_build()
        e1.record(stream)
    barrier()
    ms_hot = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop(t0, t1)

    # ---- e2e: host buffers in, host buffers out ----
    st = nl.synchronize()
    n_entries = st.number_of_pairs
    h_cnt = torch.empty(n_owned, dtype=torch.int32).pin_memory()
    h_off = torch.empty(n_owned + 1, dtype=torch.int64).pin_memory()
    h_lst = torch.empty(n_entries, dtype=torch.int32).pin_memory()
    q_dev2 = torch.empty_like(q_dev) if halo is None else q_dev  # multi-GPU: H2D straight into the assembly buffer
    # borrowed views of the library's output buffers (stable until reserve/destroy)
    v_cnt, v_off, v_lst = nl.number_of_partners(), nl.offsets(), nl.partners()

    def one_e2e():
        q_dev2.copy_(q_pinned, non_blocking=True)
        if halo is None:
            nl.build(q_dev2, stream=stream)
        else:
            halo.build(nl, q_dev2, stream, gid_owned=gid_dev)
        h_cnt.copy_(v_cnt, non_blocking=True)
        h_off.copy_(v_off, non_blocking=True)
        h_lst.copy_(v_lst, non_blocking=True)

    with torch.cuda.stream(stream):
        for _ in range(2):
            one_e2e()
    barrier()
    e2e_steps = max(3, min(args.steps, 20))
    w0 = time.perf_counter()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(e2e_steps):
            one_e2e()
        e1.record(stream)
    barrier()
    w1 = time.perf_counter()
    ms_e2e = max(e0.elapsed_time(e1), (w1 - w0) * 1e3) / e2e_steps
    assert int(h_off[-1]) == n_entries and int(h_cnt.sum()) == n_entries

    # ---- dominant kernel: per-stage device times (separate, profiled handle; L2 flushed) ----
    # rank 0 only and WITHOUT a ghost exchange (a P2P step needs every rank): it rebuilds from the records the last
    # exchange assembled
    stage_ms = {}
    if rank == 0:
        nlp = VerletListB200(SL, *box, dtype="f64", mode="full_csr", profile=True, cell_window=window)
        nlp.initialize(cap_particles, max_entries)
        reps = 10
        for r in range(reps + 2):
            with torch.cuda.stream(stream):
                flush_l2()
                if halo is None:
                    nlp.build(q_dev, stream=stream)
                else:
                    qa, ga, no = halo.last_assembled()
                    nlp.build(qa, n_owned=no, global_ids=ga, stream=stream)
            nlp.synchronize()
            if r >= 2:
                for k, v in nlp.stage_times().items():
                    stage_ms[k] = stage_ms.get(k, 0.0) + v / reps
        nlp.close()

    # ---- outside the timed region: this rank's rows (device path, NCCL ghosts) against the oracle ----
    rows_ok = None
    if halo is not None and rank == 0 and not args.no_extras:
        rows_ok = check_rank_rows(torch, nl, halo, n_owned)

    # ---- the other workloads BASELINE.json names (after the headline measurement, own handles) ----
    peak, peak_src = _peaks()
    extras = {}
    h2d_bytes = int(q_pinned.numel() * 8)
    d2h_bytes = int(h_cnt.numel() * 4 + h_off.numel() * 8 + h_lst.numel() * 4)
    if not args.no_extras:
        nl.close()
        del h_lst, v_lst, v_cnt, v_off
        torch.cuda.empty_cache()
        if world == 1:
            extras["workloads"] = side_workloads(torch, dev, stream, flush, peak)
            del flush
            torch.cuda.empty_cache()
            extras["c2"] = c2_block(torch, dev, stream, peak)
            if not args.no_cpu_baseline:
                extras["gpu_reference"] = gpu_reference_run()
        else:
            del flush
            torch.cuda.empty_cache()
            extras["c3"] = c3_block(torch, dist, dev, stream, rank, world, local_rank)
            extras["rank0_rows_match_oracle"] = rows_ok

    halo_transport = None
    if halo is not None:
        halo_transport = ("ghost exchange by peer stores of the build's binning kernel (CUDA IPC over NVLink; "
                          "nlb200_set_halo_pack), no NCCL call and no packing launch on the step"
                          if getattr(halo, "uses_peer_stores", lambda: False)() and os.environ.get("NLB_HALO_FUSED", "1") != "0"
                          else "ghost exchange by peer stores of the packing kernel (CUDA IPC over NVLink), no NCCL call "
                          "on the step" if getattr(halo, "uses_peer_stores", lambda: False)()
                          else "ghost exchange by one grouped NCCL send/recv")
        if hasattr(halo, "close"):
            halo.close()

    # ---- reduce over ranks ----
    vals = torch.tensor([ms_total, ms_hot, ms_e2e], dtype=torch.float64, device=dev)
    sums = torch.tensor([float(pairs_local), float(entries_local)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    ms_total, ms_hot, ms_e2e = [float(v) for v in vals.cpu()]
    pairs_all, entries_all = [float(v) for v in sums.cpu()]

    if rank == 0:
        ms_step = ms_total / args.steps
        value = pairs_all / (ms_step * 1e-3)
        n = n_owned
        # algorithmic bytes: B_alg = N*(V+8) + 4*P per build, V = 32 (double4)
        b_alg_build = n * (32 + 8) + 4 * entries_local
        # Dominant kernel = the longest stage of the profiled build.  Algorithmic bytes per launch:
        #   emit_kernel      writes 4 B per list entry, reads 8 B offset + 4 B id per row      -> 4 P + 12 N
        #   pairmask_kernel  reads one 16-B cell-sorted record per particle, writes nothing the algorithm asks for;
        #                    it is bounded by FP32 issue (3 FFMA + 1 FADD per distance test), reported beside it
        kern = {"emit": ("emit_kernel", 4 * entries_local + 12 * n),
                "emit3": ("emit3_kernel", 4 * entries_local + 12 * n),
                "emit_run": ("emitrun_kernel", 4 * entries_local + 12 * n),
                "runmask": ("runmask_kernel", 16 * n),
                "pairmask": ("pairmask_kernel", 16 * n),
                "rowmask": ("rowmask4_kernel", 16 * n),
                "search_fill": ("search_kernel<FILL>", n * (16 + 8) + 4 * entries_local),
                "search_count": ("search_kernel<COUNT>", n * (16 + 4))}
        dom = max((k for k in stage_ms if k in kern), key=lambda k: stage_ms[k], default=None)
        roof = {"bound": "hbm", "kernel": None, "achieved": None, "peak": peak, "unit": "GB/s",
                "frac": None, "traffic": None, "peak_source": peak_src}
        if dom:
            name, b_alg_k = kern[dom]
            roof.update({"kernel": name, "algorithmic_bytes_per_launch": b_alg_k, "kernel_ms": stage_ms[dom],
                         "achieved": b_alg_k / (stage_ms[dom] * 1e-3) / 1e9})
            roof["frac"] = roof["achieved"] / peak
            # share of the build's KERNEL time (the memset and the status copy are not kernels): comparable with the
            # per-launch durations of the ncu launch list in profiles/
            kernel_stages = [v for k, v in stage_ms.items() if k not in ("zero", "status_copy")]
            roof["kernel_share_of_build"] = stage_ms[dom] / max(sum(kernel_stages), 1e-12)
        # ncu --set full capture of the same kernel (profiles/): dram__bytes_read.sum + dram__bytes_write.sum
        traffic_file = os.path.join(ROOT, "profiles", "dram_traffic.json")
        if dom and os.path.exists(traffic_file):
            with open(traffic_file) as f:
                roof["traffic"] = json.load(f).get(kern[dom][0])
        search = next((k for k in ("runmask", "pairmask", "rowmask") if k in stage_ms), None)
        if search:
            # secondary ceiling: the distance tests themselves, 4 FP32-pipe lane-ops each at the measured
            # lane-FMA rate (profiles/r01_microbench_issue_rates.txt) x SM count x SM clock of this device
            t_fp32 = st.candidates_tested * 4 / _fp32_issue_rate(dev) * 1e3
            roof["fp32_issue_floor_ms_pairmask"] = t_fp32
            roof["pairmask_ms"] = stage_ms[search]
            roof["pairmask_frac_of_fp32_issue"] = t_fp32 / stage_ms[search]
            roof["search_kernel"] = kern[search][0]
        build_gbs = b_alg_build / (ms_step * 1e-3) / 1e9
        q_host_np = q
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            qref = np.ascontiguousarray(q_host_np)
            times, cpairs, kind = cpu_reference_run(qref, args.cpu_loops)
            best = min(times, key=times.get)
            cpu = {"value": cpairs / (times[best] * 1e-3), "unit": UNIT, "cores": 1, "kind": kind,
                   "sample": f"{args.cpu_loops} full builds per variant of the same density-1.0 default system, "
                             f"reference classes compiled from the reference sources (oracle/_ref), 1 thread "
                             f"(the reference is single-threaded); ms/build: "
                             + ", ".join(f"{k}={v:.1f}" for k, v in times.items()),
                   "ms_per_build": times, "host_nproc": os.cpu_count()}
        # kernels of one build: bin (+ cell scan), scatter, cellsort, runmask, scan(counts), emitrun, finalize
        # (a slab rank: the binning kernel runs twice — owned records + halo send, then the ghosts — or, with
        #  NLB_HALO_FUSED=0 / NCCL, one packing kernel precedes the build: one more launch either way)
        kernels_per_build = 7 if world == 1 else 7 + 1
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(world), **({"halo": halo_transport} if halo_transport else {})),
            "clocks": clocks,
            "e2e": {"value": pairs_all / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes},
            "gpu_launches": kernels_per_build * args.steps,
            "roofline": roof,
            "cpu_baseline": cpu,
            "build": {"ms_cold_l2": ms_step, "ms_hot_l2_back_to_back": ms_hot,
                      "entries_per_s": entries_all / (ms_step * 1e-3),
                      "candidate_tests_per_s": st.candidates_tested / (ms_step * 1e-3),
                      "candidates_per_pass": st.candidates_tested, "band_retests": st.band_tests,
                      "algorithmic_bytes": b_alg_build, "algorithmic_gbs": build_gbs,
                      "frac_of_hbm_peak": build_gbs / peak, "stage_ms": stage_ms,
                      "ms_per_step_each": ms_steps if len(ms_steps) <= 32 else ms_steps[:32]},
        }
        line.update(extras)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    faulthandler.enable()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-loops", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="headline measurement only: skip the side workloads, the c2 / c3 blocks and the reference GPU run")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
