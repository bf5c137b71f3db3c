"""Times one workload (not the headline bench): python tools/bench_workload.py KIND N [mode] [reps]
KIND: uniform | fcc | clustered (fcc: N is the box edge; density from the environment variable DENSITY, default 1.0).
Prints ms/build (graph replay, L2-cold), stage times and throughput."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from md_neighbor_list_b200 import VerletListB200, workloads  # noqa: E402

kind = sys.argv[1]
n = int(sys.argv[2])
mode = sys.argv[3] if len(sys.argv) > 3 else "full_csr"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
SL = 3.3
if kind == "uniform":
    L = float(round(n ** (1.0 / 3.0)))
    q = workloads.uniform(n, L)
elif kind == "fcc":
    L = float(n)  # here N is the box edge
    q = workloads.fcc(float(os.environ.get("DENSITY", "1.0")), L)
else:
    L = float(round(n ** (1.0 / 3.0)))
    q = workloads.clustered(n, L)
n = q.shape[0]
qd = torch.from_numpy(q).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = {"kind": kind, "n": n, "L": L, "mode": mode}
for profile in (False, True):
    nl = VerletListB200(SL, L, L, L, mode=mode, profile=profile, kernel_variant=int(os.environ.get("NLB_VARIANT", "0")))
    nl.initialize(n)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        nl.build(qd)
    for _ in range(4):
        try:
            st = nl.synchronize()
            break
        except Exception as e:
            if getattr(e, "status", 0) == 7:
                nl.reserve_cell_capacity(nl.stats().max_in_cell)
            else:
                nl.reserve(nl.stats().required_entries)
            with torch.cuda.stream(s):
                nl.build(qd)
    ms = []
    stages = {}
    for r in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s):
            flush.fill_(1)
            e0.record(s)
            nl.build(qd)
            e1.record(s)
        st = nl.synchronize()
        ms.append(e0.elapsed_time(e1))
        if profile:
            for k, v in nl.stage_times().items():
                stages[k] = stages.get(k, 0.0) + v / reps
    if not profile:
        out["ms_per_build"] = sorted(ms)[len(ms) // 2]
        out["entries"] = st.number_of_pairs
        out["candidates_per_pass"] = st.candidates_tested
        out["band"] = st.band_tests
        out["max_partners"] = st.max_partners
        out["max_in_cell"] = st.max_in_cell
        out["G_tests_per_s"] = st.candidates_tested / (out["ms_per_build"] * 1e-3) / 1e9
        out["ns_per_particle"] = out["ms_per_build"] * 1e6 / n
        b_alg = n * 40 + 4 * st.number_of_pairs
        out["alg_GBs"] = b_alg / (out["ms_per_build"] * 1e-3) / 1e9
    else:
        out["stage_ms"] = {k: round(v, 4) for k, v in stages.items()}
    nl.close()
    torch.cuda.empty_cache()
print(json.dumps(out))
