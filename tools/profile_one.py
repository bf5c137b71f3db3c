"""Runs a handful of builds of the reference default system (density 1.0) — the short command ncu wraps.
usage: python tools/profile_one.py [builds] [mode] [density]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from md_neighbor_list_b200 import VerletListB200, workloads  # noqa: E402

builds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
mode = sys.argv[2] if len(sys.argv) > 2 else "full_csr"
dens = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
q = workloads.fcc(dens)
qd = torch.from_numpy(q).cuda()
nl = VerletListB200(3.3, 50.0, 50.0, 50.0, mode=mode, use_graph=False)
nl.initialize(q.shape[0])
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(builds):
        nl.build(qd)
st = nl.synchronize()
print("pairs", st.number_of_pairs, "candidates", st.candidates_tested, "band", st.band_tests)
