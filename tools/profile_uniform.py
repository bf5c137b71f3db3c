"""ncu target: a few builds of a uniform-random system.  python tools/profile_uniform.py N [builds]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from md_neighbor_list_b200 import VerletListB200, workloads
n = int(sys.argv[1]); builds = int(sys.argv[2]) if len(sys.argv) > 2 else 2
L = float(round(n ** (1.0 / 3.0)))
q = workloads.uniform(n, L)
qd = torch.from_numpy(q).cuda()
nl = VerletListB200(3.3, L, L, L, mode="full_csr", use_graph=False)
nl.initialize(n)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(builds):
        nl.build(qd)
st = nl.synchronize()
print("pairs", st.number_of_pairs, "candidates", st.candidates_tested)
