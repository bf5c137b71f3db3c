"""Small builds of every mode and kernel variant (a quick crash check; compute-sanitizer is not available on the GPU pool): python tools/sanitize_small.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from md_neighbor_list_b200 import PeriodicVerletList, VerletListB200, workloads  # noqa: E402

L = 14.0
q = workloads.fcc(1.0, L)
qd = torch.from_numpy(q).cuda()
for mode in ("full_csr", "half_csr", "full_ell_transposed"):
    for variant in (0, 1, 3, 4):
        nl = VerletListB200(3.3, L, L, L, mode=mode, kernel_variant=variant, use_graph=False)
        nl.initialize(q.shape[0])
        nl.build(qd)
        st = nl.synchronize()
        print(mode, variant, st.number_of_pairs)
        nl.close()
qc = workloads.clustered(3000, 18.0, blobs=2)
nl = VerletListB200(2.3, 18.0, 18.0, 18.0, mode="full_csr", use_graph=False, max_in_cell=2048)
nl.initialize(qc.shape[0], 30_000_000)
nl.build(torch.from_numpy(qc).cuda())
print("clustered", nl.synchronize().number_of_pairs, nl.stats().max_in_cell)
nl.close()
own = np.nonzero(q[:, 2] < 7.0)[0]
gh = np.nonzero(q[:, 2] >= 7.0)[0]
qa = np.full((len(own) + len(gh) + 64, 4), np.nan)
qa[:len(own)] = q[own]
qa[len(own):len(own) + len(gh)] = q[gh]
ga = np.zeros(len(qa), dtype=np.int32)
ga[:len(own)] = own
ga[len(own):len(own) + len(gh)] = gh
nl = VerletListB200(3.3, L, L, L, mode="half_csr", use_graph=False)
nl.initialize(len(qa))
nl.build(torch.from_numpy(qa).cuda(), n_owned=len(own), global_ids=torch.from_numpy(ga).cuda())
print("subset", nl.synchronize().number_of_pairs)
nl.close()
pl = PeriodicVerletList(3.3, L, L, L, use_graph=False)
pl.initialize(q.shape[0])
pl.build(qd)
print("periodic", pl.synchronize().number_of_pairs)
