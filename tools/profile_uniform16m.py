"""Two builds of 2^24 uniform particles (the command ncu wraps for the configs[2] traffic figures)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from md_neighbor_list_b200 import VerletListB200, workloads  # noqa: E402

n = 1 << 24
L = 256.0
qd = torch.from_numpy(workloads.uniform(n, L)).cuda()
nl = VerletListB200(3.3, L, L, L, mode="full_csr", use_graph=False, kernel_variant=int(os.environ.get("NLB_VARIANT", "0")))
nl.initialize(n)
for _ in range(2):
    nl.build(qd)
    try:
        st = nl.synchronize()
    except Exception as e:  # capacity the estimate missed
        if getattr(e, "status", 0) == 7:
            nl.reserve_cell_capacity(nl.stats().max_in_cell)
        else:
            nl.reserve(nl.stats().required_entries)
print("entries", nl.synchronize().number_of_pairs)
