"""Bring-up check of the row-mask path against the oracle (small boxes, every mode): python tools/v3_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from md_neighbor_list_b200 import VerletListB200, workloads  # noqa: E402
from oracle import oracle as O  # noqa: E402


def run(q, sl, box, mode, **kw):
    nl = VerletListB200(sl, *box, mode=mode, **kw)
    nl.initialize(max(q.shape[0], 1))
    qd = torch.from_numpy(np.ascontiguousarray(q)).cuda()
    for _ in range(4):
        nl.build(qd)
        try:
            st = nl.synchronize()
            break
        except Exception as e:
            print("  retry:", e)
            if getattr(e, "status", 0) == 7:
                nl.reserve_cell_capacity(nl.stats().max_in_cell)
            else:
                nl.reserve(nl.stats().required_entries)
    off = nl.offsets().cpu().numpy().copy()
    lst = nl.partners().cpu().numpy().copy()
    cnt = nl.number_of_partners().cpu().numpy().copy()
    nl.close()
    return st, cnt, off, lst


def check(name, q, sl, box, variant=0):
    for mode, builder in (("full_csr", O.build_full), ("half_csr", O.build_half)):
        ref = builder(q, sl, box)
        st, cnt, off, lst = run(q, sl, box, mode, kernel_variant=variant)
        ok_c = np.array_equal(cnt, ref.number_of_partners)
        ok_o = np.array_equal(off, ref.offsets)
        ok_l = ok_o and np.array_equal(lst, ref.partners)
        ok_s = ok_l
        if ok_o and not ok_l:
            a = lst.copy()
            O.lib().orc_sort_rows(a.ctypes.data, off.ctypes.data, q.shape[0])
            ok_s = np.array_equal(a, ref.sorted_rows().partners)
        print(f"{name:28s} {mode:9s} v{variant} pairs {st.number_of_pairs:>10d} ref {ref.number_of_pairs:>10d} counts "
              f"{'OK' if ok_c else 'FAIL'} offsets {'OK' if ok_o else 'FAIL'} order {'OK' if ok_l else 'no'} "
              f"sets {'OK' if ok_s else 'FAIL'} cand {st.candidates_tested} band {st.band_tests}", flush=True)
        if not ok_c:
            bad = np.nonzero(cnt != ref.number_of_partners)[0]
            print("   first bad rows", bad[:10], cnt[bad[:10]], ref.number_of_partners[bad[:10]])


v = int(os.environ.get("NLB_VARIANT", "0"))
L = 14.0
check("fcc 14", workloads.fcc(1.0, L), 3.3, (L, L, L), v)
L = 24.0
check("fcc 24", workloads.fcc(1.0, L), 3.3, (L, L, L), v)
check("fcc 50 d0.5", workloads.fcc(0.5), 3.3, (50.0,) * 3, v)
check("fcc 50 d1.0", workloads.fcc(1.0), 3.3, (50.0,) * 3, v)
rng = np.random.default_rng(8)
box = (13.0, 29.5, 10.1)
q = np.zeros((5000, 4))
q[:, :3] = rng.random((5000, 3)) * np.array(box)
check("ragged 3/8/3", q, 3.3, box, v)
q = np.zeros((300, 4))
q[:, :3] = rng.random((300, 3)) * 60.0
check("sparse", q, 3.3, (60.0,) * 3, v)
check("clustered 12000", workloads.clustered(12000, 30.0, blobs=4), 2.3, (30.0,) * 3, v)
check("clustered 3000", workloads.clustered(3000, 18.0, blobs=2), 2.3, (18.0,) * 3, v)
print("v3_check done")
