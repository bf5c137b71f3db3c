# quick perf iteration: default system (+ optional parity smoke)
python - <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, '.')
from md_neighbor_list_b200 import VerletListB200, workloads
from oracle import oracle as O
L, SL = 20.0, 3.3
q = workloads.fcc(1.0, L)
qd = torch.from_numpy(q).cuda()
for mode, ref in (("full_csr", O.build_full(q, SL, (L, L, L))), ("half_csr", O.build_half(q, SL, (L, L, L)))):
    nl = VerletListB200(SL, L, L, L, mode=mode); nl.initialize(q.shape[0])
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        nl.build(qd)
    st = nl.synchronize()
    got = nl.partners().cpu().numpy().copy(); off = nl.offsets().cpu().numpy()
    O.lib().orc_sort_rows(got.ctypes.data, off.ctypes.data, q.shape[0])
    refs = ref.sorted_rows()
    ok = st.number_of_pairs == refs.number_of_pairs and np.array_equal(off, refs.offsets) and np.array_equal(got, refs.partners)
    print(mode, "parity", "OK" if ok else "FAIL")
PY
python tools/bench_workload.py fcc 50 full_csr 9 2>&1 | tail -1
