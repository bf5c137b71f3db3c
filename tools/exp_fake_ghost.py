import sys, os, numpy as np, torch
sys.path.insert(0, '/root/repo')
from md_neighbor_list_b200 import VerletListB200, _lib
L = _lib.lib()
s = (0.25) ** (-1.0 / 3.0)
sx, sy, sz = 320, 320, 20
n = L.nlb200_workload_fcc(1.0, 1.0, sx, sy, sz, 2, None, 4, 0)
q = np.zeros((n, 4)); L.nlb200_workload_fcc(1.0, 1.0, sx, sy, sz, 2, q.ctypes.data, 4, n)
qd = torch.from_numpy(q).cuda()
gid = torch.arange(n, dtype=torch.int32, device="cuda")
stream = torch.cuda.Stream()
for label, box_z, kw in (("plain", sz * s, {}), ("n_owned=n-1", sz * s, {"n_owned": n - 1}),
                         ("n_owned=n-1 + gids", sz * s, {"n_owned": n - 1, "global_ids": gid}),
                         ("plain, box 2x in z", 2 * sz * s, {}), ("plain, box 8x in z", 8 * sz * s, {})):
    nl = VerletListB200(3.3, sx * s, sy * s, box_z, profile=True)
    nl.initialize(n, int(n * 150.5 * 1.05))
    for r in range(3):
        with torch.cuda.stream(stream):
            nl.build(qd, stream=stream, **kw)
        nl.synchronize()
    print(label, {k: round(v, 3) for k, v in nl.stage_times().items() if v > 0.05})
    nl.close(); torch.cuda.empty_cache()
