"""Golden digests for BASELINE.json configs[2] (uniform-random particles, density 1.0, SL 3.3) at 2^21 and 2^24.

Run in the build container:   python tests/golden/make_golden_uniform.py
Writes tests/golden/uniform_large.json.  The lists are produced by the oracle (oracle/nlist_oracle.c, pinned against
the reference's own classes by tests/test_oracle.py); at these sizes neither the reference's fixed capacities nor its
O(N^2) self-check can serve.  2^24 needs ~12 GB of host memory and a few minutes on one core, which
is why the GPU test compares FNV-1a-64 digests instead of re-running the oracle on the GPU box.

Digest = FNV-1a-64 over the raw little-endian bytes (oracle.fnv1a64), of
  half: number_of_partners int32[n], offsets int64[n+1], the row-sorted partner list int32[P]
  full: number_of_partners int32[n]  (= half count + in-degree: every pair appears in both rows)
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402

SL = 3.3


def digest(n: int) -> dict:
    L = float(round(n ** (1.0 / 3.0)))
    t0 = time.time()
    q = O.gen_uniform(n, L)
    half = O.build_half(q, SL, (L, L, L)).sorted_rows()
    indeg = np.bincount(half.partners, minlength=n).astype(np.int64)
    full_cnt = (half.number_of_partners.astype(np.int64) + indeg).astype(np.int32)
    out = {
        "n": n, "L": L, "search_length": SL, "seed": 2,
        "positions_xyz_fnv": O.fnv1a64(q[:, :3]),
        "half": {"number_of_pairs": half.number_of_pairs,
                 "number_of_partners_fnv": O.fnv1a64(half.number_of_partners),
                 "offsets_i64_fnv": O.fnv1a64(half.offsets),
                 "list_rowsorted_fnv": O.fnv1a64(half.partners),
                 "max_partners": int(half.number_of_partners.max())},
        "full": {"number_of_pairs": 2 * half.number_of_pairs,
                 "number_of_partners_fnv": O.fnv1a64(full_cnt),
                 "max_partners": int(full_cnt.max())},
        "oracle_seconds": round(time.time() - t0, 1),
    }
    return out


if __name__ == "__main__":
    O.build()
    res = {f"n_{n}": digest(n) for n in (1 << 21, 1 << 24)}
    with open(os.path.join(HERE, "uniform_large.json"), "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)
    print(json.dumps(res, indent=1, sort_keys=True))
