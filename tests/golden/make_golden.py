"""Generates tests/golden/*.json|npz by RUNNING THE REFERENCE's own classes (oracle/_ref/libref_*.so, compiled from
/root/reference by oracle/Makefile) — run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

default_systems.json : fingerprints of the reference's output on its two default systems (make_list.cpp:17-24,
                       density 1.0 and 0.5), FNV-1a-64 over the raw little-endian bytes, plus the known-answer
                       counts of SURVEY.md §8 / BASELINE.md §4.
small_mesh3.npz      : complete reference output (half list) for a 3x3x3-cell box, plus the full list.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    O.build(with_ref=True)
    out = {}
    for dens in (1.0, 0.5):
        q = O.gen_fcc(dens)
        entry = {"n": int(q.shape[0]), "positions_xyz_fnv": O.fnv1a64(q[:, :3]),
                 "q0": [float(v) for v in q[0, :3]]}
        per_variant = {}
        for v in ("scalar", "scalar_swp", "avx2_4x1", "avx512_8x1"):
            if not O.ref_available(v):
                continue
            r, _ = O.ref_build(v, q, 3.3, (50.0, 50.0, 50.0))
            rs = r.sorted_rows()
            per_variant[v] = {
                "number_of_pairs": rs.number_of_pairs,
                "number_of_partners_fnv": O.fnv1a64(rs.number_of_partners),
                "key_pointer_i32_fnv": O.fnv1a64(rs.offsets.astype(np.int32)),
                "sorted_list_rowsorted_fnv": O.fnv1a64(rs.partners),
                "np_first8": [int(x) for x in rs.number_of_partners[:8]],
                "max_partners": int(rs.number_of_partners.max()),
            }
        vals = list(per_variant.values())
        assert all(v == vals[0] for v in vals), "reference variants disagree"
        entry["half"] = vals[0]
        entry["reference_variants_run"] = sorted(per_variant)
        full = O.build_full(q, 3.3, (50.0, 50.0, 50.0))
        bf_ok = None
        if dens == 0.5:  # the drivers' brute force (make_list.cu:79-98) — 62 500^2 tests, ~10 s
            bf = O.bruteforce(q, 3.3, full=True)
            fs = full.sorted_rows()
            bf_ok = bool(np.array_equal(bf.partners, fs.partners) and np.array_equal(bf.offsets, fs.offsets))
            assert bf_ok
        fs = full.sorted_rows()
        entry["full"] = {"number_of_pairs": fs.number_of_pairs,
                         "number_of_partners_fnv": O.fnv1a64(fs.number_of_partners),
                         "offsets_i64_fnv": O.fnv1a64(fs.offsets),
                         "list_rowsorted_fnv": O.fnv1a64(fs.partners),
                         "max_partners": int(fs.number_of_partners.max()),
                         "candidates_27": int(full.candidates),
                         "bruteforce_checked": bf_ok}
        entry["half"]["candidates_13"] = int(O.build_half(q, 3.3, (50.0, 50.0, 50.0)).candidates)
        out[f"density_{dens}"] = entry
    with open(os.path.join(HERE, "default_systems.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)

    # small complete fixture: 3 cells per axis (every cell is a neighbour of every other cell)
    L, SL = 10.5, 3.3
    q = O.gen_fcc(1.0, L)
    r, _ = O.ref_build("scalar", q, SL, (L, L, L))
    rs = r.sorted_rows()
    bf = O.bruteforce(q, SL, full=False)
    assert np.array_equal(bf.partners, rs.partners) and np.array_equal(bf.offsets, rs.offsets)
    full = O.bruteforce(q, SL, full=True)
    np.savez_compressed(os.path.join(HERE, "small_mesh3.npz"), q=q, L=L, SL=SL,
                        half_np=rs.number_of_partners, half_off=rs.offsets, half_list=rs.partners,
                        full_np=full.number_of_partners, full_off=full.offsets, full_list=full.partners)
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
