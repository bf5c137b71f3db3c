"""CPU tests: the oracle (oracle/nlist_oracle.c) against the golden vectors recorded from the reference's own
classes (tests/golden/make_golden.py), the reference brute force, and — where oracle/_ref exists — the reference
classes themselves.  No GPU, no product code."""
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(GOLD, "default_systems.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("dens", [1.0, 0.5])
def test_known_answers_of_default_systems(oracle, golden, dens):
    """SURVEY.md §8 / BASELINE.md §4 known-answer values + fingerprints from the reference run."""
    g = golden[f"density_{dens}"]
    q = oracle.gen_fcc(dens)
    assert q.shape[0] == g["n"] == {1.0: 119164, 0.5: 62500}[dens]
    assert [float(v) for v in q[0, :3]] == g["q0"]
    assert q[0, 0] == 0.018508208157401413
    assert oracle.fnv1a64(q[:, :3]) == g["positions_xyz_fnv"]
    h = oracle.build_half(q, 3.3, (50.0, 50.0, 50.0))
    assert h.number_of_pairs == g["half"]["number_of_pairs"] == {1.0: 7839886, 0.5: 2268138}[dens]
    assert h.candidates == g["half"]["candidates_13"] == {1.0: 57007044, 0.5: 15593750}[dens]
    assert int(h.number_of_partners.max()) == {1.0: 94, 0.5: 50}[dens]
    assert list(h.number_of_partners[:8]) == g["half"]["np_first8"]
    hs = h.sorted_rows()
    assert oracle.fnv1a64(hs.number_of_partners) == g["half"]["number_of_partners_fnv"]
    assert oracle.fnv1a64(hs.offsets.astype(np.int32)) == g["half"]["key_pointer_i32_fnv"]
    assert oracle.fnv1a64(hs.partners) == g["half"]["sorted_list_rowsorted_fnv"]
    f = oracle.build_full(q, 3.3, (50.0, 50.0, 50.0))
    assert f.number_of_pairs == 2 * h.number_of_pairs == g["full"]["number_of_pairs"]
    assert f.candidates == g["full"]["candidates_27"] == {1.0: 114133252, 0.5: 31250000}[dens]
    assert int(f.number_of_partners.max()) == {1.0: 149, 0.5: 78}[dens]
    fs = f.sorted_rows()
    assert oracle.fnv1a64(fs.partners) == g["full"]["list_rowsorted_fnv"]
    assert oracle.fnv1a64(fs.offsets) == g["full"]["offsets_i64_fnv"]


def test_density_one_row0(oracle):
    """row 0 of the density-1.0 half list starts 1 2 3 4 5 6 7 8 124 125 126 127."""
    q = oracle.gen_fcc(1.0)
    h = oracle.build_half(q, 3.3, (50.0, 50.0, 50.0)).sorted_rows()
    assert list(h.partners[:12]) == [1, 2, 3, 4, 5, 6, 7, 8, 124, 125, 126, 127]


def test_small_fixture_and_bruteforce(oracle):
    z = np.load(os.path.join(GOLD, "small_mesh3.npz"))
    q, L, SL = z["q"], float(z["L"]), float(z["SL"])
    h = oracle.build_half(q, SL, (L, L, L)).sorted_rows()
    assert np.array_equal(h.number_of_partners, z["half_np"])
    assert np.array_equal(h.offsets, z["half_off"])
    assert np.array_equal(h.partners, z["half_list"])
    f = oracle.build_full(q, SL, (L, L, L)).sorted_rows()
    assert np.array_equal(f.number_of_partners, z["full_np"])
    assert np.array_equal(f.partners, z["full_list"])
    bf = oracle.bruteforce(q, SL, full=True)
    assert np.array_equal(bf.partners, f.partners) and np.array_equal(bf.offsets, f.offsets)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_cell_list_equals_bruteforce_on_random_inputs(oracle, seed):
    rng = np.random.default_rng(seed)
    L = (14.0, 11.0, 17.5)
    n = 1500
    q = np.zeros((n, 4))
    q[:, :3] = rng.random((n, 3)) * np.array(L)
    for full in (False, True):
        cl = (oracle.build_full if full else oracle.build_half)(q, 2.7, L).sorted_rows()
        bf = oracle.bruteforce(q, 2.7, full=full)
        assert np.array_equal(cl.offsets, bf.offsets)
        assert np.array_equal(cl.partners, bf.partners)
    # float32 restatement too
    qf = q.astype(np.float32)
    cl = oracle.build_full(qf, 2.7, L).sorted_rows()
    bf = oracle.bruteforce(qf, 2.7, full=True)
    assert np.array_equal(cl.partners, bf.partners)


def test_binning_is_a_stable_counting_sort(oracle):
    q = oracle.gen_fcc(0.5)
    mi, pid, cell = oracle.bin_particles(q, 3.3, (50.0, 50.0, 50.0))
    assert mi[0] == 0 and mi[-1] == q.shape[0] and len(mi) == 15 ** 3 + 1
    occ = np.diff(mi)
    assert occ.min() == 13 and occ.max() == 32  # occupancy 13-32 at density 0.5
    for m in (0, 17, 3374):
        seg = pid[mi[m]:mi[m + 1]]
        assert np.all(np.diff(seg) > 0) and np.all(cell[seg] == m)


def test_band_report_default_system(oracle):
    """no pair of the density-0.5 default system depends on the rounding order."""
    q = oracle.gen_fcc(0.5)
    rep = oracle.band_report(q, 3.3, (50.0, 50.0, 50.0))
    assert rep["order_dependent"] == 0 and rep["within_1ulp"] == 0


def test_generators_match_libstdcxx(oracle):
    """oracle's C restatement of mt19937 / mt19937_64 / uniform_real_distribution vs the C++ standard library
    (the product's csrc/workloads.cpp calls the real thing)."""
    from md_neighbor_list_b200 import workloads
    assert np.array_equal(workloads.fcc(0.5), oracle.gen_fcc(0.5))
    assert np.array_equal(workloads.fcc(1.0, 20.0, seed=7, stride=3), oracle.gen_fcc(1.0, 20.0, seed=7, stride=3))
    assert np.array_equal(workloads.uniform(50000, 256.0), oracle.gen_uniform(50000, 256.0))
    c = workloads.clustered(20000, 64.0)
    assert c[:, :3].min() >= 0.0 and c[:, :3].max() < 64.0


@pytest.mark.parametrize("variant", ["scalar", "scalar_swp", "avx2_4x1", "avx512_8x1"])
def test_oracle_equals_reference_classes(oracle, variant):
    """The real thing: the reference's own classes compiled from /root/reference (oracle/_ref)."""
    if not oracle.ref_available(variant):
        pytest.skip(f"oracle/_ref/{oracle.REF_VARIANTS[variant]} not built or CPU lacks the ISA")
    q = oracle.gen_fcc(0.5)  # configs[0]: density 0.5, AVX2 4x1 is the PR1 bit-exact reference
    r, _ = oracle.ref_build(variant, q, 3.3, (50.0, 50.0, 50.0))
    r = r.sorted_rows()
    h = oracle.build_half(q, 3.3, (50.0, 50.0, 50.0)).sorted_rows()
    assert r.number_of_pairs == h.number_of_pairs == 2268138
    assert np.array_equal(r.number_of_partners, h.number_of_partners)
    assert np.array_equal(r.offsets, h.offsets)
    assert np.array_equal(r.partners, h.partners)
