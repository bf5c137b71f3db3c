// make_list_b200.cpp — a driver of the reference's shape (make_list.cu:102-201 for the GPU class, make_list.cpp:132-226
// for the CPU classes) built on include/nlist_b200_shim.hpp: generate the jittered-FCC default system, build the list
// LOOP times, print "# of particles N T[ms]", then verify against an O(N^2) brute force and print "TEST is passed."
// usage: make_list_b200.out [gpu|cpu|md|pbc] [density] [loop] [check]
//   gpu : NeighListGPU interface (full list, list[k*N + i] layout)     cpu : NeighList interface (half list, CSR)
//   md  : what the drivers' unused momenta `p` are for (make_list.cpp:135-140): LOOP velocity-Verlet steps of a
//         Lennard-Jones system in a 20^3 box on the NeighListGPU interface — forces from the list on the device, the
//         list rebuilt only when a particle has moved more than margin / 2 since the last build
//         — then the forces of the (possibly several steps old) list are checked against an O(N^2) evaluation
//   pbc : periodic boundaries (minimum image) on the NeighListPeriodicGPU shim class: FULL and HALF
//         lists of a 20^3 box checked against an O(N^2) minimum-image brute force
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "nlist_b200_shim.hpp"

namespace {

struct double4v {
  double x, y, z, w;
};
const double L = 50.0;              // make_list.cpp:22
const double SEARCH_LENGTH = 3.3;   // make_list.cpp:23 (cutoff 3.0 + margin 0.3)

int fail(const char* tag, long long a, long long b) {
  std::fprintf(stderr, "TEST fail %s %lld %lld\n", tag, a, b);
  return 1;
}

// brute force of the drivers (make_list.cpp:79-99 half, make_list.cu:79-98 full): plain distances, no minimum image,
// accept unless r2 > SL2
void bruteforce(const std::vector<double4v>& q, bool full, std::vector<int32_t>& np, std::vector<int32_t>& kp,
                std::vector<int32_t>& list) {
  const int n = (int)q.size();
  const double sl2 = SEARCH_LENGTH * SEARCH_LENGTH;
  np.assign(n, 0);
  std::vector<std::vector<int32_t>> rows(n);
  for (int i = 0; i < n; i++) {
    const double xi = q[i].x, yi = q[i].y, zi = q[i].z;
    for (int j = i + 1; j < n; j++) {
      const double dx = q[j].x - xi, dy = q[j].y - yi, dz = q[j].z - zi;
      const double r2 = dx * dx + dy * dy + dz * dz;
      if (r2 > sl2) continue;
      rows[i].push_back(j);
      if (full) rows[j].push_back(i);
    }
  }
  kp.assign(n + 1, 0);
  for (int i = 0; i < n; i++) {
    std::sort(rows[i].begin(), rows[i].end());
    np[i] = (int32_t)rows[i].size();
    kp[i + 1] = kp[i] + np[i];
  }
  list.resize(kp[n]);
  for (int i = 0; i < n; i++) std::copy(rows[i].begin(), rows[i].end(), list.begin() + kp[i]);
}

// Lennard-Jones forces by brute force: every pair inside rc, plain Euclidean distance (no minimum image, like the list)
void lj_bruteforce(const std::vector<double4v>& q, double rc, double eps, double sigma, std::vector<double>& f) {
  const int n = (int)q.size();
  f.assign((std::size_t)3 * n, 0.0);
  const double rc2 = rc * rc, s2 = sigma * sigma;
  for (int i = 0; i < n; i++)
    for (int j = i + 1; j < n; j++) {
      const double dx = q[i].x - q[j].x, dy = q[i].y - q[j].y, dz = q[i].z - q[j].z;
      const double r2 = dx * dx + dy * dy + dz * dz;
      if (!(r2 < rc2) || r2 == 0.0) continue;
      const double sr2 = s2 / r2, sr6 = sr2 * sr2 * sr2;
      const double fr = 24.0 * eps * sr6 * (2.0 * sr6 - 1.0) / r2;
      f[3 * i] += fr * dx; f[3 * i + 1] += fr * dy; f[3 * i + 2] += fr * dz;
      f[3 * j] -= fr * dx; f[3 * j + 1] -= fr * dy; f[3 * j + 2] -= fr * dz;
    }
}

int run_md(double density, int steps) {
  const double Lmd = 20.0, RC = 3.0, MARGIN = SEARCH_LENGTH - RC, DT = 0.002;
  const int64_t n64 = nlb200_workload_fcc(density, Lmd, 0, 0, 0, 2, nullptr, 4, 0);
  const int32_t N = (int32_t)n64;
  nlb200::cuda_ptr<double4v> q;
  q.allocate(N);
  nlb200_workload_fcc(density, Lmd, 0, 0, 0, 2, &q[0].x, 4, N);
  std::vector<double> p((std::size_t)3 * N);  // the momenta of make_list.cpp:135-140, finally used
  unsigned long long rng = 2;
  for (auto& v : p) {
    rng = rng * 6364136223846793005ull + 1442695040888963407ull;
    v = ((double)(rng >> 11) / 9007199254740992.0 - 0.5) * 2.0;
  }
  q.host2dev();
  nlb200::cuda_ptr<double> f;
  f.allocate((std::size_t)3 * N);
  nlb200::NeighListGPU<double4v, double> nl(SEARCH_LENGTH, Lmd, Lmd, Lmd);
  nl.Initialize(N);
  nl.MakeNeighList(q, N, true);
  nl.TrackReference(q);
  int builds = 1;
  double max_seen = 0.0;
  nl.LJForces(q, RC, 1.0, 1.0, f);
  f.dev2host();
  for (int s = 0; s < steps; s++) {
    // velocity Verlet, half kick + drift on the host (the demo's point is the list, not the integrator); particles are
    // kept inside the open box by reflection
    for (int i = 0; i < N; i++) {
      double* x = &q[i].x;
      for (int d = 0; d < 3; d++) {
        p[3 * i + d] += 0.5 * DT * f[3 * i + d];
        x[d] += DT * p[3 * i + d];
        if (x[d] < 0.0) { x[d] = -x[d]; p[3 * i + d] = -p[3 * i + d]; }
        if (x[d] > Lmd) { x[d] = 2.0 * Lmd - x[d]; p[3 * i + d] = -p[3 * i + d]; }
      }
    }
    q.host2dev();
    const double disp = nl.MaxDisplacement(q);
    max_seen = std::max(max_seen, disp);
    if (disp > 0.5 * MARGIN) {  // some particle left the safety shell: the list may miss a pair inside rc
      nl.MakeNeighList(q, N, true);
      nl.TrackReference(q);
      builds++;
    }
    nl.LJForces(q, RC, 1.0, 1.0, f);
    f.dev2host();
    for (int i = 0; i < 3 * N; i++) p[i] += 0.5 * DT * f[i];
  }
  std::printf("# of particles %d, %d steps, %d list builds, largest displacement seen %.4f (margin/2 = %.3f)\n", N, steps,
              builds, max_seen, 0.5 * MARGIN);
  // forces from the list in use (built up to steps/builds steps ago) against every pair inside rc
  std::vector<double4v> qh(N);
  for (int i = 0; i < N; i++) qh[i] = q[i];
  std::vector<double> fref;
  lj_bruteforce(qh, RC, 1.0, 1.0, fref);
  double worst = 0.0, scale = 1e-300;
  for (int i = 0; i < 3 * N; i++) {
    worst = std::max(worst, std::fabs(f[i] - fref[i]));
    scale = std::max(scale, std::fabs(fref[i]));
  }
  if (!(worst <= 1e-9 * scale)) {
    std::fprintf(stderr, "TEST fail lj_forces %.3e (scale %.3e)\n", worst, scale);
    return 1;
  }
  if (builds < 2 || builds > steps / 2) return fail("list_builds", builds, steps);
  std::fprintf(stderr, "TEST is passed.\n");
  return 0;
}

int run_pbc(double density) {
  const double Lp = 20.0;
  const int64_t n64 = nlb200_workload_fcc(density, Lp, 0, 0, 0, 2, nullptr, 4, 0);
  const int32_t N = (int32_t)n64;
  nlb200::cuda_ptr<double4v> q;
  q.allocate(N);
  nlb200_workload_fcc(density, Lp, 0, 0, 0, 2, &q[0].x, 4, N);
  q.host2dev();
  const double sl2 = SEARCH_LENGTH * SEARCH_LENGTH;
  // minimum-image brute force, rows ascending
  std::vector<std::vector<int32_t>> rows(N);
  for (int i = 0; i < N; i++)
    for (int j = i + 1; j < N; j++) {
      double d[3] = {q[j].x - q[i].x, q[j].y - q[i].y, q[j].z - q[i].z};
      double r2 = 0.0;
      for (int a = 0; a < 3; a++) {
        d[a] -= Lp * std::nearbyint(d[a] / Lp);
        r2 += d[a] * d[a];
      }
      if (r2 > sl2) continue;
      rows[i].push_back(j);
      rows[j].push_back(i);
    }
  for (int half = 0; half < 2; half++) {
    nlb200::NeighListPeriodicGPU<double4v, double> nl(SEARCH_LENGTH, Lp, Lp, Lp, half != 0);
    nl.Initialize(N);
    nl.MakeNeighList(q, N, true);
    nl.MakeNeighList(q, N, true);  // the second build replays the library's graph
    const int64_t pairs = nl.number_of_pairs64();
    auto& off = nl.offsets();
    auto& list = nl.partners();
    off.dev2host();
    list.dev2host();
    int64_t want_pairs = 0;
    std::vector<int32_t> row;
    for (int i = 0; i < N; i++) {
      std::vector<int32_t> want;
      for (int32_t j : rows[i])
        if (!half || j > i) want.push_back(j);
      std::sort(want.begin(), want.end());
      want_pairs += (int64_t)want.size();
      row.assign(&list[(std::size_t)off[i]], &list[(std::size_t)off[i]] + (off[i + 1] - off[i]));
      std::sort(row.begin(), row.end());
      if (row != want) return fail(half ? "pbc_half_row" : "pbc_full_row", i, (long long)row.size() - (long long)want.size());
    }
    if (pairs != want_pairs) return fail("pbc_pairs", pairs, want_pairs);
    std::printf("# of particles %d periodic %s list: %lld entries\n", N, half ? "half" : "full", (long long)pairs);
  }
  std::fprintf(stderr, "TEST is passed.\n");
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  if (argc > 1 && std::strcmp(argv[1], "pbc") == 0) return run_pbc(argc > 2 ? std::atof(argv[2]) : 1.0);
  if (argc > 1 && std::strcmp(argv[1], "md") == 0)
    return run_md(argc > 2 ? std::atof(argv[2]) : 1.0, argc > 3 ? std::atoi(argv[3]) : 200);
  const bool gpu = argc < 2 || std::strcmp(argv[1], "cpu") != 0;
  const double density = argc > 2 ? std::atof(argv[2]) : 1.0;
  const int LOOP = argc > 3 ? std::atoi(argv[3]) : 100;  // make_list.cpp:21
  const bool check = argc > 4 ? std::atoi(argv[4]) != 0 : true;

  const int64_t n64 = nlb200_workload_fcc(density, L, 0, 0, 0, 2, nullptr, 4, 0);
  if (n64 <= 0 || n64 > 400000) {  // driver buffer cap, make_list.cpp:20,73-76
    std::fprintf(stderr, "particle number is too large.\n");
    return 1;
  }
  const int32_t N = (int32_t)n64;
  std::vector<double4v> q(N);
  nlb200_workload_fcc(density, L, 0, 0, 0, 2, &q[0].x, 4, N);

  std::vector<int32_t> np_ref, kp_ref, list_ref;
  if (gpu) {
    nlb200::cuda_ptr<double4v> qd;
    qd.allocate(N);
    for (int i = 0; i < N; i++) qd[i] = q[i];
    qd.host2dev();
    nlb200::NeighListGPU<double4v, double> nl(SEARCH_LENGTH, L, L, L);
    nl.Initialize(N);
    nl.MakeNeighList(qd, N, true);  // warm-up: sizes the partner list
    const auto beg = std::chrono::system_clock::now();
    for (int i = 0; i < LOOP; i++) nl.MakeNeighList(qd, N, false);
    nl.synchronize();
    const auto end = std::chrono::system_clock::now();
    std::printf("# of particles %d %lld[ms]\n", N,
                (long long)std::chrono::duration_cast<std::chrono::milliseconds>(end - beg).count());
    std::printf("%.4f ms per build\n",
                std::chrono::duration_cast<std::chrono::microseconds>(end - beg).count() * 1e-3 / LOOP);
    if (!check) return 0;
    const int32_t pairs = nl.number_of_pairs();
    auto& list = nl.neigh_list();
    auto& np = nl.number_of_partners();
    list.dev2host();
    np.dev2host();
    bruteforce(q, true, np_ref, kp_ref, list_ref);
    if (pairs != kp_ref[N]) return fail("number_of_pairs", pairs, kp_ref[N]);
    std::vector<int32_t> row;
    for (int i = 0; i < N; i++) {
      if (np[i] != np_ref[i]) return fail("number_of_partners", np[i], np_ref[i]);
      row.resize(np[i]);
      for (int k = 0; k < np[i]; k++) row[k] = list[(std::size_t)N * k + i];  // transposed layout, make_list.cu:180-181
      std::sort(row.begin(), row.end());
      for (int k = 0; k < np[i]; k++)
        if (row[k] != list_ref[kp_ref[i] + k]) return fail("neigh_list", row[k], list_ref[kp_ref[i] + k]);
    }
  } else {
    nlb200::NeighList<double4v> nl(SEARCH_LENGTH, L, L, L);
    nl.Initialize(N);
    nl.MakeNeighList(q.data(), N);  // warm-up
    const auto beg = std::chrono::system_clock::now();
    for (int i = 0; i < LOOP; i++) nl.MakeNeighList(q.data(), N);
    const auto end = std::chrono::system_clock::now();
    std::printf("# of particles %d %lld[ms]\n", N,
                (long long)std::chrono::duration_cast<std::chrono::milliseconds>(end - beg).count());
    if (!check) return 0;
    bruteforce(q, false, np_ref, kp_ref, list_ref);
    if (nl.number_of_pairs() != kp_ref[N]) return fail("number_of_pairs", nl.number_of_pairs(), kp_ref[N]);
    for (int i = 0; i < N; i++)
      if (nl.number_of_partners()[i] != np_ref[i]) return fail("number_of_partners", nl.number_of_partners()[i], np_ref[i]);
    for (int i = 0; i <= N; i++)
      if (nl.key_pointer()[i] != kp_ref[i]) return fail("key_pointer", nl.key_pointer()[i], kp_ref[i]);
    std::vector<int32_t> row;
    for (int i = 0; i < N; i++) {  // rows sorted before comparing, make_list.cpp:120-128,211
      row.assign(nl.sorted_list() + kp_ref[i], nl.sorted_list() + kp_ref[i + 1]);
      std::sort(row.begin(), row.end());
      for (std::size_t k = 0; k < row.size(); k++)
        if (row[k] != list_ref[kp_ref[i] + k]) return fail("sorted_list", row[k], list_ref[kp_ref[i] + k]);
    }
  }
  std::fprintf(stderr, "TEST is passed.\n");
  return 0;
}
