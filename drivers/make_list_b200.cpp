// make_list_b200.cpp — a driver of the reference's shape (make_list.cu:102-201 for the GPU class, make_list.cpp:132-226
// for the CPU classes) built on include/nlist_b200_shim.hpp: generate the jittered-FCC default system, build the list
// LOOP times, print "# of particles N T[ms]", then verify against an O(N^2) brute force and print "TEST is passed."
// usage: make_list_b200.out [gpu|cpu] [density] [loop] [check]
//   gpu : NeighListGPU interface (full list, list[k*N + i] layout)     cpu : NeighList interface (half list, CSR)
#include <algorithm>
#include <chrono>
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

}  // namespace

int main(int argc, char** argv) {
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
