// make_list_b200.cpp — a driver of the reference's shape (make_list.cu:102-201 for the GPU class, make_list.cpp:132-226
// for the CPU classes) built on include/nlist_b200_shim.hpp: generate the jittered-FCC default system, build the list
// LOOP times, print "# of particles N T[ms]", then verify against an O(N^2) brute force and print "TEST is passed."
// usage: make_list_b200.out [gpu|cpu|md|pbc|slab] [density] [loop] [check]
//   gpu : NeighListGPU interface (full list, list[k*N + i] layout)     cpu : NeighList interface (half list, CSR)
//   md  : what the drivers' unused momenta `p` are for (make_list.cpp:135-140): LOOP velocity-Verlet steps of a
//         Lennard-Jones system in a 20^3 box on the NeighListGPU interface — forces from the list on the device, the
//         list rebuilt only when a particle has moved more than margin / 2 since the last build
//         — then the forces of the (possibly several steps old) list are checked against an O(N^2) evaluation
//   pbc : periodic boundaries (minimum image) on the NeighListPeriodicGPU shim class: FULL and HALF
//         lists of a 20^3 box checked against an O(N^2) minimum-image brute force
//   slab G : the multi-GPU build from a C++ host, C ABI only: G processes (one per GPU; on a box with
//         fewer GPUs they share devices), each owns one L^3 block of a box L x L x G L stacked along z.  The 64-byte
//         CUDA IPC handles of the assembly buffers travel once over pipes (an MPI host would use MPI_Allgather); after
//         that a step is ONE call per rank, nlb200_build_subset: its binning kernel sends the face particles into the
//         neighbours' buffers over NVLink, flags on the device order the steps (nlb200_set_halo_sync / _pack).  Every
//         rank checks its rows against an O(N^2) brute force over the GLOBAL system.
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

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

// ---- slab mode: G ranks, halo exchange by peer stores, everything through the C ABI -----------------------------------
struct Pipes {
  int up[2];    // child -> parent
  int down[2];  // parent -> child
};
bool write_all(int fd, const void* p, size_t n) {
  const char* c = static_cast<const char*>(p);
  while (n > 0) {
    const ssize_t w = ::write(fd, c, n);
    if (w <= 0) return false;
    c += w;
    n -= (size_t)w;
  }
  return true;
}
bool read_all(int fd, void* p, size_t n) {
  char* c = static_cast<char*>(p);
  while (n > 0) {
    const ssize_t r = ::read(fd, c, n);
    if (r <= 0) return false;
    c += r;
    n -= (size_t)r;
  }
  return true;
}

#define SLAB_CK(call)                                                                          \
  do {                                                                                         \
    const int st_ = (call);                                                                    \
    if (st_ != NLB200_OK) {                                                                    \
      std::fprintf(stderr, "rank %d: %s -> %s\n", rank, #call, nlb200_status_string(st_));     \
      return 1;                                                                                \
    }                                                                                          \
  } while (0)

int slab_rank(int rank, int world, double density, int loop, const Pipes& io) {
  const double Ls = 20.0;  // block edge: small enough for the O(N^2) check
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
    std::fprintf(stderr, "rank %d: no CUDA device\n", rank);
    return 1;
  }
  cudaSetDevice(rank % ndev);
  // every rank generates EVERY block (the check needs the global system); block r is shifted to z in [r Ls, (r+1) Ls)
  const int64_t n = nlb200_workload_fcc(density, Ls, 0, 0, 0, 2, nullptr, 4, 0);
  std::vector<double4v> global((size_t)n * world);
  for (int r = 0; r < world; r++) {
    nlb200_workload_fcc(density, Ls, 0, 0, 0, 2 + (uint32_t)r, &global[(size_t)n * r].x, 4, n);
    for (int64_t i = 0; i < n; i++) global[(size_t)n * r + i].z += Ls * r;
  }
  // layout of a rank's assembly buffer (same on every rank): [control 256 B | records n_cap + 2 cap | global ids]
  const int64_t cap = (((int64_t)((double)n * (SEARCH_LENGTH / Ls) * 1.5) + 1024 + 31) / 32) * 32;
  const int64_t n_cap = (n + 31) / 32 * 32, n_total = n_cap + 2 * cap;
  const int64_t o_q = 256, o_g = (o_q + n_total * 32 + 255) / 256 * 256, nbytes = o_g + n_total * 4;
  void* base = nullptr;
  char handle[64];
  SLAB_CK(nlb200_p2p_alloc(nbytes, &base, handle));
  // all-gather of the handles through the parent
  std::vector<char> all((size_t)64 * world);
  if (!write_all(io.up[1], handle, 64) || !read_all(io.down[0], all.data(), all.size())) return 1;
  char* peer[2] = {nullptr, nullptr};  // lower, upper neighbour's buffer
  for (int f = 0; f < 2; f++) {
    const int pr = f == 0 ? rank - 1 : rank + 1;
    if (pr < 0 || pr >= world) continue;
    void* pp = nullptr;
    SLAB_CK(nlb200_p2p_open(&all[(size_t)64 * pr], &pp));
    peer[f] = static_cast<char*>(pp);
  }
  char* b = static_cast<char*>(base);
  double* q_all = reinterpret_cast<double*>(b + o_q);
  int32_t* g_all = reinterpret_cast<int32_t*>(b + o_g);
  {
    // owned records in place, every other slot an absent (NaN) record, control block zeroed
    std::vector<double> hq((size_t)n_total * 4, std::nan(""));
    std::memcpy(hq.data(), &global[(size_t)n * rank].x, sizeof(double) * 4 * (size_t)n);
    std::vector<int32_t> hg((size_t)n_total, 0);
    for (int64_t i = 0; i < n; i++) hg[(size_t)i] = (int32_t)(n * rank + i);
    cudaMemset(base, 0, 256);
    cudaMemcpy(q_all, hq.data(), sizeof(double) * hq.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(g_all, hg.data(), sizeof(int32_t) * hg.size(), cudaMemcpyHostToDevice);
  }
  void* state = nullptr;  // cursors, ticket, previous counts; + the two counts of the step
  cudaMalloc(&state, 128);
  cudaMemset(state, 0, 128);
  int64_t* counts = reinterpret_cast<int64_t*>(static_cast<char*>(state) + 64);
  cudaDeviceSynchronize();
  // barrier: nobody writes into a neighbour before it has initialised its buffer
  char tok = 1;
  if (!write_all(io.up[1], &tok, 1) || !read_all(io.down[0], &tok, 1)) return 1;

  // the rank's window of the global grid: its slab plus one search length (+ one cell of slack) on either side
  const double Lz = Ls * world;
  const int m = (int)(Lz / SEARCH_LENGTH);
  const double ms = Lz / m, lo = Ls * rank, hi = Ls * (rank + 1);
  int c_lo = rank == 0 ? 0 : (int)((lo - SEARCH_LENGTH) / ms) - 1;
  int c_hi = rank + 1 == world ? m - 1 : (int)((hi + SEARCH_LENGTH) / ms) + 1;
  c_lo = std::max(c_lo, 0);
  c_hi = std::min(c_hi, m - 1);
  if (c_hi - c_lo + 1 == 3 && m > 3) {
    if (c_hi + 1 < m) c_hi++; else c_lo--;
  }
  nlb200_handle h = nullptr;
  SLAB_CK(nlb200_create(SEARCH_LENGTH, Ls, Ls, Lz, NLB200_F64, NLB200_FULL_CSR, &h));
  if (world > 1) SLAB_CK(nlb200_set_cell_window(h, 2, c_lo, c_hi - c_lo + 1));
  SLAB_CK(nlb200_initialize(h, n_total, (int64_t)((double)n * 4.18879 * 35.937 * density * 1.4) + 4096));
  auto ctrl_of = [&](char* buf, int word) { return buf ? buf + 8 * word : nullptr; };  // HaloCtrl: step, ready[2], free_from[2]
  if (world > 1) {
    // my lower neighbour sees me as its UPPER face (ready[1], free_from[1] = words 2, 4); my upper one as its lower
    SLAB_CK(nlb200_set_halo_sync(h, base, ctrl_of(peer[0], 4), ctrl_of(peer[1], 3)));
    SLAB_CK(nlb200_set_halo_pack(
        h, 2, peer[0] ? lo + SEARCH_LENGTH : -INFINITY, peer[1] ? hi - SEARCH_LENGTH : INFINITY,
        peer[0] ? peer[0] + o_q + (n_cap + cap) * 32 : nullptr,
        peer[0] ? reinterpret_cast<int32_t*>(peer[0] + o_g + (n_cap + cap) * 4) : nullptr,
        peer[1] ? peer[1] + o_q + n_cap * 32 : nullptr,
        peer[1] ? reinterpret_cast<int32_t*>(peer[1] + o_g + n_cap * 4) : nullptr, cap, counts, state,
        ctrl_of(peer[0], 2), ctrl_of(peer[1], 1), nullptr, nullptr));
  }
  const auto t0 = std::chrono::steady_clock::now();
  for (int k = 0; k < loop; k++) {
    if (world > 1)
      SLAB_CK(nlb200_build_subset(h, q_all, n_total, n, g_all, nullptr));
    else
      SLAB_CK(nlb200_build(h, q_all, n, nullptr));
  }
  SLAB_CK(nlb200_synchronize(h));
  const double ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  // rows of the owned particles (global ids) against the brute force over the global system
  std::vector<int32_t> np((size_t)n);
  std::vector<int64_t> off((size_t)n + 1);
  cudaMemcpy(np.data(), nlb200_number_of_partners(h), sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost);
  cudaMemcpy(off.data(), nlb200_offsets(h), sizeof(int64_t) * ((size_t)n + 1), cudaMemcpyDeviceToHost);
  std::vector<int32_t> list((size_t)off[(size_t)n]);
  cudaMemcpy(list.data(), nlb200_partners(h), sizeof(int32_t) * list.size(), cudaMemcpyDeviceToHost);
  const double sl2 = SEARCH_LENGTH * SEARCH_LENGTH;
  const int64_t ng = n * world;
  std::vector<int32_t> want, got;
  for (int64_t i = 0; i < n; i++) {
    const double4v& a = global[(size_t)(n * rank + i)];
    want.clear();
    for (int64_t j = 0; j < ng; j++) {
      if (j == n * rank + i) continue;
      const double dx = global[(size_t)j].x - a.x, dy = global[(size_t)j].y - a.y, dz = global[(size_t)j].z - a.z;
      if (dx * dx + dy * dy + dz * dz > sl2) continue;
      want.push_back((int32_t)j);
    }
    got.assign(list.begin() + off[(size_t)i], list.begin() + off[(size_t)i + 1]);
    std::sort(got.begin(), got.end());
    if (np[(size_t)i] != (int32_t)want.size() || got != want) {
      std::fprintf(stderr, "rank %d: TEST fail slab_row %lld %zu %zu\n", rank, (long long)i, got.size(), want.size());
      return 1;
    }
  }
  std::printf("# rank %d of %d: %lld owned particles, %lld entries, %d builds in %.1f ms\n", rank, world, (long long)n,
              (long long)off[(size_t)n], loop, ms_total);
  std::fflush(stdout);
  // nobody frees memory a neighbour may still write into
  if (!write_all(io.up[1], &tok, 1) || !read_all(io.down[0], &tok, 1)) return 1;
  nlb200_destroy(h);
  for (int f = 0; f < 2; f++)
    if (peer[f]) nlb200_p2p_close(peer[f]);
  cudaFree(state);
  nlb200_p2p_free(base);
  return 0;
}

int run_slab(int world, double density, int loop) {
  if (world < 1 || world > 16) return fail("slab_world", world, 16);
  // fork BEFORE anything touches CUDA: every rank creates its own context
  std::vector<Pipes> io((size_t)world);
  std::vector<pid_t> pid((size_t)world);
  for (int r = 0; r < world; r++)
    if (pipe(io[(size_t)r].up) != 0 || pipe(io[(size_t)r].down) != 0) return fail("pipe", r, 0);
  for (int r = 0; r < world; r++) {
    pid[(size_t)r] = fork();
    if (pid[(size_t)r] == 0) _exit(slab_rank(r, world, density, loop, io[(size_t)r]));
    if (pid[(size_t)r] < 0) return fail("fork", r, 0);
  }
  bool ok = true;
  // all-gather of the IPC handles, then two barriers (buffers initialised; builds checked)
  std::vector<char> all((size_t)64 * world);
  for (int r = 0; r < world && ok; r++) ok = read_all(io[(size_t)r].up[0], &all[(size_t)64 * r], 64);
  for (int r = 0; r < world && ok; r++) ok = write_all(io[(size_t)r].down[1], all.data(), all.size());
  for (int round = 0; round < 2 && ok; round++) {
    char tok = 0;
    for (int r = 0; r < world && ok; r++) ok = read_all(io[(size_t)r].up[0], &tok, 1);
    for (int r = 0; r < world && ok; r++) ok = write_all(io[(size_t)r].down[1], &tok, 1);
  }
  if (!ok)  // a rank died: release the others from their pipe reads
    for (int r = 0; r < world; r++) close(io[(size_t)r].down[1]);
  for (int r = 0; r < world; r++) {
    int st = 0;
    waitpid(pid[(size_t)r], &st, 0);
    if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) ok = false;
  }
  if (!ok) return fail("slab", world, 0);
  std::fprintf(stderr, "TEST is passed.\n");
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  if (argc > 1 && std::strcmp(argv[1], "slab") == 0)
    return run_slab(argc > 2 ? std::atoi(argv[2]) : 2, argc > 3 ? std::atof(argv[3]) : 1.0, argc > 4 ? std::atoi(argv[4]) : 10);
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
