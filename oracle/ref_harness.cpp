// ORACLE — TEST INFRASTRUCTURE ONLY.
// Thin C-ABI harness around the UNMODIFIED reference list-builder classes, compiled from the sources where they
// lie under /root/reference (scalar) or from sed-patched copies under oracle/_ref/patched/ (AVX2 / AVX-512: the
// only edit is `alignas(64)` on shfl_table_, SURVEY.md §8c — the shipped headers read it with aligned loads and
// SIGSEGV under g++).  Built by oracle/Makefile into oracle/_ref/libref_<variant>.so; never committed.
//
// It plays the role of make_list.cpp:132-163 (construct, Initialize, LOOP x MakeNeighList, fetch the accessors)
// and times the loop exactly as make_list.cpp:152-157 does (std::chrono around LOOP builds on identical input).
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <numeric>

#if defined USE_AVX512
#include "neighlist_cpu_avx512.hpp"
#elif defined USE_AVX2
#include "neighlist_cpu_avx2.hpp"
#else
#include "neighlist_cpu.hpp"
#endif

typedef double Dtype;
struct Vec {  // make_list.cpp:26-32
#if defined USE_AVX512 || defined USE_AVX2
  Dtype x, y, z, w;
#else
  Dtype x, y, z;
#endif
};

#if defined USE_AVX512
typedef NeighListAVX512<Vec> RefList;
#elif defined USE_AVX2
typedef NeighListAVX2<Vec> RefList;
#else
typedef NeighList<Vec> RefList;
#endif

extern "C" {

// MAX_PARTNERS = 100 is an *average* capacity (neighlist_cpu.hpp:37,76-78): total pair slots = 100*n.
int64_t ref_pair_capacity(int64_t n) { return 100 * n; }

const char* ref_variant() {
#if defined USE_AVX512
  return "avx512_8x1";
#elif defined USE_AVX2
  return "avx2_4x1";
#elif defined LOOP_FUSION_SWP
  return "scalar_loop_fusion_swp";
#else
  return "scalar_loop_fusion";
#endif
}

// q_xyzw: n x 4 doubles.  Outputs sized n, n+1, list_cap.  Returns 0, or <0 on error.
// `warmup` untimed builds on the SAME instance come first: the reference's pair buffers (3 x 100 x n int32,
// neighlist_cpu.hpp:76-78) are touched for the first time by the first build, and those page faults are not part of
// the steady-state LOOP the reference times (make_list.cpp:152-157: 100 builds on one instance).
int ref_build_warm(const double* q_xyzw, int32_t n, double search_length, double lx, double ly, double lz,
                   int32_t warmup, int32_t loops, int32_t* number_of_partners, int32_t* key_pointer,
                   int32_t* sorted_list, int64_t list_cap, int64_t* number_of_pairs, double* ms_per_build) {
  Vec* q = static_cast<Vec*>(aligned_alloc(64, ((sizeof(Vec) * (size_t)(n + 8) + 63) / 64) * 64));
  if (!q) return -2;
  for (int32_t i = 0; i < n; i++) {
    q[i].x = q_xyzw[4 * i + 0];
    q[i].y = q_xyzw[4 * i + 1];
    q[i].z = q_xyzw[4 * i + 2];
#if defined USE_AVX512 || defined USE_AVX2
    q[i].w = 0.0;
#endif
  }
  {
    RefList nlist(search_length, lx, ly, lz);
    nlist.Initialize(n);
    for (int32_t l = 0; l < warmup; l++) nlist.MakeNeighList(q, n);
    const auto beg = std::chrono::system_clock::now();
    for (int32_t l = 0; l < loops; l++) nlist.MakeNeighList(q, n);
    const auto end = std::chrono::system_clock::now();
    *ms_per_build = std::chrono::duration<double, std::milli>(end - beg).count() / (loops > 0 ? loops : 1);
    const int64_t np = nlist.number_of_pairs();
    *number_of_pairs = np;
    if (np > list_cap) {
      free(q);
      return -3;
    }
    std::memcpy(number_of_partners, nlist.number_of_partners(), sizeof(int32_t) * (size_t)n);
    std::memcpy(key_pointer, nlist.key_pointer(), sizeof(int32_t) * ((size_t)n + 1));
    std::memcpy(sorted_list, nlist.sorted_list(), sizeof(int32_t) * (size_t)np);
  }
  free(q);
  return 0;
}

int ref_build(const double* q_xyzw, int32_t n, double search_length, double lx, double ly, double lz, int32_t loops,
              int32_t* number_of_partners, int32_t* key_pointer, int32_t* sorted_list, int64_t list_cap,
              int64_t* number_of_pairs, double* ms_per_build) {
  return ref_build_warm(q_xyzw, n, search_length, lx, ly, lz, 0, loops, number_of_partners, key_pointer, sorted_list,
                        list_cap, number_of_pairs, ms_per_build);
}
}
