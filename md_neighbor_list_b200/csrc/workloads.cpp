// workloads.cpp — host-side input generators of the drivers (not part of the list build).
//
// The reference's workload is produced by its driver, not by the list-builder class: make_list.cpp:34-77 /
// make_list.cu:26-66 (`add_particle` + `init`).  A drop-in driver needs the same particles, so the generator is
// provided here with the standard library's own engines/distributions — the identical std::mt19937 /
// std::uniform_real_distribution calls the reference makes — plus the synthetic distributions of SURVEY.md §8d
// (C2 uniform, C4 clustered).  oracle/nlist_oracle.c restates the same generators independently in C; the tests
// compare the two bit for bit.
#include <cmath>
#include <cstdint>
#include <random>
#include <vector>

#include "../../include/nlist_b200.h"

extern "C" {

int64_t nlb200_workload_fcc(double density, double L, int sx, int sy, int sz, uint32_t seed, double* q, int stride,
                            int64_t capacity) {
  if (!(density > 0) || !(L > 0) || (stride != 3 && stride != 4)) return -1;
  const double s = 1.0 / std::pow(density * 0.25, 1.0 / 3.0);  // make_list.cpp:54
  const double hs = s * 0.5;
  if (sx <= 0) sx = static_cast<int>(L / s);
  if (sy <= 0) sy = static_cast<int>(L / s);
  if (sz <= 0) sz = static_cast<int>(L / s);
  const int64_t total = 4ll * sx * sy * sz;
  if (q == nullptr) return total;
  if (total > capacity) return -1;
  std::mt19937 mt(seed);  // make_list.cpp:40 (`static std::mt19937 mt(2)`)
  int64_t n = 0;
  auto add = [&](double x, double y, double z) {  // make_list.cpp:34-49
    std::uniform_real_distribution<double> ud(0.0, 0.1);
    double* p = q + n * stride;
    p[0] = x + ud(mt);
    p[1] = y + ud(mt);
    p[2] = z + ud(mt);
    if (stride == 4) p[3] = 0.0;
    n++;
  };
  for (int iz = 0; iz < sz; iz++)
    for (int iy = 0; iy < sy; iy++)
      for (int ix = 0; ix < sx; ix++) {
        const double x = ix * s, y = iy * s, z = iz * s;
        add(x, y, z);  // make_list.cpp:65-68
        add(x, y + hs, z + hs);
        add(x + hs, y, z + hs);
        add(x + hs, y + hs, z);
      }
  return n;
}

int64_t nlb200_workload_uniform(int64_t n, double L, uint64_t seed, double* q, int stride) {
  if (n < 0 || !(L > 0) || (stride != 3 && stride != 4) || q == nullptr) return -1;
  std::mt19937_64 mt(seed);
  std::uniform_real_distribution<double> ud(0.0, L);
  for (int64_t i = 0; i < n; i++) {
    double* p = q + i * stride;
    p[0] = ud(mt);
    p[1] = ud(mt);
    p[2] = ud(mt);
    if (stride == 4) p[3] = 0.0;
  }
  return n;
}

static double reflect_into(double v, double L) {
  // reflect at the walls until inside [0, L)
  for (int it = 0; it < 64; it++) {
    if (v < 0)
      v = -v;
    else if (v >= L)
      v = 2.0 * L - v;
    else
      break;
  }
  if (!(v >= 0 && v < L)) v = 0.5 * L;
  if (v >= L) v = std::nextafter(L, 0.0);
  return v;
}

int64_t nlb200_workload_clustered(int64_t n, double L, int blobs, uint64_t seed, double* q, int stride) {
  if (n < 0 || !(L > 0) || blobs < 1 || (stride != 3 && stride != 4) || q == nullptr) return -1;
  std::mt19937_64 mt(seed);
  std::uniform_real_distribution<double> ud(0.0, L);
  std::normal_distribution<double> nd(0.0, L / 40.0);
  std::vector<double> c(3 * (size_t)blobs);
  for (auto& v : c) v = ud(mt);
  const int64_t nbg = n / 2;
  for (int64_t i = 0; i < n; i++) {
    double* p = q + i * stride;
    if (i < nbg) {
      p[0] = ud(mt);
      p[1] = ud(mt);
      p[2] = ud(mt);
    } else {
      const size_t b = (size_t)((i - nbg) % blobs);
      p[0] = reflect_into(c[3 * b + 0] + nd(mt), L);
      p[1] = reflect_into(c[3 * b + 1] + nd(mt), L);
      p[2] = reflect_into(c[3 * b + 2] + nd(mt), L);
    }
    if (stride == 4) p[3] = 0.0;
  }
  return n;
}
}
