/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).
 *
 * Plain-C restatement of the Verlet-list build of kohnakagawa/md_neighbor_list for parity testing of the
 * B200 library.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  Each function cites the reference file:line it follows (paths relative to /root/reference).
 *
 * Parity is PINNED: tests/test_oracle.py checks this file against (1) the golden fingerprints the survey
 * recorded by running the reference (SURVEY.md §8c / BASELINE.md §4), (2) the reference's own classes compiled
 * from /root/reference into oracle/_ref/ (oracle/Makefile), (3) the reference drivers' brute force.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fPIC -shared).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------------------
 * std::mt19937 / std::mt19937_64 and libstdc++'s uniform_real_distribution<double>, restated so that the
 * workload of make_list.cpp:34-77 (`static std::mt19937 mt(2)`, `uniform_real_distribution<Dtype> ud(0.0, 0.1)`)
 * can be regenerated without C++.  libstdc++ generate_canonical<double,53> draws k = ceil(53/32) = 2 words from
 * the 32-bit engine: sum = w0 + w1*2^32 (accumulated in double), ret = sum / 2^64, clamped below 1.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
  uint32_t mt[624];
  int idx;
} orc_mt32;

void orc_mt32_seed(orc_mt32* s, uint32_t seed) {
  s->mt[0] = seed;
  for (int i = 1; i < 624; i++) s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
  s->idx = 624;
}

uint32_t orc_mt32_next(orc_mt32* s) {
  if (s->idx >= 624) {
    for (int i = 0; i < 624; i++) {
      const uint32_t y = (s->mt[i] & 0x80000000u) | (s->mt[(i + 1) % 624] & 0x7fffffffu);
      s->mt[i] = s->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    s->idx = 0;
  }
  uint32_t y = s->mt[s->idx++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

typedef struct {
  uint64_t mt[312];
  int idx;
} orc_mt64;

void orc_mt64_seed(orc_mt64* s, uint64_t seed) {
  s->mt[0] = seed;
  for (int i = 1; i < 312; i++)
    s->mt[i] = 6364136223846793005ull * (s->mt[i - 1] ^ (s->mt[i - 1] >> 62)) + (uint64_t)i;
  s->idx = 312;
}

uint64_t orc_mt64_next(orc_mt64* s) {
  if (s->idx >= 312) {
    for (int i = 0; i < 312; i++) {
      const uint64_t x = (s->mt[i] & 0xffffffff80000000ull) | (s->mt[(i + 1) % 312] & 0x7fffffffull);
      s->mt[i] = s->mt[(i + 156) % 312] ^ (x >> 1) ^ ((x & 1ull) ? 0xb5026f5aa96619e9ull : 0ull);
    }
    s->idx = 0;
  }
  uint64_t x = s->mt[s->idx++];
  x ^= (x >> 29) & 0x5555555555555555ull;
  x ^= (x << 17) & 0x71d67fffeda60000ull;
  x ^= (x << 37) & 0xfff7eee000000000ull;
  x ^= x >> 43;
  return x;
}

static double orc_canonical32(orc_mt32* s) {
  double sum = 0.0, tmp = 1.0;
  for (int k = 0; k < 2; k++) {
    sum += (double)orc_mt32_next(s) * tmp;
    tmp *= 4294967296.0;
  }
  double r = sum / tmp;
  if (r >= 1.0) r = nextafter(1.0, 0.0);
  return r;
}

static double orc_canonical64(orc_mt64* s) {
  double r = (double)orc_mt64_next(s) / 18446744073709551616.0;
  if (r >= 1.0) r = nextafter(1.0, 0.0);
  return r;
}

/* make_list.cpp:51-77 (init) + 34-49 (add_particle): FCC lattice, constant s = (0.25*density)^(-1/3),
 * sx = int(L/s) lattice cells per axis (sx,sy,sz overridable for the multi-GPU config C3), 4 atoms per lattice cell
 * in the order of lines 65-68, each coordinate + U[0,0.1) drawn x,y,z from mt19937(seed).
 * Writes stride reals per particle (w = 0 when stride == 4).  Returns the particle count (or the required count
 * if q == NULL). */
int64_t orc_gen_fcc(double density, double L, int sx, int sy, int sz, uint32_t seed, double* q, int stride,
                    int64_t cap) {
  const double s = 1.0 / pow(density * 0.25, 1.0 / 3.0);
  const double hs = s * 0.5;
  if (sx <= 0) sx = (int)(L / s);
  if (sy <= 0) sy = (int)(L / s);
  if (sz <= 0) sz = (int)(L / s);
  const int64_t total = 4ll * sx * sy * sz;
  if (!q) return total;
  if (total > cap) return -1;
  orc_mt32 mt;
  orc_mt32_seed(&mt, seed);
  int64_t n = 0;
  static const int off[4][3] = {{0, 0, 0}, {0, 1, 1}, {1, 0, 1}, {1, 1, 0}};
  for (int iz = 0; iz < sz; iz++)
    for (int iy = 0; iy < sy; iy++)
      for (int ix = 0; ix < sx; ix++) {
        const double x = ix * s, y = iy * s, z = iz * s;
        for (int a = 0; a < 4; a++) {
          const double bx = off[a][0] ? x + hs : x;
          const double by = off[a][1] ? y + hs : y;
          const double bz = off[a][2] ? z + hs : z;
          double* p = q + n * stride;
          p[0] = bx + (orc_canonical32(&mt) * (0.1 - 0.0) + 0.0);
          p[1] = by + (orc_canonical32(&mt) * (0.1 - 0.0) + 0.0);
          p[2] = bz + (orc_canonical32(&mt) * (0.1 - 0.0) + 0.0);
          if (stride == 4) p[3] = 0.0;
          n++;
        }
      }
  return n;
}

/* x,y,z ~ U[0,L) from std::mt19937_64(seed), order x,y,z per particle. */
int64_t orc_gen_uniform(int64_t n, double L, uint64_t seed, double* q, int stride) {
  orc_mt64 mt;
  orc_mt64_seed(&mt, seed);
  for (int64_t i = 0; i < n; i++) {
    double* p = q + i * stride;
    for (int d = 0; d < 3; d++) p[d] = orc_canonical64(&mt) * (L - 0.0) + 0.0;
    if (stride == 4) p[3] = 0.0;
  }
  return n;
}

/* FNV-1a-64 over raw bytes — the fingerprint function of SURVEY.md §8c. */
uint64_t orc_fnv1a64(const void* data, int64_t nbytes) {
  const unsigned char* p = (const unsigned char*)data;
  uint64_t h = 0xcbf29ce484222325ull;
  for (int64_t i = 0; i < nbytes; i++) {
    h ^= p[i];
    h *= 0x100000001b3ull;
  }
  return h;
}

static int orc_cmp_i32(const void* a, const void* b) {
  const int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
  return (x > y) - (x < y);
}

/* make_list.cpp:120-128 (sort_neighlist): sort every CSR row ascending before comparing. */
void orc_sort_rows(int32_t* list, const int64_t* offsets, int64_t n) {
  for (int64_t i = 0; i < n; i++)
    qsort(list + offsets[i], (size_t)(offsets[i + 1] - offsets[i]), sizeof(int32_t), orc_cmp_i32);
}

/* kernel_impl.cuh:30 / make_list.cu:178-182: the reference GPU layout list[k*N + i], padded with -1
 * (neighlist_gpu.hpp:271-274). */
int orc_ell_from_csr(const int32_t* list, const int64_t* offsets, int64_t n, int32_t rows, int32_t* ell) {
  for (int64_t t = 0; t < (int64_t)rows * n; t++) ell[t] = -1;
  for (int64_t i = 0; i < n; i++) {
    const int64_t c = offsets[i + 1] - offsets[i];
    if (c > rows) return -1;
    for (int64_t k = 0; k < c; k++) ell[k * n + i] = list[offsets[i] + k];
  }
  return 0;
}

void orc_free(void* p) { free(p); }

#define REAL double
#define SUF _f64
#define FMA fma
#define NEXTAFTER nextafter
#include "nlist_oracle_impl.h"
#undef REAL
#undef SUF
#undef FMA
#undef NEXTAFTER

#define REAL float
#define SUF _f32
#define FMA fmaf
#define NEXTAFTER nextafterf
#include "nlist_oracle_impl.h"
