/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Precision-generic body of the CPU restatement of the reference's
 * Verlet-list build.  Included twice by nlist_oracle.c with REAL = double / float and SUF = _f64 / _f32.
 *
 * Every function cites the reference file:line (relative to /root/reference) whose behaviour it restates.
 * Nothing here is compiled into, linked by, or called from the product library (md_neighbor_list_b200/).
 */

#ifndef REAL
#error "include from nlist_oracle.c"
#endif

#define CAT2(a, b) a##b
#define CAT(a, b) CAT2(a, b)
#define FN(name) CAT(name, SUF)

/* r2 in the three rounding orders that occur in the reference builds (SURVEY.md §8c "Rounding sensitivity"):
 *   order 0 (canonical): fma(dz,dz, fma(dy,dy, dx*dx))  — g++/nvcc contraction of `dx*dx + dy*dy + dz*dz`
 *                         (neighlist_cpu.hpp:222, kernel_impl.cuh:28)
 *   order 1            : fma(dx,dx, fma(dy,dy, dz*dz))  — explicit intrinsics, neighlist_cpu_avx2.hpp:555,
 *                         neighlist_cpu_avx512.hpp:364
 *   order 2            : (dx*dx + dy*dy) + dz*dz with every operation rounded (no contraction)
 * The file is compiled with -ffp-contract=off so only the explicit fma calls fuse. */
static inline REAL FN(orc_r2)(const REAL* qi, const REAL* qj, int order) {
  const REAL dx = qj[0] - qi[0];
  const REAL dy = qj[1] - qi[1];
  const REAL dz = qj[2] - qi[2];
  if (order == 0) return FMA(dz, dz, FMA(dy, dy, dx * dx));
  if (order == 1) return FMA(dx, dx, FMA(dy, dy, dz * dz));
  {
    volatile REAL a = dx * dx, b = dy * dy, c = dz * dz;
    volatile REAL ab = a + b;
    return ab + c;
  }
}

/* neighlist_cpu.hpp:380-395 (ctor) + 408-411 (Initialize): mesh_size = int(L/SL), ms = L/mesh_size,
 * ims = 1/ms, SL2 = SL*SL — all in the working precision for the GPU class (neighlist_gpu.hpp:236-255);
 * the CPU class keeps SL2 in double (neighlist_cpu.hpp:12,394).  With REAL=double both agree. */
typedef struct {
  int32_t mesh[3];
  int64_t nmesh;
  REAL ms[3], ims[3];
  REAL sl2;
} FN(orc_grid);

static int FN(orc_make_grid)(double sl, double lx, double ly, double lz, FN(orc_grid) * g) {
  const REAL slr = (REAL)sl;
  const REAL l[3] = {(REAL)lx, (REAL)ly, (REAL)lz};
  for (int d = 0; d < 3; d++) {
    g->mesh[d] = (int32_t)(l[d] / slr);
    if (g->mesh[d] < 3) return -1; /* mesh_size >= 3 is a precondition */
    g->ms[d] = l[d] / (REAL)g->mesh[d];
    g->ims[d] = (REAL)(1.0 / (double)g->ms[d]); /* `1.0 / ms_` is a double division, neighlist_gpu.hpp:250-252 */
  }
  g->nmesh = (int64_t)g->mesh[0] * g->mesh[1] * g->mesh[2];
  g->sl2 = slr * slr;
  return 0;
}

/* neighlist_cpu.hpp:61-66 (ApplyPBC): one-period wrap of a cell index. */
static inline void FN(orc_apply_pbc)(const FN(orc_grid) * g, int32_t* idx) {
  for (int d = 0; d < 3; d++) {
    if (idx[d] < 0) idx[d] += g->mesh[d];
    if (idx[d] >= g->mesh[d]) idx[d] -= g->mesh[d];
  }
}

/* neighlist_cpu.hpp:42-59 (GenHash): idx = int(q*ims) (truncation, reciprocal multiply), wrap, linearise.
 * `gpu_clamp` selects the GPU kernel's rule instead (neighlist_gpu.hpp:33-39: idx==mesh -> idx-1, no wrap). */
static inline int64_t FN(orc_hash)(const FN(orc_grid) * g, const REAL* q, int gpu_clamp) {
  int32_t idx[3] = {(int32_t)(q[0] * g->ims[0]), (int32_t)(q[1] * g->ims[1]), (int32_t)(q[2] * g->ims[2])};
  if (gpu_clamp) {
    for (int d = 0; d < 3; d++)
      if (idx[d] == g->mesh[d]) idx[d]--;
  } else {
    FN(orc_apply_pbc)(g, idx);
  }
  for (int d = 0; d < 3; d++)
    if (idx[d] < 0 || idx[d] >= g->mesh[d]) return -1; /* reference: out-of-range write (UB) */
  return idx[0] + ((int64_t)idx[1] + (int64_t)idx[2] * g->mesh[1]) * g->mesh[0];
}

/* neighlist_cpu.hpp:134-165 (MakeMeshidOfPtcl + MakeNextDest): histogram, exclusive scan -> mesh_index[M+1],
 * stable counting-sort permutation ptcl_id_in_mesh[N] (ids ascending inside a cell). */
int FN(orc_bin)(const REAL* q, int64_t n, int stride, double sl, double lx, double ly, double lz, int gpu_clamp,
                int64_t* mesh_index /*[M+1]*/, int32_t* ptcl_id_in_mesh /*[n]*/, int32_t* cell_of /*[n] or NULL*/) {
  FN(orc_grid) g;
  if (FN(orc_make_grid)(sl, lx, ly, lz, &g)) return -1;
  const int64_t M = g.nmesh;
  int64_t* cursor = (int64_t*)calloc((size_t)M + 1, sizeof(int64_t));
  int32_t* cell = (int32_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(int32_t));
  if (!cursor || !cell) return -2;
  for (int64_t i = 0; i < n; i++) {
    const int64_t h = FN(orc_hash)(&g, q + i * stride, gpu_clamp);
    if (h < 0) {
      free(cursor);
      free(cell);
      return -3;
    }
    cell[i] = (int32_t)h;
    cursor[h + 1]++;
  }
  mesh_index[0] = 0;
  for (int64_t m = 0; m < M; m++) mesh_index[m + 1] = mesh_index[m] + cursor[m + 1];
  for (int64_t m = 0; m < M; m++) cursor[m] = mesh_index[m];
  for (int64_t i = 0; i < n; i++) ptcl_id_in_mesh[cursor[cell[i]]++] = (int32_t)i;
  if (cell_of) memcpy(cell_of, cell, (size_t)n * sizeof(int32_t));
  free(cursor);
  free(cell);
  return 0;
}

/* Growable (key, partner) pair buffer — replaces the reference's fixed MAX_PARTNERS*N arrays
 * (neighlist_cpu.hpp:37,76-78) whose overflow is silent UB. */
typedef struct {
  int32_t* key;
  int32_t* partner;
  int64_t n, cap;
} FN(orc_pairs);

static int FN(orc_push)(FN(orc_pairs) * p, int32_t k, int32_t j) {
  if (p->n == p->cap) {
    const int64_t nc = p->cap ? p->cap * 2 : (1 << 20);
    int32_t* nk = (int32_t*)realloc(p->key, (size_t)nc * sizeof(int32_t));
    if (!nk) return -1;
    p->key = nk;
    int32_t* np_ = (int32_t*)realloc(p->partner, (size_t)nc * sizeof(int32_t));
    if (!np_) return -1;
    p->partner = np_;
    p->cap = nc;
  }
  p->key[p->n] = k;
  p->partner[p->n] = j;
  p->n++;
  return 0;
}

/*
 * HALF list (CPU semantics).  Restates neighlist_cpu.hpp:
 *   107-132 MakeNeighMeshId   — the first 13 of the 27 (jz,jy,jx) offsets, cell indices wrapped periodically
 *   196-213 MakeNeighMeshPtclId — per-cell id list = own cell ++ 13 lower neighbour cells
 *   272-293 MakePairListFusedLoop — particle at local slot l tests list entries k > l
 *   215-237 RegistInteractPair — d = qj - qi, reject iff r2 > SL2, key = min(i,j), partner = max(i,j)
 *   361-377 MakeNeighListForEachPtcl — exclusive scan -> key_pointer[N+1], bucket scatter in discovery order
 * Offsets are 64-bit (SURVEY.md §7 "Index width").  Distances are NOT periodic.
 * Outputs: number_of_partners[n], key_pointer[n+1], *sorted_list (malloc'ed, caller frees with orc_free).
 * Returns total pairs or a negative error.
 */
int64_t FN(orc_build_half)(const REAL* q, int64_t n, int stride, double sl, double lx, double ly, double lz,
                           int order, int32_t* number_of_partners, int64_t* key_pointer, int32_t** sorted_list,
                           int64_t* candidates_tested) {
  FN(orc_grid) g;
  if (FN(orc_make_grid)(sl, lx, ly, lz, &g)) return -1;
  const int64_t M = g.nmesh;
  int64_t* mesh_index = (int64_t*)malloc(((size_t)M + 1) * sizeof(int64_t));
  int32_t* pid = (int32_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(int32_t));
  if (!mesh_index || !pid) return -2;
  int rc = FN(orc_bin)(q, n, stride, sl, lx, ly, lz, 0, mesh_index, pid, NULL);
  if (rc) {
    free(mesh_index);
    free(pid);
    return rc;
  }
  FN(orc_pairs) pairs = {0, 0, 0, 0};
  memset(number_of_partners, 0, (size_t)n * sizeof(int32_t));
  int64_t tested = 0;
  int32_t* fused = NULL;
  int64_t fused_cap = 0;
  for (int32_t iz = 0; iz < g.mesh[2]; iz++)
    for (int32_t iy = 0; iy < g.mesh[1]; iy++)
      for (int32_t ix = 0; ix < g.mesh[0]; ix++) {
        const int64_t imesh = ix + ((int64_t)iy + (int64_t)iz * g.mesh[1]) * g.mesh[0];
        const int64_t ibeg = mesh_index[imesh], iend = mesh_index[imesh + 1];
        const int64_t isize = iend - ibeg;
        /* fused id list: own cell first, then the 13 neighbour cells in table order */
        int64_t nf = 0;
        int32_t jm = 0;
        int done = 0;
        int64_t need = isize;
        int64_t nbeg[13], nend[13];
        for (int32_t jz = -1; jz < 2 && !done; jz++)
          for (int32_t jy = -1; jy < 2 && !done; jy++)
            for (int32_t jx = -1; jx < 2 && !done; jx++) {
              int32_t idx[3] = {ix + jx, iy + jy, iz + jz};
              FN(orc_apply_pbc)(&g, idx);
              const int64_t jmesh = idx[0] + ((int64_t)idx[1] + (int64_t)idx[2] * g.mesh[1]) * g.mesh[0];
              nbeg[jm] = mesh_index[jmesh];
              nend[jm] = mesh_index[jmesh + 1];
              need += nend[jm] - nbeg[jm];
              jm++;
              if (jm == 13) done = 1;
            }
        if (need > fused_cap) {
          fused_cap = need * 2 + 64;
          fused = (int32_t*)realloc(fused, (size_t)fused_cap * sizeof(int32_t));
          if (!fused) return -2;
        }
        for (int64_t k = ibeg; k < iend; k++) fused[nf++] = pid[k];
        for (int32_t m = 0; m < 13; m++)
          for (int64_t k = nbeg[m]; k < nend[m]; k++) fused[nf++] = pid[k];
        for (int64_t l = 0; l < isize; l++) {
          const int32_t i = pid[l + ibeg];
          const REAL* qi = q + (int64_t)i * stride;
          for (int64_t k = l + 1; k < nf; k++) {
            const int32_t j = fused[k];
            tested++;
            const REAL r2 = FN(orc_r2)(qi, q + (int64_t)j * stride, order);
            if (r2 > g.sl2) continue;
            const int32_t a = i < j ? i : j, b = i < j ? j : i;
            if (FN(orc_push)(&pairs, a, b)) return -2;
            number_of_partners[a]++;
          }
        }
      }
  free(fused);
  key_pointer[0] = 0;
  for (int64_t i = 0; i < n; i++) key_pointer[i + 1] = key_pointer[i] + number_of_partners[i];
  int32_t* list = (int32_t*)malloc((size_t)(pairs.n > 0 ? pairs.n : 1) * sizeof(int32_t));
  int64_t* cur = (int64_t*)malloc(((size_t)n + 1) * sizeof(int64_t));
  if (!list || !cur) return -2;
  memcpy(cur, key_pointer, ((size_t)n + 1) * sizeof(int64_t));
  for (int64_t p = 0; p < pairs.n; p++) list[cur[pairs.key[p]]++] = pairs.partner[p];
  free(cur);
  free(pairs.key);
  free(pairs.partner);
  free(mesh_index);
  free(pid);
  *sorted_list = list;
  if (candidates_tested) *candidates_tested = tested;
  return key_pointer[n];
}

/*
 * FULL list (GPU semantics).  Restates the reference kernels' contract in CSR form:
 *   neighlist_gpu.hpp:125-142 MakeNeighMeshId — all 27 offsets, wrapped cell indices
 *   kernel_impl.cuh:17-32 make_neighlist_naive — for every stencil cell, every j in it, d = qi - qj,
 *       skip iff (r2 > SL2 || j == i), row i receives j
 *   neighlist_gpu.hpp:33-39 make_mesh — binning with the `idx == mesh_size -> idx-1` clamp
 * The reference stores rows in a transposed ELL matrix list[k*N+i] (kernel_impl.cuh:30) without offsets; here the
 * same rows are stored in CSR with 64-bit offsets, and orc_ell_from_csr() below produces the transposed view.
 * Row order = discovery order (stencil-cell order, ids ascending inside a cell).
 */
int64_t FN(orc_build_full)(const REAL* q, int64_t n, int stride, double sl, double lx, double ly, double lz,
                           int order, int32_t* number_of_partners, int64_t* offsets, int32_t** list_out,
                           int64_t* candidates_tested) {
  FN(orc_grid) g;
  if (FN(orc_make_grid)(sl, lx, ly, lz, &g)) return -1;
  const int64_t M = g.nmesh;
  int64_t* mesh_index = (int64_t*)malloc(((size_t)M + 1) * sizeof(int64_t));
  int32_t* pid = (int32_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(int32_t));
  int32_t* cell_of = (int32_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(int32_t));
  if (!mesh_index || !pid || !cell_of) return -2;
  int rc = FN(orc_bin)(q, n, stride, sl, lx, ly, lz, 1, mesh_index, pid, cell_of);
  if (rc) return rc;
  int64_t tested = 0;
  int32_t* list = NULL;
  for (int pass = 0; pass < 2; pass++) {
    for (int64_t i = 0; i < n; i++) {
      const int64_t c = cell_of[i];
      const int32_t ix = (int32_t)(c % g.mesh[0]);
      const int32_t iy = (int32_t)((c / g.mesh[0]) % g.mesh[1]);
      const int32_t iz = (int32_t)(c / ((int64_t)g.mesh[0] * g.mesh[1]));
      const REAL* qi = q + i * stride;
      int64_t w = pass ? offsets[i] : 0;
      int32_t cnt = 0;
      for (int32_t jz = -1; jz < 2; jz++)
        for (int32_t jy = -1; jy < 2; jy++)
          for (int32_t jx = -1; jx < 2; jx++) {
            int32_t idx[3] = {ix + jx, iy + jy, iz + jz};
            FN(orc_apply_pbc)(&g, idx);
            const int64_t jmesh = idx[0] + ((int64_t)idx[1] + (int64_t)idx[2] * g.mesh[1]) * g.mesh[0];
            for (int64_t k = mesh_index[jmesh]; k < mesh_index[jmesh + 1]; k++) {
              const int32_t j = pid[k];
              if (!pass) tested++;
              /* d = qi - qj in the kernels; the squares make the sign irrelevant */
              const REAL r2 = FN(orc_r2)(qi, q + (int64_t)j * stride, order);
              if (r2 > g.sl2 || j == i) continue;
              if (pass) list[w++] = j;
              cnt++;
            }
          }
      if (!pass) number_of_partners[i] = cnt;
    }
    if (!pass) {
      offsets[0] = 0;
      for (int64_t i = 0; i < n; i++) offsets[i + 1] = offsets[i] + number_of_partners[i];
      list = (int32_t*)malloc((size_t)(offsets[n] > 0 ? offsets[n] : 1) * sizeof(int32_t));
      if (!list) return -2;
    }
  }
  free(mesh_index);
  free(pid);
  free(cell_of);
  *list_out = list;
  if (candidates_tested) *candidates_tested = tested;
  return offsets[n];
}

/* make_list.cpp:79-99 (half, j > i) and make_list.cu:79-98 (full, all j != i): the O(N^2) brute force the
 * reference drivers use as their self-check.  Emits CSR (rows ascending by construction). */
int64_t FN(orc_bruteforce)(const REAL* q, int64_t n, int stride, double sl, int full, int order,
                           int32_t* number_of_partners, int64_t* offsets, int32_t** list_out) {
  const REAL slr = (REAL)sl;
  const REAL sl2 = slr * slr; /* make_list.cpp:24, make_list.cu:24 */
  int32_t* list = NULL;
  for (int pass = 0; pass < 2; pass++) {
    for (int64_t i = 0; i < n; i++) {
      int64_t w = pass ? offsets[i] : 0;
      int32_t cnt = 0;
      const REAL* qi = q + i * stride;
      for (int64_t j = full ? 0 : i + 1; j < n; j++) {
        if (j == i) continue;
        const REAL r2 = FN(orc_r2)(qi, q + j * stride, order);
        if (r2 > sl2) continue;
        if (pass) list[w++] = (int32_t)j;
        cnt++;
      }
      if (!pass) number_of_partners[i] = cnt;
    }
    if (!pass) {
      offsets[0] = 0;
      for (int64_t i = 0; i < n; i++) offsets[i + 1] = offsets[i] + number_of_partners[i];
      list = (int32_t*)malloc((size_t)(offsets[n] > 0 ? offsets[n] : 1) * sizeof(int32_t));
      if (!list) return -2;
    }
  }
  *list_out = list;
  return offsets[n];
}

/* The "1-ulp band" report of BASELINE.json's north_star / SURVEY.md §8c: over every candidate pair of the 27-cell
 * stencil (each unordered pair once), count pairs whose verdict differs between the three rounding orders and
 * pairs whose canonical r2 lies within 1 ulp of SL2.  Writes up to `cap` such pairs (i,j) into band_pairs. */
int64_t FN(orc_band_report)(const REAL* q, int64_t n, int stride, double sl, double lx, double ly, double lz,
                            int64_t* n_order_dependent, int64_t* n_within_1ulp, int32_t* band_pairs, int64_t cap) {
  FN(orc_grid) g;
  if (FN(orc_make_grid)(sl, lx, ly, lz, &g)) return -1;
  const int64_t M = g.nmesh;
  int64_t* mesh_index = (int64_t*)malloc(((size_t)M + 1) * sizeof(int64_t));
  int32_t* pid = (int32_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(int32_t));
  int32_t* cell_of = (int32_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(int32_t));
  if (!mesh_index || !pid || !cell_of) return -2;
  int rc = FN(orc_bin)(q, n, stride, sl, lx, ly, lz, 1, mesh_index, pid, cell_of);
  if (rc) return rc;
  const REAL lo = NEXTAFTER(g.sl2, (REAL)0), hi = NEXTAFTER(g.sl2, (REAL)1e30);
  int64_t nod = 0, nulp = 0, nb = 0;
  for (int64_t i = 0; i < n; i++) {
    const int64_t c = cell_of[i];
    const int32_t ix = (int32_t)(c % g.mesh[0]);
    const int32_t iy = (int32_t)((c / g.mesh[0]) % g.mesh[1]);
    const int32_t iz = (int32_t)(c / ((int64_t)g.mesh[0] * g.mesh[1]));
    const REAL* qi = q + i * stride;
    for (int32_t jz = -1; jz < 2; jz++)
      for (int32_t jy = -1; jy < 2; jy++)
        for (int32_t jx = -1; jx < 2; jx++) {
          int32_t idx[3] = {ix + jx, iy + jy, iz + jz};
          FN(orc_apply_pbc)(&g, idx);
          const int64_t jmesh = idx[0] + ((int64_t)idx[1] + (int64_t)idx[2] * g.mesh[1]) * g.mesh[0];
          for (int64_t k = mesh_index[jmesh]; k < mesh_index[jmesh + 1]; k++) {
            const int32_t j = pid[k];
            if (j <= i) continue;
            const REAL* qj = q + (int64_t)j * stride;
            const REAL r0 = FN(orc_r2)(qi, qj, 0), r1 = FN(orc_r2)(qi, qj, 1), r2 = FN(orc_r2)(qi, qj, 2);
            const int v0 = !(r0 > g.sl2), v1 = !(r1 > g.sl2), v2 = !(r2 > g.sl2);
            const int od = (v0 != v1) || (v0 != v2);
            const int ulp = (r0 >= lo && r0 <= hi);
            nod += od;
            nulp += ulp;
            if ((od || ulp) && nb < cap) {
              band_pairs[2 * nb] = (int32_t)i;
              band_pairs[2 * nb + 1] = j;
              nb++;
            }
          }
        }
  }
  free(mesh_index);
  free(pid);
  free(cell_of);
  *n_order_dependent = nod;
  *n_within_1ulp = nulp;
  return nb;
}

#undef FN
#undef CAT
#undef CAT2
