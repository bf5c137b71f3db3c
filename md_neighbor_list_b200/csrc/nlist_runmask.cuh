// nlist_runmask.cuh — "run masks": the search + emission pair of FULL and HALF lists (round 2).
//
// What changed against the pair masks (nlist_kernels.cuh) and why.  There the unit of work is (cell A, 256 candidates
// of A's 27-cell stencil): the ~35 particles i of A are the BITS of a word, the candidates j sit on the lanes, and by
// symmetry the word a lane ends with is a piece of ROW j ("which particles of A are partners of j").  A cell of the
// default system holds 35.3 particles on average: two 32-bit words per (row, stencil cell), the second one a tenth
// full, 27 x 2-3 sparse words per row for the popcount pass and the emission to walk, and a per-unit set-up
// (run table, 27 translation entries, candidate look-ups) that is 44 % of the kernel's instructions and runs at a
// fraction of the main loop's issue rate (profiles/r01_ncu_final_pairmask.txt).
//
// Here the bits of a word are the particles of a whole x-RUN: the <= 3 cells (cx-1..cx+1, cy, cz) around a centre cell
// C, which are contiguous in the cell-sorted arrays (~106 particles: 3.3 words, dense).  The candidates on the lanes
// are the particles of C's COLUMN: the <= 9 cells (cx, cy-1..cy+1, cz-1..cz+1).  Every (run particle i, column
// particle j) pair lies inside each other's stencil, each ordered pair of the system is met exactly once, and the
// words lane j ends with are the complete, dense bit vector of ROW j over one of its nine runs:
//     mask[(rho * WR + w) * n_cap + slot_j],  bit (31 - b)  <=>  the particle in slot  run_start + 32 w + b  is a partner
// with rho = (oz, oy) the run's ordinal in j's 3 x 3 stencil of runs.  Consequences:
//   * 9 x ~3.3 dense words per row instead of 27 x 2-3 sparse ones, and bit -> slot is run_start + bit index: the
//     emission needs 2 cell_start values per run instead of 4 and no per-cell word bookkeeping;
//   * a unit tests ~106 rows against its candidates instead of ~35: the per-unit set-up (9-entry column table, row
//     staging, candidate look-ups) is amortised over three times the main-loop work;
//   * the row's own bit is cleared where it is produced, and the row length leaves the search kernel by one
//     RED.ADD per (row, run): the popcount pass (15 us on the default system) is gone.
//   * HALF lists (row j keeps the larger ids): a cell is ordered by the id its rows report, so the kept rows of each
//     of a run's <= 3 cells are a suffix — one binary search per (candidate, cell), range masks per word, no per-test
//     cost (runmask_kernel<HALF>).
// The test itself is round 1's: dot form d = xi.xj - |xj|^2/2 - (|xi|^2 - SL^2)/2 in FP32 with packed FFMA2/FADD2,
// sign bit funnel-shifted into the word, min|d| tracked, exact input-precision re-test inside the band E.
#pragma once

#include "nlist_kernels.cuh"

namespace nlb {

#ifndef NLB_RN_THREADS
#define NLB_RN_THREADS 128
#endif
constexpr int RN_THREADS = NLB_RN_THREADS;  // warps are independent; the CTA only groups them
#ifndef NLB_RN_MINB
#define NLB_RN_MINB 5
#endif
#ifndef NLB_RN_RJ
#define NLB_RN_RJ 4
#endif
constexpr int RN_RJ = NLB_RN_RJ;      // candidates per lane of a full chunk (packed in pairs)
constexpr int RN_CH = 32 * RN_RJ;     // candidates per chunk
constexpr int RN_SW = 8;              // row words (of 32 rows) staged per round: runs of up to 256 particles in one
__host__ __device__ inline int rn_staged_words(int wr) { return wr < RN_SW ? wr : RN_SW; }
// per-warp shared memory: staged rows {xi, yi, zi, -ai} (+ one row of padding: the main loop requests a row ahead);
// column table: 9 int4 {end of the cell in the candidate list, first slot minus start in the list, mask plane
// rho * WR (or -1), own-cell flag}; 9 float2 {ty, tz}
__host__ __device__ inline size_t rn_warp_bytes(int wr) {
  return (size_t)(rn_staged_words(wr) * 32 + 1) * sizeof(float4) + 9 * sizeof(int4) + 10 * sizeof(float2);
}

template <typename T>
struct RunMaskArgs {
  const T* q;  // caller's positions (band re-test only)
  GridParams<T> gp;
  const int32_t* cell_start;
  const float4* rec;  // cell-sorted records relative to the particle's own cell corner, .w = local id
  const int32_t* sorted_ids;
  const int32_t* cut_ids;  // HALF: the id per slot the rows are cut by — sorted_ids, or the global id per slot
  int32_t cut_by_slot_table;  // HALF: the candidate's key is cut_ids[its slot] (global ids) instead of its local id
  int32_t n_owned;
  uint32_t* mask;  // [9][wr][n_cap]
  long long n_cap;
  int32_t wr;
  int32_t fits32;  // 9 * wr * n_cap < 2^32
  FastDiv d_parts, d_mx, d_my;  // item -> (cell, part), cell -> (cx, cy, cz)
  float band;
  unsigned long long* queue;
  int32_t parts;  // units per cell: part p takes the candidate chunks p, p + parts, ...
  int32_t grab;   // units drawn per atomic
  int32_t* counts;  // zeroed by cellsort_kernel
  DeviceStatus* st;
};

// One chunk of RJ x 32 candidates against every row of the run.  `staged` tells whether the (single) row round of
// the run already sits in shared memory (runs of more than RN_SW words are re-staged round by round).
template <typename T, int STRIDE, int RJ, bool HALF>
__device__ __forceinline__ void rn_chunk(const RunMaskArgs<T>& a, float4* si, const int4* t_col, const float2* t_tr,
                                         const int lane, const int32_t c0, const int32_t ncand, const int32_t rs,
                                         const int32_t n_r, const int32_t b1, const int32_t b2, const float tx0,
                                         bool& staged, unsigned long long& band_local) {
  const GridParams<T>& gp = a.gp;
  const float msx = gp.msf[0], msy = gp.msf[1], msz = gp.msf[2];
  const float hx = 0.5f * msx, hy = 0.5f * msy, hz = 0.5f * msz;
  const uint32_t ncap32 = (uint32_t)a.n_cap;
  const int32_t nw = (n_r + 31) >> 5;

  float xj[RJ], yj[RJ], zj[RJ], wj[RJ];
  int32_t sj[RJ];    // candidate's slot
  int32_t oj[RJ];    // first mask plane of this run in the candidate's row (rho * wr); -1: tail lane or ghost row
  int32_t idj[RJ];   // candidate's local id
  int32_t selfw[RJ];  // word of the run that holds the candidate itself (-1: the run is not the candidate's own)
  uint32_t selfm[RJ];
  int32_t pc[RJ];    // partners found in this run
  {
    // slots first (table only), then every record load, then the translation: the loads are in flight together
    int32_t run = 0;
    int4 cur = t_col[0];
    int32_t tc[RJ];
#pragma unroll
    for (int k = 0; k < RJ; k++) {
      const int32_t c = c0 + k * 32 + lane;
      sj[k] = rs;  // tail lanes read a present record
      tc[k] = -1;
      oj[k] = -1;
      selfw[k] = -1;
      selfm[k] = 0u;
      if (c < ncand) {
        while (c >= cur.x) cur = t_col[++run];  // c < ncand = end of entry 8: stops at run <= 8
        sj[k] = cur.y + c;
        tc[k] = run;
        oj[k] = cur.z;
        if (cur.w) {
          const int32_t bs = sj[k] - rs;
          selfw[k] = bs >> 5;
          selfm[k] = 0x80000000u >> (bs & 31);
        }
      }
    }
    float4 rjv[RJ];
#pragma unroll
    for (int k = 0; k < RJ; k++) rjv[k] = __ldg(a.rec + sj[k]);
#pragma unroll
    for (int k = 0; k < RJ; k++) {
      xj[k] = yj[k] = zj[k] = 0.f;
      wj[k] = -1.0e30f;  // d = -1e30: a miss, far from the band
      idj[k] = 0;
      pc[k] = 0;
      if (tc[k] >= 0) {
        const float4 rj = rjv[k];
        const float2 tr = t_tr[tc[k]];
        xj[k] = rj.x - hx;
        yj[k] = fmaf(tr.x, msy, rj.y);
        zj[k] = fmaf(tr.y, msz, rj.z);
        wj[k] = -0.5f * fmaf(xj[k], xj[k], fmaf(yj[k], yj[k], zj[k] * zj[k]));
        idj[k] = __float_as_int(rj.w);
        if (idj[k] >= a.n_owned) oj[k] = -1;  // a ghost: no row
      }
    }
  }
  {
    // a chunk whose candidates are all ghosts (the outer cell layers of a slab rank) produces no row
    bool row_needed = false;
#pragma unroll
    for (int k = 0; k < RJ; k++) row_needed = row_needed || oj[k] >= 0;
    if (!__any_sync(0xffffffffu, row_needed)) return;
  }
  // HALF lists (row j keeps the partners with a larger id — the id its rows report): the ids of a cell ascend with the slot, so inside each of
  // the run's <= 3 cells the kept rows are a SUFFIX.  cut[k][c] = run-relative index of the first kept row of cell c for
  // candidate k: cell start + #{ids of the cell <= id_j}, found by a branch-free binary search (uniform step count, the
  // 3 * RJ searches of a lane interleaved); every word is then cut with at most three range masks — no per-test cost.
  int32_t cut[RJ][3];
  int32_t cs_[4];  // run-relative starts of the three cells and the end of the run
  if (HALF) {
    cs_[0] = 0;
    cs_[1] = min(b1 == 0x7fffffff ? n_r : b1 - rs, n_r);
    cs_[2] = min(b2 == 0x7fffffff ? n_r : b2 - rs, n_r);
    cs_[3] = n_r;
    const int32_t maxlen = max(cs_[1], max(cs_[2] - cs_[1], n_r - cs_[2]));
    int32_t pos[RJ][3];
#pragma unroll
    for (int k = 0; k < RJ; k++)
#pragma unroll
      for (int c = 0; c < 3; c++) pos[k][c] = 0;
    // (a local -> global id map: cellsort_kernel orders a cell by GLOBAL id, so the suffix property holds for the
    //  global ids the rows report; key and table are then the global ids per slot)
    const int32_t* ids = a.cut_ids + rs;
    int32_t key[RJ];
#pragma unroll
    for (int k = 0; k < RJ; k++) key[k] = a.cut_by_slot_table ? __ldg(a.cut_ids + sj[k]) : idj[k];
    for (int32_t bit = 1 << (31 - __clz(max(maxlen, 1))); bit > 0; bit >>= 1) {
#pragma unroll
      for (int k = 0; k < RJ; k++)
#pragma unroll
        for (int c = 0; c < 3; c++) {
          const int32_t t = pos[k][c] + bit;
          const bool in = t <= cs_[c + 1] - cs_[c];
          const int32_t v = in ? __ldg(ids + cs_[c] + t - 1) : 0x7fffffff;
          if (v <= key[k]) pos[k][c] = t;
        }
    }
#pragma unroll
    for (int k = 0; k < RJ; k++)
#pragma unroll
      for (int c = 0; c < 3; c++) cut[k][c] = cs_[c] + pos[k][c];
  }
  f32x2 X[RJ / 2], Y[RJ / 2], Z[RJ / 2], W[RJ / 2];
#pragma unroll
  for (int h = 0; h < RJ / 2; h++) {
    X[h] = pack2(xj[2 * h], xj[2 * h + 1]);
    Y[h] = pack2(yj[2 * h], yj[2 * h + 1]);
    Z[h] = pack2(zj[2 * h], zj[2 * h + 1]);
    W[h] = pack2(wj[2 * h], wj[2 * h + 1]);
  }
  for (int32_t iw0 = 0; iw0 < nw; iw0 += RN_SW) {
    const int32_t iw1 = min(iw0 + RN_SW, nw);
    if (!staged) {
      __syncwarp();  // the previous round's readers are done
      for (int32_t k = iw0 * 32 + lane; k < min(n_r, iw1 * 32); k += 32) {
        const int32_t s = rs + k;
        const float4 r = __ldg(a.rec + s);
        const float tx = tx0 + ((s >= b1) ? 1.f : 0.f) + ((s >= b2) ? 1.f : 0.f);
        const float x = fmaf(tx, msx, r.x), y = r.y - hy, z = r.z - hz;
        const float nai = -0.5f * (fmaf(x, x, fmaf(y, y, z * z)) - gp.sl2f);
        si[k - iw0 * 32] = make_float4(x, y, z, nai);
      }
      __syncwarp();
      staged = nw <= RN_SW;  // a run of one round stays staged for the unit's later chunks
    }
    for (int32_t w = iw0; w < iw1; w++) {
      const int32_t cnt = min(32, n_r - w * 32);
      const float4* sp = si + (w - iw0) * 32;
      uint32_t miss[RJ];
#pragma unroll
      for (int k = 0; k < RJ; k++) miss[k] = 0u;
      float mh[RJ / 2];  // min |d| per candidate pair
#pragma unroll
      for (int h = 0; h < RJ / 2; h++) mh[h] = 3.0e38f;
      // one broadcast LDS.128 per row, requested a row ahead; ptxas folds the {v, v} pairs into scalar-broadcast
      // operands of FFMA2 / FADD2 (R.F32), so a row costs no register moves
      float4 nxt = sp[0];
#pragma unroll 4
      for (int32_t ii = 0; ii < cnt; ii++) {
        const float4 r = nxt;
        nxt = sp[ii + 1];  // row cnt: the next block's first row or the padding row, never used
        const f32x2 xx = pack2(r.x, r.x), yy = pack2(r.y, r.y), zz = pack2(r.z, r.z), aa = pack2(r.w, r.w);
#pragma unroll
        for (int h = 0; h < RJ / 2; h++) {
          const f32x2 d2 = add2(fma2(xx, X[h], fma2(yy, Y[h], fma2(zz, Z[h], W[h]))), aa);
          float d0, d1;
          unpack2(d2, d0, d1);
          miss[2 * h] = __funnelshift_l(__float_as_uint(d0), miss[2 * h], 1);  // shift the sign bit in
          miss[2 * h + 1] = __funnelshift_l(__float_as_uint(d1), miss[2 * h + 1], 1);
          mh[h] = fminf(mh[h], fminf(fabsf(d0), fabsf(d1)));
        }
      }
      uint32_t hits[RJ];
#pragma unroll
      for (int k = 0; k < RJ; k++) hits[k] = (~miss[k]) << (32 - cnt);  // bit (31 - ii) <-> row w*32 + ii of the run
      // tests inside the pre-filter's uncertainty band are decided exactly, in the caller's precision, by the whole
      // warp (lane = row, the triggering lane's candidate broadcast)
      float mall = mh[0];
#pragma unroll
      for (int h = 1; h < RJ / 2; h++) mall = fminf(mall, mh[h]);
      unsigned trig = __ballot_sync(0xffffffffu, mall < a.band);
      while (trig) {
        const int src = __ffs(trig) - 1;
        trig &= trig - 1;
#pragma unroll
        for (int h = 0; h < RJ / 2; h++) {
          if (!(__shfl_sync(0xffffffffu, mh[h], src) < a.band)) continue;  // warp-uniform
          const f32x2 xs = __shfl_sync(0xffffffffu, X[h], src), ys = __shfl_sync(0xffffffffu, Y[h], src);
          const f32x2 zs = __shfl_sync(0xffffffffu, Z[h], src), ws = __shfl_sync(0xffffffffu, W[h], src);
          float cx2[2], cy2[2], cz2[2], cw2[2];
          unpack2(xs, cx2[0], cx2[1]);
          unpack2(ys, cy2[0], cy2[1]);
          unpack2(zs, cz2[0], cz2[1]);
          unpack2(ws, cw2[0], cw2[1]);
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int k = 2 * h + e;
            const int32_t o_src = __shfl_sync(0xffffffffu, oj[k], src);
            const int32_t jid = __shfl_sync(0xffffffffu, idj[k], src);
            bool fix = false, hit = false;
            if (lane < cnt && o_src >= 0) {
              const float4 qi = sp[lane];
              const float d = pre_d(qi.x, qi.y, qi.z, qi.w, cx2[e], cy2[e], cz2[e], cw2[e]);
              if (fabsf(d) < a.band) {
                const int32_t iid = __ldg(a.sorted_ids + rs + w * 32 + lane);
                hit = exact_within(load_pos<T, STRIDE>(a.q, iid), load_pos<T, STRIDE>(a.q, jid), gp.sl2);
                fix = true;
                band_local++;
              }
            }
            // lane ii <-> bit (31 - ii)
            const uint32_t fixm = __brev(__ballot_sync(0xffffffffu, fix));
            const uint32_t hitm = __brev(__ballot_sync(0xffffffffu, hit));
            if (lane == src) hits[k] = (hits[k] & ~fixm) | hitm;
          }
        }
      }
      if (HALF) {
        // drop the rows [cell start, cut) of every cell that meets this word (the own bit goes with them: id == id)
        const int32_t wb = w * 32;
#pragma unroll
        for (int c = 0; c < 3; c++) {
          if (cs_[c] >= wb + 32 || cs_[c + 1] <= wb) continue;  // warp-uniform
          const uint32_t from = __funnelshift_rc(0xffffffffu, 0u, (uint32_t)max(cs_[c] - wb, 0));
#pragma unroll
          for (int k = 0; k < RJ; k++)
            hits[k] &= ~(from & ~__funnelshift_rc(0xffffffffu, 0u, (uint32_t)max(cut[k][c] - wb, 0)));
        }
      }
#pragma unroll
      for (int k = 0; k < RJ; k++) {
        if (!HALF && selfw[k] == w) hits[k] &= ~selfm[k];  // FULL rows hold j != i (kernel_impl.cuh:29)
        pc[k] += __popc(hits[k]);
      }
      if (a.fits32) {
#pragma unroll
        for (int k = 0; k < RJ; k++)
          if (oj[k] >= 0) a.mask[(uint32_t)(oj[k] + w) * ncap32 + (uint32_t)sj[k]] = hits[k];
      } else {
#pragma unroll
        for (int k = 0; k < RJ; k++)
          if (oj[k] >= 0) a.mask[(unsigned long long)(uint32_t)(oj[k] + w) * ncap32 + (uint32_t)sj[k]] = hits[k];
      }
    }
    if (iw1 < nw) staged = false;  // the next round overwrites the staged rows
  }
  // the run's share of the row length: one RED per (row, run)
#pragma unroll
  for (int k = 0; k < RJ; k++)
    if (oj[k] >= 0 && pc[k] != 0) atomicAdd(a.counts + idj[k], pc[k]);
}

template <typename T, int STRIDE, bool HALF = false>
__global__ void __launch_bounds__(RN_THREADS, NLB_RN_MINB) runmask_kernel(RunMaskArgs<T> a) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  unsigned char* wbase = smem_raw + (size_t)warp * rn_warp_bytes(a.wr);
  float4* si = reinterpret_cast<float4*>(wbase);  // [32 * min(wr, RN_SW) + 1]
  int4* t_col = reinterpret_cast<int4*>(wbase + (size_t)(rn_staged_words(a.wr) * 32 + 1) * sizeof(float4));  // [9]
  float2* t_tr = reinterpret_cast<float2*>(t_col + 9);                                                        // [9]

  const GridParams<T>& gp = a.gp;
  const int32_t mx = gp.mesh[0], my = gp.mesh[1], mz = gp.mesh[2];
  unsigned long long band_local = 0, cand_local = 0;

  const long long n_items = (long long)gp.n_cells * a.parts;
  // the first batch of every warp is static; later batches come from the queue, which starts behind the static ones
  const long long n_warps = (long long)gridDim.x * (blockDim.x >> 5);
  const long long first_dyn = n_warps * a.grab;
  long long base = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * a.grab;
  while (base < n_items) {
    long long next = 0;
    if (lane == 0) next = first_dyn + (long long)atomicAdd(a.queue, (unsigned long long)a.grab);  // in flight meanwhile
    for (long long item = base; item < base + a.grab && item < n_items; item++) {
      const int32_t cell = (int32_t)fdiv((uint32_t)item, a.d_parts), part = (int32_t)item - cell * a.parts;
      const int32_t cyz = (int32_t)fdiv((uint32_t)cell, a.d_mx);
      const int32_t cx = cell - cyz * mx;
      const int32_t cz = (int32_t)fdiv((uint32_t)cyz, a.d_my);
      const int32_t cy = cyz - cz * my;
      int xlo, xhi, ylo, yhi, zlo, zhi;
      axis_range(cx, mx, xlo, xhi);
      axis_range(cy, my, ylo, yhi);
      axis_range(cz, mz, zlo, zhi);
      const int32_t nx = xhi - xlo + 1, ny = yhi - ylo + 1, ncc = ny * (zhi - zlo + 1);
      // lanes 0..8: the column cells (cx, y, z) in stencil order; lanes 16..19: the run's cell boundaries
      int32_t v0 = 0, v1 = 0, plane = -1, own = 0;
      float ty = 0.f, tz = 0.f;
      if (lane < ncc) {
        const int lz = ny == 3 ? (lane * 11) >> 5 : (ny == 2 ? lane >> 1 : lane);  // lane / ny for lane < 9
        const int z = zlo + lz, y = ylo + lane - lz * ny;
        const int32_t* cs = a.cell_start + (y + z * my) * mx + cx;
        v0 = __ldg(cs);
        v1 = __ldg(cs + 1);
        ty = (float)(y - cy) - 0.5f;
        tz = (float)(z - cz) - 0.5f;
        // ordinal of C's run in the 3 x 3 runs of a row of that column cell
        plane = ((cz - axis_lo(z, mz)) * 3 + (cy - axis_lo(y, my))) * a.wr;
        own = (y == cy && z == cz) ? 1 : 0;
      } else if (lane >= 16 && lane < 20) {
        v0 = __ldg(a.cell_start + (cy + cz * my) * mx + xlo + min(lane - 16, nx));
      }
      const int32_t len = lane < ncc ? v1 - v0 : 0;
      int32_t incl = len;
#pragma unroll
      for (int d = 1; d < 16; d <<= 1) {
        const int32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
      }
      const int32_t ncand = __shfl_sync(0xffffffffu, incl, 8);
      const int32_t rs = __shfl_sync(0xffffffffu, v0, 16);
      const int32_t re = __shfl_sync(0xffffffffu, v0, 16 + nx);
      const int32_t b1v = __shfl_sync(0xffffffffu, v0, 17), b2v = __shfl_sync(0xffffffffu, v0, 18);
      if (part * RN_CH >= ncand || re == rs) continue;  // warp-uniform: nothing for this unit
      const int32_t b1 = nx >= 2 ? b1v : 0x7fffffff, b2 = nx >= 3 ? b2v : 0x7fffffff;
      int32_t n_r = re - rs;
      if (n_r > 32 * a.wr) {
        if (lane == 0 && part == 0) atomicOr(&a.st->flags, FLAG_CELL_WORDS);  // the build fails; stay in range
        n_r = 32 * a.wr;
      }
      __syncwarp();  // the previous unit's readers of the tables and the staged rows are done
      if (lane < 9) {
        t_col[lane] = make_int4(incl, v0 - (incl - len), plane, own);  // absent cells: len 0, end = previous end
        t_tr[lane] = make_float2(ty, tz);
      }
      __syncwarp();
      if (lane == 0 && part == 0) cand_local += (unsigned long long)(re - rs) * (unsigned long long)ncand;
      const float tx0 = (float)(xlo - cx) - 0.5f;
      bool staged = false;
      for (int32_t c0 = part * RN_CH; c0 < ncand; c0 += a.parts * RN_CH) {
        const int32_t rem = ncand - c0;
        if (RN_RJ >= 4 && rem <= RN_CH / 2) {
          if (RN_RJ >= 8 && rem <= RN_CH / 4)
            rn_chunk<T, STRIDE, (RN_RJ >= 8 ? RN_RJ / 4 : 2), HALF>(a, si, t_col, t_tr, lane, c0, ncand, rs, n_r, b1, b2, tx0,
                                                              staged, band_local);
          else
            rn_chunk<T, STRIDE, (RN_RJ >= 4 ? RN_RJ / 2 : 2), HALF>(a, si, t_col, t_tr, lane, c0, ncand, rs, n_r, b1, b2, tx0,
                                                              staged, band_local);
        } else {
          rn_chunk<T, STRIDE, RN_RJ, HALF>(a, si, t_col, t_tr, lane, c0, ncand, rs, n_r, b1, b2, tx0, staged, band_local);
        }
      }
    }
    base = __shfl_sync(0xffffffffu, next, 0);
  }
  if (cand_local) atomicAdd(&a.st->candidates, cand_local);
  if (band_local) atomicAdd(&a.st->band_tests, band_local);
}

// ---------------------------------------------------------------------------------------------------------------
// emission from run masks.  Thread = row, warp = 32 consecutive cell-sorted slots, warps independent.  Per z-plane of
// the row's stencil the 6 cell_start values and the first ER_PRE words of its three runs are requested together; a
// word's set bits are expanded MSB-first (one FLO each) into the lane's line of the warp's shared-memory tile as
// cell-sorted SLOTS (run start + bit index — nothing else to look up), and lines are flushed to the rows as in round 1:
// slot -> partner id by a gather, scalar stores up to 16-byte alignment of the row position, then 16-byte stores.
// ---------------------------------------------------------------------------------------------------------------
struct EmitRunArgs {
  const int32_t* cell_start;
  const int32_t* sorted_ids;
  const int32_t* slot_cell;
  const int32_t* slot_pid;  // partner id reported for a slot: sorted_ids, or the global ids in slot order
  int32_t mesh[3];
  FastDiv d_mx, d_my;
  int32_t n_total, n_owned, n_cells;
  const uint32_t* mask;
  long long n_cap;
  int32_t wr;
  const int64_t* offsets;
  int32_t* partners;
  long long capacity;
};

#ifndef NLB_ER_WARPS
#define NLB_ER_WARPS 2
#endif
constexpr int ER_WARPS = NLB_ER_WARPS;
#ifndef NLB_ER_MINB
#define NLB_ER_MINB 14
#endif
#ifndef NLB_ER_VEC
#define NLB_ER_VEC 8
#endif
constexpr int ER_VEC = NLB_ER_VEC;  // entries per vector store of the flush: 4 (16 bytes) or 8 (32 bytes)
constexpr int ER_PRE = 4;  // words of a run requested ahead (runs of up to 128 particles)

__global__ void __launch_bounds__(ER_WARPS * 32, NLB_ER_MINB) emitrun_kernel(EmitRunArgs a) {
  pdl_enter();
  extern __shared__ __align__(16) int32_t er_smem[];
  if (a.offsets[a.n_owned] > a.capacity) return;  // overflow already flagged by the offsets scan
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  int32_t* line = er_smem + (warp * 32 + lane) * EM_LINE;
  const uint32_t line_sa = (uint32_t)__cvta_generic_to_shared(line);
  const int32_t slot = (blockIdx.x * ER_WARPS + warp) * 32 + lane;
  int32_t id = 0x7fffffff;
  if (slot < a.n_total && slot < __ldg(a.cell_start + a.n_cells)) id = __ldg(a.sorted_ids + slot);
  const bool owned = id < a.n_owned;
  if (!__any_sync(0xffffffffu, owned)) return;
  const int32_t cell = owned ? __ldg(a.slot_cell + slot) : 0;
  const long long dst = owned ? (long long)a.offsets[id] : 0;
  int32_t fill = 0;  // entries staged in this lane's line
  int32_t done = 0;  // entries of this row already written

  // Every lane copies ITS OWN line to its row: staged slots -> partner ids (gather), scalar stores until the row
  // position is aligned, then vector stores; what does not fill a vector stays in the line.  ER_VEC = 8: 32-byte
  // stores (STG.256, sm_100), one full sector per lane — rows of a warp lie on 32 different pages when ids are random,
  // and half-sector stores then reach HBM as partial writes (profiles/r02_large_systems.md).
  auto flush = [&](bool final) {
    int32_t k = 0;
    int32_t* out = a.partners + dst + done;
    while (k < fill && ((reinterpret_cast<uintptr_t>(out + k) & (4 * ER_VEC - 1)) != 0)) {
      out[k] = __ldg(a.slot_pid + line[k]);
      k++;
    }
    while (k + ER_VEC <= fill) {
      int32_t v[ER_VEC];
#pragma unroll
      for (int u = 0; u < ER_VEC; u++) v[u] = __ldg(a.slot_pid + line[k + u]);
      if (ER_VEC == 8)
        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(out + k), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                     "r"(v[3]), "r"(v[4 % ER_VEC]), "r"(v[5 % ER_VEC]), "r"(v[6 % ER_VEC]), "r"(v[7 % ER_VEC])
                     : "memory");
      else
        *reinterpret_cast<int4*>(out + k) = make_int4(v[0], v[1], v[2], v[3]);
      k += ER_VEC;
    }
    if (final) {
      while (k < fill) {
        out[k] = __ldg(a.slot_pid + line[k]);
        k++;
      }
    }
    done += k;
    const int32_t left = fill - k;
    for (int32_t t = 0; t < left; t++) line[t] = line[k + t];
    fill = left;
  };

  const int32_t mx = a.mesh[0], my = a.mesh[1], mz = a.mesh[2];
  const int32_t byz = (int32_t)fdiv((uint32_t)cell, a.d_mx), bx = cell - byz * mx;
  const int32_t bz = (int32_t)fdiv((uint32_t)byz, a.d_my), by = byz - bz * my;
  int xlo, xhi, ylo, yhi, zlo, zhi;
  axis_range(bx, mx, xlo, xhi);
  axis_range(by, my, ylo, yhi);
  axis_range(bz, mz, zlo, zhi);
  const int32_t ny = yhi - ylo + 1, nz = zhi - zlo + 1;
  const uint32_t* mrow = a.mask + min((long long)slot, a.n_cap - 1);  // lanes past the last slot load in range
  const long long run_stride = (long long)a.wr * a.n_cap;  // words between the planes of two runs
  const int32_t pre = min(a.wr, ER_PRE);

  // The nine runs in stencil order, software-pipelined: the 2 cell_start values and the first ER_PRE words of run
  // r + 1 are requested before run r is expanded.  Every load is unconditional at a clamped address; runs a boundary
  // row does not have are dropped by nw = 0 (their planes were written by nobody).
  auto expand = [&](uint32_t word, int32_t last) {
    uint32_t wa = line_sa + 4u * (uint32_t)fill;  // shared-memory byte address of the next free entry
    fill += __popc(word);
    if (word) {
      // one FLO per entry; the loop is written in PTX so that the bit clear feeds the loop predicate directly
      // (7 instructions per entry; the compiled C loop carried a register move and a separate compare: 9)
      asm volatile(
          "{\n\t"
          ".reg .pred p;\n\t"
          ".reg .u32 pos, bit, sl;\n"
          "EXPAND_%=:\n\t"
          "bfind.u32 pos, %0;\n\t"
          "sub.s32 sl, %2, pos;\n\t"
          "shl.b32 bit, 1, pos;\n\t"
          "st.shared.s32 [%1], sl;\n\t"
          "xor.b32 %0, %0, bit;\n\t"
#ifndef NLB_ER_NO_UNROLL2
          // second entry of the pair, predicated: when it is absent the loop ends and the address is not used again
          "setp.ne.u32 p, %0, 0;\n\t"
          "@p bfind.u32 pos, %0;\n\t"
          "@p sub.s32 sl, %2, pos;\n\t"
          "@p shl.b32 bit, 1, pos;\n\t"
          "@p st.shared.s32 [%1+4], sl;\n\t"
          "@p xor.b32 %0, %0, bit;\n\t"
          "add.u32 %1, %1, 8;\n\t"
#else
          "add.u32 %1, %1, 4;\n\t"
#endif
          "setp.ne.u32 p, %0, 0;\n\t"
          "@p bra EXPAND_%=;\n\t"
          "}"
          : "+r"(word), "+r"(wa)
          : "r"(last)
          : "memory");
    }
  };
  int32_t s0n, s1n;
  uint32_t mn[ER_PRE];
  const uint32_t* mrun = mrow;
  auto request = [&](int r) {
    const int oz = r / 3, oy = r - oz * 3;
    const int32_t* cs = a.cell_start + (min(ylo + oy, my - 1) + min(zlo + oz, mz - 1) * my) * mx;
    s0n = __ldg(cs + xlo);
    s1n = __ldg(cs + xhi + 1);
#pragma unroll
    for (int u = 0; u < ER_PRE; u++) mn[u] = __ldg(mrun + (long long)min(u, pre - 1) * a.n_cap);
    mrun += run_stride;
  };
  request(0);
#pragma unroll 1
  for (int r = 0; r < 9; r++) {
    const int oz = r / 3, oy = r - oz * 3;
    const int32_t s0 = s0n;
    const bool rv = owned && (oz < nz) && (oy < ny);
    const int32_t nw = rv ? min((s1n - s0 + 31) >> 5, a.wr) : 0;  // > wr only after FLAG_CELL_WORDS
    uint32_t m[ER_PRE];
    int32_t tot = 0;
#pragma unroll
    for (int u = 0; u < ER_PRE; u++) {
      m[u] = u < nw ? mn[u] : 0u;
      tot += __popc(m[u]);
    }
    if (r < 8) request(r + 1);
    const int32_t nwmax = __reduce_max_sync(0xffffffffu, nw);
    if (nwmax == 0) continue;
    if (__any_sync(0xffffffffu, fill + tot > EM_TILE)) flush(false);  // leaves fill <= ER_VEC - 1
    if (!__any_sync(0xffffffffu, tot > EM_TILE - (ER_VEC - 1) || nw > ER_PRE)) {
      // the run fits the line: no check between its words
#pragma unroll
      for (int u = 0; u < ER_PRE; u++)
        if (u < nwmax) expand(m[u], s0 + 32 * u + 31);
    } else {
      // a long run (the row's own run can hold more partners than a line, crowded cells more than ER_PRE words):
      // word by word with a check before each — the requested words from their registers, the rest on demand
#pragma unroll
      for (int u = 0; u < ER_PRE; u++) {
        if (u < nwmax && __any_sync(0xffffffffu, m[u] != 0u)) {
          if (__any_sync(0xffffffffu, fill + __popc(m[u]) > EM_TILE)) flush(false);
          expand(m[u], s0 + 32 * u + 31);
        }
      }
      for (int32_t w = ER_PRE; w < nwmax; w++) {
        const uint32_t word = w < nw ? __ldg(mrow + ((long long)(r * a.wr + w)) * a.n_cap) : 0u;
        if (!__any_sync(0xffffffffu, word != 0u)) continue;
        if (__any_sync(0xffffffffu, fill + __popc(word) > EM_TILE)) flush(false);
        expand(word, s0 + 32 * w + 31);
      }
    }
  }
  flush(true);
}

// ---------------------------------------------------------------------------------------------------------------
// emission from run masks, partner ids through a shared-memory WINDOW.  Same decomposition as emitrun_kernel (thread =
// row, warp = 32 consecutive cell-sorted slots, lines of the warp's tile flushed by their own lanes), but the
// slot -> partner id translation happens where the bit is expanded: the 32 rows of a warp share, for a given run
// ordinal, at most a few neighbouring cells, so the ids of every slot their run can name form ONE contiguous piece of
// slot_pid (~140 ids on the default system).  The warp stages that piece once per run (coalesced), the expansion loop
// reads the id with an LDS and the lines hold final ids: the flush is a plain copy (no per-entry gather from global
// memory — in emitrun_kernel 16 M LDG whose lanes all name different cache lines: 12.5 % of the L1 data-pipe
// wavefronts and the long-scoreboard stalls of its flush).  Warps whose window does not fit (rows that straddle the
// end of a cell row, crowded cells) take the same loop with the id read from global memory.
// ---------------------------------------------------------------------------------------------------------------
#ifndef NLB_EW_WIN
#define NLB_EW_WIN 192
#endif
constexpr int EW_WIN = NLB_EW_WIN;
#ifndef NLB_EW_MINB
#define NLB_EW_MINB 13
#endif
__host__ __device__ inline size_t ew_smem_bytes() {
  return (size_t)ER_WARPS * (32 * EM_LINE + EW_WIN) * sizeof(int32_t);
}

__global__ void __launch_bounds__(ER_WARPS * 32, NLB_EW_MINB) emitwin_kernel(EmitRunArgs a) {
  pdl_enter();
  extern __shared__ __align__(16) int32_t er_smem[];
  if (a.offsets[a.n_owned] > a.capacity) return;  // overflow already flagged by the offsets scan
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  int32_t* line = er_smem + (warp * 32 + lane) * EM_LINE;
  int32_t* win = er_smem + ER_WARPS * 32 * EM_LINE + warp * EW_WIN;
  const int32_t slot = (blockIdx.x * ER_WARPS + warp) * 32 + lane;
  int32_t id = 0x7fffffff;
  if (slot < a.n_total && slot < __ldg(a.cell_start + a.n_cells)) id = __ldg(a.sorted_ids + slot);
  const bool owned = id < a.n_owned;
  if (!__any_sync(0xffffffffu, owned)) return;
  const int32_t cell = owned ? __ldg(a.slot_cell + slot) : 0;
  const long long dst = owned ? (long long)a.offsets[id] : 0;
  int32_t fill = 0;  // entries staged in this lane's line
  int32_t done = 0;  // entries of this row already written

  // Every lane copies ITS OWN line (final ids) to its row: scalar stores until the row position is 32-byte aligned,
  // then one STG.256 per 8 entries; what does not fill a vector stays in the line.
  auto flush = [&](bool final) {
    int32_t k = 0;
    int32_t* out = a.partners + dst + done;
    while (k < fill && ((reinterpret_cast<uintptr_t>(out + k) & 31) != 0)) {
      out[k] = line[k];
      k++;
    }
    while (k + 8 <= fill) {
      int32_t v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) v[u] = line[k + u];
      asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(out + k), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                   "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                   : "memory");
      k += 8;
    }
    if (final) {
      while (k < fill) {
        out[k] = line[k];
        k++;
      }
    }
    done += k;
    const int32_t left = fill - k;
    for (int32_t t = 0; t < left; t++) line[t] = line[k + t];
    fill = left;
  };

  const int32_t mx = a.mesh[0], my = a.mesh[1], mz = a.mesh[2];
  const int32_t byz = (int32_t)fdiv((uint32_t)cell, a.d_mx), bx = cell - byz * mx;
  const int32_t bz = (int32_t)fdiv((uint32_t)byz, a.d_my), by = byz - bz * my;
  int xlo, xhi, ylo, yhi, zlo, zhi;
  axis_range(bx, mx, xlo, xhi);
  axis_range(by, my, ylo, yhi);
  axis_range(bz, mz, zlo, zhi);
  const int32_t ny = yhi - ylo + 1, nz = zhi - zlo + 1;
  const uint32_t* mrow = a.mask + min((long long)slot, a.n_cap - 1);  // lanes past the last slot load in range
  const long long run_stride = (long long)a.wr * a.n_cap;  // words between the planes of two runs
  const int32_t pre = min(a.wr, ER_PRE);

  // Bit position p of a word <-> the slot `last - p` (last = the slot of bit 0).  In the window: the id sits at the
  // shared-memory byte address last_sa - 4 p.  Two entries per trip, the second predicated; both id loads are issued
  // before the first store into the line, so a trip waits for shared memory once.  (Four entries per trip: 88 bytes
  // of spills at 72 registers, 2^24 uniform particles 8.27 -> 9.12 ms.)
  const uint32_t line_sa = (uint32_t)__cvta_generic_to_shared(line);
  const uint32_t win_sa = (uint32_t)__cvta_generic_to_shared(win);
  auto expand_win = [&](uint32_t word, uint32_t last_sa) {
    uint32_t wa = line_sa + 4u * (uint32_t)fill;
    fill += __popc(word);
    if (word) {
      asm volatile(
          "{\n\t"
          ".reg .pred p, q;\n\t"
          ".reg .u32 pos, bit, ad, v0, v1;\n"
          "EXPANDW_%=:\n\t"
          "bfind.u32 pos, %0;\n\t"
          "shl.b32 bit, 1, pos;\n\t"
          "mad.lo.s32 ad, pos, -4, %2;\n\t"
          "xor.b32 %0, %0, bit;\n\t"
          "ld.shared.s32 v0, [ad];\n\t"
          "setp.ne.u32 p, %0, 0;\n\t"
          "@p bfind.u32 pos, %0;\n\t"
          "@p shl.b32 bit, 1, pos;\n\t"
          "@p mad.lo.s32 ad, pos, -4, %2;\n\t"
          "@p xor.b32 %0, %0, bit;\n\t"
          "@p ld.shared.s32 v1, [ad];\n\t"
          "st.shared.s32 [%1], v0;\n\t"
          "@p st.shared.s32 [%1+4], v1;\n\t"
          "add.u32 %1, %1, 8;\n\t"
          "setp.ne.u32 q, %0, 0;\n\t"
          "@q bra EXPANDW_%=;\n\t"
          "}"
          : "+r"(word), "+r"(wa)
          : "r"(last_sa)
          : "memory");
    }
  };
  // the same with the id read from global memory (warps without a window)
  auto expand_glb = [&](uint32_t word, const int32_t* last) {
    int32_t* wp = line + fill;
    fill += __popc(word);
    while (word) {
      const int p = 31 - __clz(word);
      word ^= 1u << p;
      *wp++ = __ldg(last - p);
    }
  };
  auto expand = [&](uint32_t word, const int32_t* glb_last, uint32_t sa_last, bool in_smem) {
    if (in_smem)
      expand_win(word, sa_last);
    else
      expand_glb(word, glb_last);
  };
  int32_t s0n, s1n;
  uint32_t mn[ER_PRE];
  const uint32_t* mrun = mrow;
  auto request = [&](int r) {
    const int oz = r / 3, oy = r - oz * 3;
    const int32_t* cs = a.cell_start + (min(ylo + oy, my - 1) + min(zlo + oz, mz - 1) * my) * mx;
    s0n = __ldg(cs + xlo);
    s1n = __ldg(cs + xhi + 1);
#pragma unroll
    for (int u = 0; u < ER_PRE; u++) mn[u] = __ldg(mrun + (long long)min(u, pre - 1) * a.n_cap);
    mrun += run_stride;
  };
  request(0);
#pragma unroll 1
  for (int r = 0; r < 9; r++) {
    const int oz = r / 3, oy = r - oz * 3;
    const int32_t s0 = s0n, s1 = s1n;
    const bool rv = owned && (oz < nz) && (oy < ny);
    const int32_t nw = rv ? min((s1 - s0 + 31) >> 5, a.wr) : 0;  // > wr only after FLAG_CELL_WORDS
    uint32_t m[ER_PRE];
    int32_t tot = 0;
#pragma unroll
    for (int u = 0; u < ER_PRE; u++) {
      m[u] = u < nw ? mn[u] : 0u;
      tot += __popc(m[u]);
    }
    if (r < 8) request(r + 1);
    const int32_t nwmax = __reduce_max_sync(0xffffffffu, nw);
    if (nwmax == 0) continue;
    // the window: every slot a row of this warp can name in this run
    const int32_t w0 = __reduce_min_sync(0xffffffffu, nw > 0 ? s0 : 0x7fffffff);
    const int32_t w1 = __reduce_max_sync(0xffffffffu, nw > 0 ? min(s1, s0 + 32 * nw) : 0);
    const bool in_smem = w1 - w0 <= EW_WIN;
    if (in_smem) {
      __syncwarp();  // the previous run's readers are done
      for (int32_t k = lane; k < w1 - w0; k += 32) win[k] = __ldg(a.slot_pid + w0 + k);
      __syncwarp();
    }
    const int32_t* src = a.slot_pid + s0 + 31;
    const uint32_t ssa = win_sa + 4u * (uint32_t)(s0 - w0 + 31);
    if (__any_sync(0xffffffffu, fill + tot > EM_TILE)) flush(false);  // leaves fill <= 7
    if (!__any_sync(0xffffffffu, tot > EM_TILE - 7 || nw > ER_PRE)) {
      // the run fits the line: no check between its words
#pragma unroll
      for (int u = 0; u < ER_PRE; u++)
        if (u < nwmax) expand(m[u], src + 32 * u, ssa + 128u * u, in_smem);
    } else {
      // a long run (the row's own run can hold more partners than a line, crowded cells more than ER_PRE words):
      // word by word with a check before each — the requested words from their registers, the rest on demand
#pragma unroll
      for (int u = 0; u < ER_PRE; u++) {
        if (u < nwmax && __any_sync(0xffffffffu, m[u] != 0u)) {
          if (__any_sync(0xffffffffu, fill + __popc(m[u]) > EM_TILE)) flush(false);
          expand(m[u], src + 32 * u, ssa + 128u * u, in_smem);
        }
      }
      for (int32_t w = ER_PRE; w < nwmax; w++) {
        const uint32_t word = w < nw ? __ldg(mrow + ((long long)(r * a.wr + w)) * a.n_cap) : 0u;
        if (!__any_sync(0xffffffffu, word != 0u)) continue;
        if (__any_sync(0xffffffffu, fill + __popc(word) > EM_TILE)) flush(false);
        expand(word, src + 32 * w, ssa + 128u * (uint32_t)w, in_smem);
      }
    }
  }
  flush(true);
}

}  // namespace nlb
