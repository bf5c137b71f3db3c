// nlist_rowmask.cuh — the round-2 search + emission pair ("row masks").
//
//   rowmask_kernel   one CTA per cell A (persistent CTAs, dynamic queue).  The candidates of A — the <= 9 contiguous
//                    x-runs of its stencil in the cell-sorted record array — are pulled into shared memory by the
//                    TMA engine (cp.async.bulk, one copy per run, completion on an mbarrier), two cells ahead of the
//                    one being computed: the runs land back to back, so the staged window IS A's candidate list and
//                    candidate c is simply win[c].  Records hold ABSOLUTE FP32 coordinates; a unit shifts them into
//                    the frame of A's centre with three subtractions per candidate, nothing is looked up per cell.
//                    Test as in round 1: candidates on the lanes (RJ per lane, packed in pairs), the rows i of A
//                    broadcast from shared memory, dot form d = xi.xj - |xj|^2/2 - (|xi|^2 - SL^2)/2 with
//                    3 FFMA2 + 1 FADD2 per two tests, sign bit funnel-shifted into a word, min|d| tracked, exact
//                    input-precision re-test inside the band E.
//                    New: the 32 x 32 verdict block a warp holds after 32 rows (lane = candidate, bit = row) is
//                    TRANSPOSED in registers (5 butterfly stages of SHFL + SHF + LOP3), so that lane = row i holds
//                    "which of these 32 candidates are partners of i".  The row's popcount accumulates in shared
//                    memory — counts[id] leaves the kernel with it, the separate popcount pass of round 1 is gone —
//                    and the words are stored row-major per cell: mask[base_A + k * n_A + i], k = 32-candidate block
//                    of A's list: dense (K = ceil(nj / 32) ~ 30 words per row instead of 27 * 3 = 81), coalesced for
//                    this writer (lanes = rows) and for the reader (lanes = rows).  base_A comes from one atomicAdd
//                    per cell on a cursor: the mask holds exactly one bit per evaluated test (+ padding), whatever
//                    the density profile.  Per cell a 88-byte CellRec tells the emission where the 9 runs start.
//   emit3_kernel     thread = row, warp = 32 consecutive cell-sorted slots.  Expands its row's words (coalesced loads)
//                    into a per-lane shared-memory line of candidate SLOTS (candidate index + run delta, run cursor
//                    in registers), flushes lines with 16-byte stores as round 1 did.
//
// Cells of more than RM_ROWCAP rows are walked in row rounds, windows larger than the staged capacity in window
// rounds (fetched synchronously): clustered inputs take the same code path, only slower per byte.
#pragma once

#include "nlist_kernels.cuh"

namespace nlb {

constexpr int RM_THREADS = 128;
constexpr int RM_WARPS = RM_THREADS / 32;
constexpr int RM_ROWCAP = 128;  // rows of a cell staged per round
#ifndef NLB_RM_MINB
#define NLB_RM_MINB 4
#endif

struct RmDesc {
  int32_t cell, n_a, slot_a0, nj, self_base, nruns;
  unsigned long long mask_base;
  float ox, oy, oz;
  int32_t store;  // 0: the mask buffer is full, compute counts only (the build fails with FLAG_MASK_WORDS)
  int32_t s0[9], cs[9], ce[9];  // first slot, start and end in the candidate list of run r
  int32_t pad;
};
static_assert(sizeof(RmDesc) % 8 == 0, "RmDesc must keep the mbarriers behind it 8-byte aligned");

template <typename T>
struct RowMaskArgs {
  const T* q;  // caller's positions (band re-test only)
  GridParams<T> gp;
  const int32_t* cell_start;
  const float4* rec;          // cell-sorted records: absolute FP32 coordinates, .w = local id
  const int32_t* global_ids;  // HALFMODE 2: local -> global id map
  int32_t n_owned;
  uint32_t* mask;
  unsigned long long mask_cap;  // words
  CellRec* cellrec;
  int32_t* counts;
  FastDiv d_mx, d_my;
  float band;
  int32_t win_cap;  // candidates staged per window round (multiple of 256)
  unsigned long long* queue;
  DeviceStatus* st;
};

__host__ __device__ inline size_t rm_smem_bytes(int win_cap) {
  return (size_t)2 * win_cap * sizeof(float4) + (size_t)2 * RM_ROWCAP * sizeof(float4) +
         (size_t)RM_ROWCAP * 2 * sizeof(float4) + (size_t)3 * RM_ROWCAP * sizeof(int32_t) + 2 * sizeof(RmDesc) +
         2 * sizeof(unsigned long long);
}

// ---- mbarrier / bulk-copy primitives (PTX; CUDA 12.9) ------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, uint32_t parity) {
  const uint32_t addr = smem_u32(b);
  uint32_t ok;
#ifdef NLB_WAIT_GUARD
  uint32_t spins = 0;
#endif
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
#ifdef NLB_WAIT_GUARD
    if (!ok && ++spins > (1u << 24)) __trap();  // bring-up: a lost completion must fail the launch, not hang the GPU
#endif
  } while (!ok);
}
// global -> shared bulk copy by the TMA engine; bytes a positive multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(b))
               : "memory");
}

// One fetch = the window pieces [c_lo, c_hi) of the candidate list (lane r < 9 copies its run's part) and the raw
// records of rows [r_lo, r_hi) (lane 9).  Executed by one full warp; completion on `bar`.
__device__ __forceinline__ void rm_fetch(const float4* __restrict__ rec, float4* win, float4* rowraw,
                                         unsigned long long* bar, int lane, int32_t s0, int32_t cs, int32_t ce,
                                         int32_t c_lo, int32_t c_hi, int32_t slot_a0, int32_t r_lo, int32_t r_hi) {
  uint32_t bytes = 0;
  const float4* src = rec;
  float4* dst = win;
  if (lane < 9) {
    const int32_t lo = max(c_lo, cs), hi = min(c_hi, ce);
    if (hi > lo) {
      bytes = (uint32_t)(hi - lo) * 16u;
      src = rec + s0 + (lo - cs);
      dst = win + (lo - c_lo);
    }
  } else if (lane == 9 && r_hi > r_lo) {
    bytes = (uint32_t)(r_hi - r_lo) * 16u;
    src = rec + slot_a0 + r_lo;
    dst = rowraw;
  }
  uint32_t total = bytes;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) total += __shfl_xor_sync(0xffffffffu, total, d);
  if (lane == 0) mbar_expect_tx(bar, total);
  __syncwarp();
  if (bytes) bulk_g2s(dst, src, bytes, bar);
}

// 32 x 32 bit-matrix transpose across the lanes of a warp: in  lane c holds word H[c], out lane r holds T[r] with
// bit c of T[r] = bit r of H[c].  Five butterfly stages; per stage one SHFL, one rotate (SHF) and one bit-select (LOP3).
struct TransposeConsts {
  uint32_t keep[5], rot[5];
};
__device__ __forceinline__ TransposeConsts make_transpose_consts(int lane) {
  TransposeConsts t;
  const uint32_t m[5] = {0x0000ffffu, 0x00ff00ffu, 0x0f0f0f0fu, 0x33333333u, 0x55555555u};
#pragma unroll
  for (int i = 0; i < 5; i++) {
    const int s = 16 >> i;
    const bool up = (lane & s) != 0;
    t.keep[i] = up ? ~m[i] : m[i];
    t.rot[i] = up ? (uint32_t)(32 - s) : (uint32_t)s;
  }
  return t;
}
__device__ __forceinline__ uint32_t transpose32(uint32_t x, const TransposeConsts& t) {
#pragma unroll
  for (int i = 0; i < 5; i++) {
    const uint32_t y = __shfl_xor_sync(0xffffffffu, x, 16 >> i);
    const uint32_t r = __funnelshift_l(y, y, t.rot[i]);
    x = (x & t.keep[i]) | (r & ~t.keep[i]);
  }
  return x;
}

// HALFMODE: 0 FULL (every j != i), 1 HALF by local id (ids ascend with the slot inside a cell: prefix cut),
//           2 HALF by global id (multi-GPU: per-bit comparison)
template <typename T, int STRIDE, int HALFMODE, int RJ>
__global__ void __launch_bounds__(RM_THREADS, NLB_RM_MINB) rowmask_kernel(RowMaskArgs<T> a) {
  pdl_enter();
  extern __shared__ __align__(128) unsigned char rm_smem[];
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int32_t WC = a.win_cap;
  float4* win0 = reinterpret_cast<float4*>(rm_smem);               // [2][WC]
  float4* rowraw0 = win0 + 2 * (size_t)WC;                         // [2][RM_ROWCAP]
  float4* srow = rowraw0 + 2 * RM_ROWCAP;                          // [RM_ROWCAP][2]: {xi,xi,yi,yi},{zi,zi,-ai,-ai}
  int32_t* sid = reinterpret_cast<int32_t*>(srow + 2 * RM_ROWCAP);  // local id of the staged rows
  int32_t* scmp = sid + RM_ROWCAP;                                  // id the HALF rule compares
  int32_t* scnt = scmp + RM_ROWCAP;                                 // row lengths of the round
  RmDesc* desc = reinterpret_cast<RmDesc*>(scnt + RM_ROWCAP);       // [2]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(desc + 2);  // [2]

  const GridParams<T>& gp = a.gp;
  const int32_t mx = gp.mesh[0], my = gp.mesh[1], mz = gp.mesh[2], M = gp.n_cells;
  const float msx = gp.msf[0], msy = gp.msf[1], msz = gp.msf[2];
  const TransposeConsts tc = make_transpose_consts(lane);
  unsigned long long band_local = 0, cand_local = 0;

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int r = threadIdx.x; r < RM_ROWCAP; r += RM_THREADS) scnt[r] = 0;
  __syncthreads();

  // ---- the prefetch pipeline (warp 0).  Position n of this CTA's cell sequence is computed while the window of
  //      position n+1 is in flight and the cell_start values of position n+2 are being loaded; the cell index of
  //      position n+3 comes from the queue.  The first three positions are static. ----
  int32_t c_cur = (int32_t)blockIdx.x;                      // position n
  int32_t c_nx1 = c_cur < M - (int32_t)gridDim.x ? c_cur + (int32_t)gridDim.x : M;  // n+1
  int32_t c_nx2 = c_nx1 < M - (int32_t)gridDim.x ? c_nx1 + (int32_t)gridDim.x : M;  // n+2
  unsigned long long q_pending = 0;  // queue ticket for position n+3 (lane 0 of warp 0; in flight)
  int32_t pf_s0 = 0, pf_s1 = 0;      // raw cell_start values of position n+1 (lane r < 9: run r; lane 9: own cell)
  int32_t pf2_s0 = 0, pf2_s1 = 0;    // ... of position n+2 (in flight)
  unsigned long long mb_pending = 0;  // mask base of position n+1 (lane 0; atomic in flight)
  uint32_t phases = 0u;  // bit b: parity the next wait on bars[b] expects

  // cell -> (cx, cy, cz), stencil ranges; lane r < nruns loads the bounds of run r = (z, y), lane 9 the own cell
  auto issue_loads = [&](int32_t cell, int32_t& s0, int32_t& s1) {
    s0 = 0;
    s1 = 0;
    if (cell >= M) return;
    const int32_t cyz = (int32_t)fdiv((uint32_t)cell, a.d_mx), cx = cell - cyz * mx;
    const int32_t cz = (int32_t)fdiv((uint32_t)cyz, a.d_my), cy = cyz - cz * my;
    int xlo, xhi, ylo, yhi, zlo, zhi;
    axis_range(cx, mx, xlo, xhi);
    axis_range(cy, my, ylo, yhi);
    axis_range(cz, mz, zlo, zhi);
    const int32_t ny = yhi - ylo + 1, nruns = ny * (zhi - zlo + 1);
    if (lane < nruns) {
      const int lz = ny == 3 ? (lane * 11) >> 5 : (ny == 2 ? lane >> 1 : lane);  // lane / ny for lane < 9
      const int z = zlo + lz, y = ylo + lane - lz * ny;
      const int32_t* cs = a.cell_start + (y + z * my) * mx;
      s0 = __ldg(cs + xlo);
      s1 = __ldg(cs + xhi + 1);
    } else if (lane == 9) {
      s0 = __ldg(a.cell_start + cell);
      s1 = __ldg(a.cell_start + cell + 1);
    }
  };
  // builds the descriptor of `cell` in desc[b] from its loaded cell_start values, writes its CellRec, requests its
  // mask block and starts the fetch of its first window round + first row round
  auto make_unit = [&](int32_t cell, int b, int32_t s0, int32_t s1) {
    RmDesc& d = desc[b];
    if (cell >= M) {
      if (lane == 0) {
        d.cell = M;
        d.n_a = 0;
      }
      return;
    }
    const int32_t cyz = (int32_t)fdiv((uint32_t)cell, a.d_mx), cx = cell - cyz * mx;
    const int32_t cz = (int32_t)fdiv((uint32_t)cyz, a.d_my), cy = cyz - cz * my;
    int xlo, xhi, ylo, yhi, zlo, zhi;
    axis_range(cx, mx, xlo, xhi);
    axis_range(cy, my, ylo, yhi);
    axis_range(cz, mz, zlo, zhi);
    const int32_t ny = yhi - ylo + 1, nruns = ny * (zhi - zlo + 1);
    const int32_t len = lane < nruns ? s1 - s0 : 0;
    int32_t incl = len;
#pragma unroll
    for (int dd = 1; dd < 16; dd <<= 1) {
      const int32_t v = __shfl_up_sync(0xffffffffu, incl, dd);
      if (lane >= dd) incl += v;
    }
    const int32_t nj = __shfl_sync(0xffffffffu, incl, 8);
    const int32_t cs = incl - len;
    const int32_t a0 = __shfl_sync(0xffffffffu, s0, 9), a1 = __shfl_sync(0xffffffffu, s1, 9);
    const int32_t n_a = a1 - a0;
    const int r_own = (cz - zlo) * ny + (cy - ylo);
    const int32_t self_base = __shfl_sync(0xffffffffu, cs, r_own) + (a0 - __shfl_sync(0xffffffffu, s0, r_own));
    if (lane < 9) {
      d.s0[lane] = s0;
      d.cs[lane] = cs;
      d.ce[lane] = incl;
    }
    if (lane == 0) {
      d.cell = cell;
      d.n_a = n_a;
      d.slot_a0 = a0;
      d.nj = nj;
      d.self_base = self_base;
      d.nruns = nruns;
      d.ox = ((float)(cx + gp.coff[0]) + 0.5f) * msx;
      d.oy = ((float)(cy + gp.coff[1]) + 0.5f) * msy;
      d.oz = ((float)(cz + gp.coff[2]) + 0.5f) * msz;
    }
    if (n_a == 0) return;
    CellRec& cr = a.cellrec[cell];
    if (lane < 9) cr.run[lane] = make_int2(incl, s0 - cs);
    if (lane == 0) {
      cr.nj = nj;
      cr.self_base = self_base;
      const unsigned long long need = (unsigned long long)n_a * (unsigned long long)((nj + 31) >> 5);
      mb_pending = atomicAdd(&a.st->mask_words, need);
      cand_local += (unsigned long long)n_a * (unsigned long long)nj;
    }
    rm_fetch(a.rec, win0 + (size_t)b * WC, rowraw0 + b * RM_ROWCAP, &bars[b], lane, s0, cs, incl, 0, min(nj, WC), a0,
             0, min(n_a, RM_ROWCAP));
  };
  // lane 0: the mask base requested by make_unit has arrived — publish it (descriptor + CellRec)
  auto publish_base = [&](int b) {
    RmDesc& d = desc[b];
    if (lane == 0 && d.cell < M && d.n_a > 0) {
      const unsigned long long need = (unsigned long long)d.n_a * (unsigned long long)((d.nj + 31) >> 5);
      const bool fits = mb_pending + need <= a.mask_cap;
      if (!fits) atomicOr(&a.st->flags, FLAG_MASK_WORDS);
      d.mask_base = mb_pending;
      d.store = fits ? 1 : 0;
      a.cellrec[d.cell].mask_base = mb_pending;
    }
  };

  if (warp == 0) {
    // prologue: position 0 synchronously, position 1's loads in flight
    int32_t s0, s1;
    issue_loads(c_cur, s0, s1);
    make_unit(c_cur, 0, s0, s1);
    publish_base(0);
    issue_loads(c_nx1, pf_s0, pf_s1);
    if (lane == 0) q_pending = atomicAdd(a.queue, 1ull);
  }
  __syncthreads();

  for (int n = 0; c_cur < M; n++) {
    const int b = n & 1;
    if (warp == 0) {
      // position n+1: descriptor, CellRec, mask request, TMA fetch into the other buffer (its previous user,
      // position n-1, was finished by every warp before the barrier that ended iteration n-1)
      make_unit(c_nx1, b ^ 1, pf_s0, pf_s1);
      // position n+2: loads in flight until the next iteration; position n+3: cell index from the queue
      issue_loads(c_nx2, pf2_s0, pf2_s1);
    }
    const RmDesc& d = desc[b];
    const int32_t n_a = d.n_a;
    if (n_a > 0) {
      const int32_t nj = d.nj, self_base = d.self_base, slot_a0 = d.slot_a0;
      const float ox = d.ox, oy = d.oy, oz = d.oz;
      float4* win = win0 + (size_t)b * WC;
      float4* rowraw = rowraw0 + b * RM_ROWCAP;
      const bool store = d.store != 0;
      uint32_t* mcell = a.mask + d.mask_base;
      bool first_fetch = true;  // the prefetched window round 0 + row round 0 are (still) in the buffers
      for (int32_t rr = 0; rr < n_a; rr += RM_ROWCAP) {
        const int32_t nrows = min(RM_ROWCAP, n_a - rr);
        if (rr > 0) {
          // later row rounds of a crowded cell: rows + window round 0 fetched synchronously
          __syncthreads();
          if (warp == 0)
            rm_fetch(a.rec, win, rowraw, &bars[b], lane, lane < 9 ? d.s0[lane] : 0, lane < 9 ? d.cs[lane] : 0,
                     lane < 9 ? d.ce[lane] : 0, 0, min(nj, WC), slot_a0, rr, rr + nrows);
          first_fetch = true;
        }
        for (int32_t sc0 = 0; sc0 < nj; sc0 += WC) {
          const int32_t ncand = min(WC, nj - sc0);
          if (!first_fetch) {
            __syncthreads();  // every warp is done with the previous window round
            if (warp == 0)
              rm_fetch(a.rec, win, rowraw, &bars[b], lane, lane < 9 ? d.s0[lane] : 0, lane < 9 ? d.cs[lane] : 0,
                       lane < 9 ? d.ce[lane] : 0, sc0, sc0 + ncand, slot_a0, 0, 0);
          }
          mbar_wait(&bars[b], (phases >> b) & 1u);
          phases ^= 1u << b;
          if (sc0 == 0) {
            // stage the round's rows: frame of A's centre, pre-duplicated for the packed FMAs
            for (int32_t r = threadIdx.x; r < nrows; r += RM_THREADS) {
              const float4 v = rowraw[r];
              const float x = v.x - ox, y = v.y - oy, z = v.z - oz;
              const float nai = -0.5f * (fmaf(x, x, fmaf(y, y, z * z)) - gp.sl2f);
              srow[2 * r] = make_float4(x, x, y, y);
              srow[2 * r + 1] = make_float4(z, z, nai, nai);
              const int32_t id = __float_as_int(v.w);
              sid[r] = id;
              if (HALFMODE == 2) scmp[r] = __ldg(a.global_ids + id);
              if (HALFMODE == 1) scmp[r] = id;
            }
          }
          first_fetch = false;
          __syncthreads();

          // ---- this warp's candidate chunks of the window round ----
          for (int32_t c0 = warp * (32 * RJ); c0 < ncand; c0 += RM_WARPS * 32 * RJ) {
            float xj[RJ], yj[RJ], zj[RJ], wj[RJ];
            int32_t cj[RJ];   // HALF: id the rule compares for candidate k
            int32_t cut[RJ];  // HALFMODE 1: rows of the round with an id below the candidate's (a prefix)
#pragma unroll
            for (int k = 0; k < RJ; k++) {
              const int32_t c = c0 + k * 32 + lane;
              const float4 v = win[min(c, ncand - 1)];
              xj[k] = v.x - ox;
              yj[k] = v.y - oy;
              zj[k] = v.z - oz;
              wj[k] = -0.5f * fmaf(xj[k], xj[k], fmaf(yj[k], yj[k], zj[k] * zj[k]));
              cj[k] = __float_as_int(v.w);
              if (c >= ncand) {
                xj[k] = yj[k] = zj[k] = 0.f;
                wj[k] = -1.0e30f;  // d = -1e30: a miss, far from the band
                cj[k] = 0x80000000;
              } else if (HALFMODE == 2) {
                cj[k] = __ldg(a.global_ids + cj[k]);
              }
              cut[k] = 0;
            }
            if (HALFMODE == 1) {
              // rows keep partners with a LARGER id (neighlist_cpu.hpp:225-236): seen from candidate j, the rows
              // with id_i < id_j.  Ids ascend with the slot inside a cell, so these rows are a prefix of the round:
              // lower bound over the staged ids, branch-free, the RJ searches interleaved
#pragma unroll
              for (int k = 0; k < RJ; k++) cut[k] = 0;
              for (int32_t step = 1 << (31 - __clz(nrows)); step > 0; step >>= 1) {
#pragma unroll
                for (int k = 0; k < RJ; k++) {
                  const int32_t t = cut[k] + step;
                  if (t <= nrows && scmp[min(t, nrows) - 1] < cj[k]) cut[k] = t;
                }
              }
            }
            f32x2 X[RJ / 2], Y[RJ / 2], Z[RJ / 2], W[RJ / 2];
#pragma unroll
            for (int h = 0; h < RJ / 2; h++) {
              X[h] = pack2(xj[2 * h], xj[2 * h + 1]);
              Y[h] = pack2(yj[2 * h], yj[2 * h + 1]);
              Z[h] = pack2(zj[2 * h], zj[2 * h + 1]);
              W[h] = pack2(wj[2 * h], wj[2 * h + 1]);
            }
            const int32_t nblk = min(RJ, (ncand - c0 + 31) >> 5);  // blocks of this chunk that hold candidates
            const int32_t kblk0 = (sc0 + c0) >> 5;                  // their index in the cell's list
            for (int32_t w = 0; w * 32 < nrows; w++) {
              const int32_t cnt = min(32, nrows - w * 32);
              const ulonglong2* sp = reinterpret_cast<const ulonglong2*>(srow + w * 64);
              // rows are walked downwards so that row ii ends at bit ii; rows >= cnt keep the initial ones (= miss)
              uint32_t miss[RJ];
#pragma unroll
              for (int k = 0; k < RJ; k++) miss[k] = 0xffffffffu;
              float mh[RJ / 2];  // min |d| per candidate pair
#pragma unroll
              for (int h = 0; h < RJ / 2; h++) mh[h] = 3.0e38f;
#pragma unroll 2
              for (int32_t ii = cnt - 1; ii >= 0; ii--) {
                const ulonglong2 p0 = sp[2 * ii];      // {xi, xi}, {yi, yi}
                const ulonglong2 p1 = sp[2 * ii + 1];  // {zi, zi}, {-ai, -ai}
#pragma unroll
                for (int h = 0; h < RJ / 2; h++) {
                  const f32x2 d2 = add2(fma2(p0.x, X[h], fma2(p0.y, Y[h], fma2(p1.x, Z[h], W[h]))), p1.y);
                  float d0, d1;
                  unpack2(d2, d0, d1);
                  miss[2 * h] = __funnelshift_l(__float_as_uint(d0), miss[2 * h], 1);  // shift the sign bit in
                  miss[2 * h + 1] = __funnelshift_l(__float_as_uint(d1), miss[2 * h + 1], 1);
                  mh[h] = fminf(mh[h], fminf(fabsf(d0), fabsf(d1)));
                }
              }
              uint32_t hits[RJ];
#pragma unroll
              for (int k = 0; k < RJ; k++) hits[k] = ~miss[k];  // bit ii <-> row w*32 + ii
              // tests inside the pre-filter's uncertainty band are decided exactly, in the caller's precision, by the
              // whole warp (lane = row, the triggering lane's candidate broadcast)
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
                    const int32_t c_src = c0 + k * 32 + src;
                    bool fix = false, hit = false;
                    if (lane < cnt && c_src < ncand) {
                      const float4 q0 = srow[w * 64 + 2 * lane], q1 = srow[w * 64 + 2 * lane + 1];
                      const float dd = pre_d(q0.x, q0.z, q1.x, q1.z, cx2[e], cy2[e], cz2[e], cw2[e]);
                      if (fabsf(dd) < a.band) {
                        const int32_t iid = sid[w * 32 + lane];
                        const int32_t jid = __float_as_int(win[c_src].w);
                        hit = exact_within(load_pos<T, STRIDE>(a.q, iid), load_pos<T, STRIDE>(a.q, jid), gp.sl2);
                        fix = true;
                        band_local++;
                      }
                    }
                    const uint32_t fixm = __ballot_sync(0xffffffffu, fix);  // lane ii <-> bit ii
                    const uint32_t hitm = __ballot_sync(0xffffffffu, hit);
                    if (lane == src) hits[k] = (hits[k] & ~fixm) | hitm;
                  }
                }
              }
              if (HALFMODE == 1) {
#pragma unroll
                for (int k = 0; k < RJ; k++) {
                  const int32_t keep = cut[k] - w * 32;  // rows w*32 .. w*32 + keep - 1 have an id below the candidate's
                  hits[k] = keep <= 0 ? 0u : (keep >= 32 ? hits[k] : (hits[k] & ((1u << keep) - 1u)));
                }
              }
              if (HALFMODE == 2) {
                // ids are not monotone in the slot when they come from a map: compare per set bit
#pragma unroll
                for (int k = 0; k < RJ; k++) {
                  uint32_t m = hits[k];
                  while (m) {
                    const int bpos = __ffs(m) - 1;
                    m &= m - 1;
                    if (!(cj[k] > scmp[w * 32 + bpos])) hits[k] &= ~(1u << bpos);
                  }
                }
              }
              // lane = candidate, bit = row   ->   lane = row, bit = candidate
              const int32_t rg = rr + w * 32 + lane;  // this lane's row inside the cell
              const int32_t cself = self_base + rg;   // ... and its own position in the candidate list (FULL: j != i)
              const bool owned = lane < cnt && sid[w * 32 + lane] < a.n_owned;
              int32_t pc = 0;
#pragma unroll
              for (int k = 0; k < RJ; k++) {
                if (k < nblk) {
                  uint32_t t = transpose32(hits[k], tc);
                  if (HALFMODE == 0 && (cself >> 5) == kblk0 + k) t &= ~(1u << (cself & 31));
                  pc += __popc(t);
                  if (owned && store) mcell[(size_t)(kblk0 + k) * (size_t)n_a + (size_t)rg] = t;
                }
              }
              if (owned && pc) atomicAdd(&scnt[w * 32 + lane], pc);
            }
          }
        }
        // ---- the round's row lengths leave with the kernel: no separate popcount pass ----
        __syncthreads();
        for (int32_t r = threadIdx.x; r < nrows; r += RM_THREADS) {
          const int32_t id = sid[r];
          if (id < a.n_owned) a.counts[id] = scnt[r];
          scnt[r] = 0;
        }
      }
    }
    if (warp == 0) {
      // the mask base of position n+1 (requested at the top of this iteration) is consumed only now
      publish_base(b ^ 1);
      unsigned long long t = __shfl_sync(0xffffffffu, q_pending, 0);
      if (lane == 0) q_pending = atomicAdd(a.queue, 1ull);  // position n+4's ticket
      const long long nxt = 3ll * (long long)gridDim.x + (long long)t;
      c_cur = c_nx1;
      c_nx1 = c_nx2;
      c_nx2 = nxt < (long long)M ? (int32_t)nxt : M;
      pf_s0 = pf2_s0;
      pf_s1 = pf2_s1;
    }
    __syncthreads();  // desc[b ^ 1] complete, every warp done with buffer b
    // every warp follows warp 0's sequence: the next cell index travels through the descriptor
    c_cur = desc[b ^ 1].cell;
  }
  if (warp == 0 && lane == 0) {
    if (cand_local) atomicAdd(&a.st->candidates, cand_local);
  }
  if (band_local) atomicAdd(&a.st->band_tests, band_local);
}

// ---------------------------------------------------------------------------------------------------------------
// rowmask4_kernel — the same search with WARP-AUTONOMOUS units.
// Profile of rowmask_kernel (profiles/r02_ncu_rowmask_v3a.txt): the main loop is 54 % of the instructions but 20 % of
// the stall samples; 36 % of the samples sit at the three CTA barriers a cell costs (rows staged -> chunks done ->
// counts written), with the four warps of a CTA in lock step per cell and warp 0 alone feeding the pipeline.
// Here a unit is (cell, part): ONE warp takes the part's 32 * RJ candidates of the cell's list against all rows of the
// cell and shares nothing with other warps — no CTA barrier exists.  Every warp runs its own three-deep pipeline:
// ticket of the unit three ahead (atomic in flight), CellRec + cell bounds of the unit two ahead (loads in flight),
// window piece + rows of the next unit by TMA into the warp's second buffer (mbarrier per buffer), compute the current
// one.  The CellRec (run table, list length, mask block) is written by cellsort_kernel, which already has a warp per
// cell, so the units of a cell agree on the block without talking to each other; row lengths accumulate with global
// atomics (zeroed by cellsort_kernel).  Cells with more candidates than UPC parts cover, or more rows than RM4_ROWCAP,
// loop inside the unit (later chunks / row rounds fetched synchronously).
// ---------------------------------------------------------------------------------------------------------------
constexpr int RM4_ROWCAP = 64;  // rows staged per round
struct Rm4Desc {
  int32_t n_a, slot_a0, nj, self_base, part, pad0;
  unsigned long long mask_base;
  float ox, oy, oz;
  int32_t pad1;
  int32_t s0[9], cs[9], ce[9];  // first slot, start and end in the candidate list of run r
  int32_t pad2;
};
template <int RJ>
struct Rm4Smem {
  float4 win[2][32 * RJ];
  float4 rowraw[2][RM4_ROWCAP];
  float4 srow[RM4_ROWCAP][2];
  int32_t sid[RM4_ROWCAP];
  int32_t scmp[RM4_ROWCAP];
  Rm4Desc desc[2];
  unsigned long long bars[2];
};

template <typename T>
struct RowMask4Args {
  const T* q;  // caller's positions (band re-test only)
  GridParams<T> gp;
  const int32_t* cell_start;
  const float4* rec;          // cell-sorted records: absolute FP32 coordinates, .w = local id
  const int32_t* global_ids;  // HALFMODE 2: local -> global id map
  int32_t n_owned;
  uint32_t* mask;
  unsigned long long mask_cap;  // words
  const CellRec* cellrec;
  int32_t* counts;  // zeroed by cellsort_kernel
  FastDiv d_mx, d_my, d_upc;
  float band;
  int32_t upc;      // parts (units) per cell
  unsigned int* queue;
  DeviceStatus* st;
};

template <typename T, int STRIDE, int HALFMODE, int RJ>
__global__ void __launch_bounds__(RM_THREADS, NLB_RM_MINB) rowmask4_kernel(RowMask4Args<T> a) {
  pdl_enter();
  extern __shared__ __align__(128) unsigned char rm4_smem[];
  constexpr int CH = 32 * RJ;
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  Rm4Smem<RJ>& sm = reinterpret_cast<Rm4Smem<RJ>*>(rm4_smem)[warp];
  const GridParams<T>& gp = a.gp;
  const int32_t mx = gp.mesh[0], my = gp.mesh[1];
  const float msx = gp.msf[0], msy = gp.msf[1], msz = gp.msf[2];
  const TransposeConsts tc = make_transpose_consts(lane);
  unsigned long long band_local = 0, cand_local = 0;
  const unsigned int n_units = (unsigned int)gp.n_cells * (unsigned int)a.upc;  // < 2^31 (checked on the host)

  if (lane == 0) {
    mbar_init(&sm.bars[0], 1);
    mbar_init(&sm.bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  // ---- per-warp pipeline state ----
  const unsigned int n_warps = gridDim.x * RM_WARPS, gwarp = blockIdx.x * RM_WARPS + warp;
  unsigned int u_cur = gwarp;                                            // position n
  unsigned int u_nx1 = u_cur < n_units ? min(u_cur + n_warps, n_units) : n_units;  // n+1
  unsigned int u_nx2 = u_nx1 < n_units ? min(u_nx1 + n_warps, n_units) : n_units;  // n+2
  unsigned int q_pending = 0;   // ticket of position n+3 (lane 0; atomic in flight)
  int32_t pf_x = 0, pf_y = 0;   // CellRec / cell bounds of position n+1 (lane-distributed, see issue_loads)
  int32_t pf2_x = 0, pf2_y = 0; // ... of position n+2 (loads in flight)
  uint32_t phases = 0u;

  // lane r < 9: run[r]; lane 9: {nj, self_base}; lane 10: mask_base; lane 11: {cell_start[cell], cell_start[cell+1]}
  auto issue_loads = [&](unsigned int u, int32_t& x, int32_t& y) {
    x = 0;
    y = 0;
    if (u >= n_units) return;
    const int32_t cell = (int32_t)fdiv(u, a.d_upc);
    const int32_t* crw = reinterpret_cast<const int32_t*>(a.cellrec + cell);  // 22 words
    if (lane < 11) {
      const int2 v = __ldg(reinterpret_cast<const int2*>(crw) + lane);
      x = v.x;
      y = v.y;
    } else if (lane == 11) {
      x = __ldg(a.cell_start + cell);
      y = __ldg(a.cell_start + cell + 1);
    }
  };
  // descriptor of unit u in sm.desc[b] (scalars by lane 0, run values by lanes 0..8) + TMA fetch of its first chunk
  // and first row round into buffer b
  auto make_unit = [&](unsigned int u, int b, int32_t x, int32_t y) {
    Rm4Desc& d = sm.desc[b];
    __syncwarp();  // the previous reader of this descriptor is done
    if (lane == 0) d.n_a = 0;
    if (u >= n_units) return;
    const int32_t cell = (int32_t)fdiv(u, a.d_upc);
    const int32_t part = (int32_t)u - cell * a.upc;
    const int32_t a0 = __shfl_sync(0xffffffffu, x, 11), a1 = __shfl_sync(0xffffffffu, y, 11);
    const int32_t n_a = a1 - a0;
    if (n_a == 0) return;  // warp-uniform
    const int32_t nj = __shfl_sync(0xffffffffu, x, 9);
    if (part * CH >= nj) return;  // this part has no chunk of the list
    const int32_t self_base = __shfl_sync(0xffffffffu, y, 9);
    const uint32_t mlo = (uint32_t)__shfl_sync(0xffffffffu, x, 10), mhi = (uint32_t)__shfl_sync(0xffffffffu, y, 10);
    const int32_t cyz = (int32_t)fdiv((uint32_t)cell, a.d_mx), cx = cell - cyz * mx;
    const int32_t cz = (int32_t)fdiv((uint32_t)cyz, a.d_my), cy = cyz - cz * my;
    // run r: [cs, ce) in the list, first slot s0 = cs + delta
    const int32_t prev_end = __shfl_up_sync(0xffffffffu, x, 1);
    const int32_t ce = lane < 9 ? x : 0;
    const int32_t cs = lane < 9 ? (lane == 0 ? 0 : prev_end) : 0;
    const int32_t s0 = cs + y;
    if (lane < 9) {
      d.s0[lane] = s0;
      d.cs[lane] = cs;
      d.ce[lane] = ce;
    }
    if (lane == 0) {
      d.n_a = n_a;
      d.slot_a0 = a0;
      d.nj = nj;
      d.self_base = self_base;
      d.part = part;
      d.mask_base = ((unsigned long long)mhi << 32) | mlo;
      d.ox = ((float)(cx + gp.coff[0]) + 0.5f) * msx;
      d.oy = ((float)(cy + gp.coff[1]) + 0.5f) * msy;
      d.oz = ((float)(cz + gp.coff[2]) + 0.5f) * msz;
      if (part == 0) cand_local += (unsigned long long)n_a * (unsigned long long)nj;
    }
    rm_fetch(a.rec, sm.win[b], sm.rowraw[b], &sm.bars[b], lane, s0, cs, ce, part * CH, min(nj, part * CH + CH), a0, 0,
             min(n_a, RM4_ROWCAP));
  };

  {
    // prologue: position 0 synchronously, position 1's loads in flight
    int32_t x, y;
    issue_loads(u_cur, x, y);
    make_unit(u_cur, 0, x, y);
    issue_loads(u_nx1, pf_x, pf_y);
    if (lane == 0) q_pending = atomicAdd(a.queue, 1u);
    __syncwarp();
  }

  for (int n = 0; u_cur < n_units; n++) {
    const int b = n & 1;
    // position n+1: unit + TMA fetch into the other buffer (this warp finished its previous user, position n-1);
    // position n+2: loads in flight until the next iteration
    make_unit(u_nx1, b ^ 1, pf_x, pf_y);
    issue_loads(u_nx2, pf2_x, pf2_y);
    __syncwarp();
    const Rm4Desc& d = sm.desc[b];
    const int32_t n_a = d.n_a;
    if (n_a > 0) {
      const int32_t nj = d.nj, self_base = d.self_base, slot_a0 = d.slot_a0, part = d.part;
      const float ox = d.ox, oy = d.oy, oz = d.oz;
      float4* win = sm.win[b];
      float4* rowraw = sm.rowraw[b];
      const unsigned long long need = (unsigned long long)n_a * (unsigned long long)((nj + 31) >> 5);
      const bool store = d.mask_base + need <= a.mask_cap;  // else FLAG_MASK_WORDS was raised by cellsort_kernel
      uint32_t* mcell = a.mask + d.mask_base;
      bool fetched = true;  // the prefetched first chunk + first row round are (still) in the buffers
      for (int32_t rr = 0; rr < n_a; rr += RM4_ROWCAP) {
        const int32_t nrows = min(RM4_ROWCAP, n_a - rr);
        for (int32_t c_lo = part * CH; c_lo < nj; c_lo += a.upc * CH) {
          const int32_t ncand = min(CH, nj - c_lo);
          const bool first_chunk = c_lo == part * CH;
          if (!fetched) {
            // later chunks / row rounds of a crowded cell: fetched synchronously into the same buffer
            __syncwarp();
            rm_fetch(a.rec, win, rowraw, &sm.bars[b], lane, lane < 9 ? d.s0[lane] : 0, lane < 9 ? d.cs[lane] : 0,
                     lane < 9 ? d.ce[lane] : 0, c_lo, c_lo + ncand, slot_a0, first_chunk ? rr : 0,
                     first_chunk ? rr + nrows : 0);
          }
          mbar_wait(&sm.bars[b], (phases >> b) & 1u);
          phases ^= 1u << b;
          fetched = false;
          if (first_chunk) {
            // stage the round's rows: frame of the cell's centre, pre-duplicated for the packed FMAs
            for (int32_t r = lane; r < nrows; r += 32) {
              const float4 v = rowraw[r];
              const float x = v.x - ox, y = v.y - oy, z = v.z - oz;
              const float nai = -0.5f * (fmaf(x, x, fmaf(y, y, z * z)) - gp.sl2f);
              sm.srow[r][0] = make_float4(x, x, y, y);
              sm.srow[r][1] = make_float4(z, z, nai, nai);
              const int32_t id = __float_as_int(v.w);
              sm.sid[r] = id;
              if (HALFMODE == 2) sm.scmp[r] = __ldg(a.global_ids + id);
              if (HALFMODE == 1) sm.scmp[r] = id;
            }
            __syncwarp();
          }
          // ---- the chunk's candidates: RJ per lane, shifted into the cell's frame ----
          float xj[RJ], yj[RJ], zj[RJ], wj[RJ];
          int32_t cj[RJ];   // HALF: id the rule compares for candidate k
          int32_t cut[RJ];  // HALFMODE 1: rows of the round with an id below the candidate's (a prefix)
#pragma unroll
          for (int k = 0; k < RJ; k++) {
            const int32_t c = k * 32 + lane;
            const float4 v = win[min(c, ncand - 1)];
            xj[k] = v.x - ox;
            yj[k] = v.y - oy;
            zj[k] = v.z - oz;
            wj[k] = -0.5f * fmaf(xj[k], xj[k], fmaf(yj[k], yj[k], zj[k] * zj[k]));
            cj[k] = __float_as_int(v.w);
            if (c >= ncand) {
              xj[k] = yj[k] = zj[k] = 0.f;
              wj[k] = -1.0e30f;  // d = -1e30: a miss, far from the band
              cj[k] = 0x80000000;
            } else if (HALFMODE == 2) {
              cj[k] = __ldg(a.global_ids + cj[k]);
            }
            cut[k] = 0;
          }
          if (HALFMODE == 1) {
            // rows keep partners with a LARGER id (neighlist_cpu.hpp:225-236): seen from candidate j, the rows with
            // id_i < id_j — a prefix of the round, ids ascend with the slot inside a cell.  Lower bound over the
            // staged ids, branch-free, the RJ searches interleaved
            for (int32_t step = 1 << (31 - __clz(nrows)); step > 0; step >>= 1) {
#pragma unroll
              for (int k = 0; k < RJ; k++) {
                const int32_t t = cut[k] + step;
                if (t <= nrows && sm.scmp[min(t, nrows) - 1] < cj[k]) cut[k] = t;
              }
            }
          }
          f32x2 X[RJ / 2], Y[RJ / 2], Z[RJ / 2], W[RJ / 2];
#pragma unroll
          for (int h = 0; h < RJ / 2; h++) {
            X[h] = pack2(xj[2 * h], xj[2 * h + 1]);
            Y[h] = pack2(yj[2 * h], yj[2 * h + 1]);
            Z[h] = pack2(zj[2 * h], zj[2 * h + 1]);
            W[h] = pack2(wj[2 * h], wj[2 * h + 1]);
          }
          const int32_t nblk = (ncand + 31) >> 5;  // blocks of this chunk that hold candidates
          const int32_t kblk0 = c_lo >> 5;          // their index in the cell's list
          for (int32_t w = 0; w * 32 < nrows; w++) {
            const int32_t cnt = min(32, nrows - w * 32);
            const ulonglong2* sp = reinterpret_cast<const ulonglong2*>(&sm.srow[w * 32][0]);
            // rows are walked downwards so that row ii ends at bit ii; rows >= cnt keep the initial ones (= miss)
            uint32_t miss[RJ];
#pragma unroll
            for (int k = 0; k < RJ; k++) miss[k] = 0xffffffffu;
            float mh[RJ / 2];  // min |d| per candidate pair
#pragma unroll
            for (int h = 0; h < RJ / 2; h++) mh[h] = 3.0e38f;
#pragma unroll 2
            for (int32_t ii = cnt - 1; ii >= 0; ii--) {
              const ulonglong2 p0 = sp[2 * ii];      // {xi, xi}, {yi, yi}
              const ulonglong2 p1 = sp[2 * ii + 1];  // {zi, zi}, {-ai, -ai}
#pragma unroll
              for (int h = 0; h < RJ / 2; h++) {
                const f32x2 d2 = add2(fma2(p0.x, X[h], fma2(p0.y, Y[h], fma2(p1.x, Z[h], W[h]))), p1.y);
                float d0, d1;
                unpack2(d2, d0, d1);
                miss[2 * h] = __funnelshift_l(__float_as_uint(d0), miss[2 * h], 1);  // shift the sign bit in
                miss[2 * h + 1] = __funnelshift_l(__float_as_uint(d1), miss[2 * h + 1], 1);
                mh[h] = fminf(mh[h], fminf(fabsf(d0), fabsf(d1)));
              }
            }
            uint32_t hits[RJ];
#pragma unroll
            for (int k = 0; k < RJ; k++) hits[k] = ~miss[k];  // bit ii <-> row w*32 + ii
            // tests inside the pre-filter's uncertainty band are decided exactly, in the caller's precision, by the
            // whole warp (lane = row, the triggering lane's candidate broadcast)
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
                  const int32_t c_src = k * 32 + src;
                  bool fix = false, hit = false;
                  if (lane < cnt && c_src < ncand) {
                    const float4 q0 = sm.srow[w * 32 + lane][0], q1 = sm.srow[w * 32 + lane][1];
                    const float dd = pre_d(q0.x, q0.z, q1.x, q1.z, cx2[e], cy2[e], cz2[e], cw2[e]);
                    if (fabsf(dd) < a.band) {
                      const int32_t iid = sm.sid[w * 32 + lane];
                      const int32_t jid = __float_as_int(win[c_src].w);
                      hit = exact_within(load_pos<T, STRIDE>(a.q, iid), load_pos<T, STRIDE>(a.q, jid), gp.sl2);
                      fix = true;
                      band_local++;
                    }
                  }
                  const uint32_t fixm = __ballot_sync(0xffffffffu, fix);  // lane ii <-> bit ii
                  const uint32_t hitm = __ballot_sync(0xffffffffu, hit);
                  if (lane == src) hits[k] = (hits[k] & ~fixm) | hitm;
                }
              }
            }
            if (HALFMODE == 1) {
#pragma unroll
              for (int k = 0; k < RJ; k++) {
                const int32_t keep = cut[k] - w * 32;  // rows w*32 .. w*32 + keep - 1 have an id below the candidate's
                hits[k] = keep <= 0 ? 0u : (keep >= 32 ? hits[k] : (hits[k] & ((1u << keep) - 1u)));
              }
            }
            if (HALFMODE == 2) {
              // ids are not monotone in the slot when they come from a map: compare per set bit
#pragma unroll
              for (int k = 0; k < RJ; k++) {
                uint32_t m = hits[k];
                while (m) {
                  const int bpos = __ffs(m) - 1;
                  m &= m - 1;
                  if (!(cj[k] > sm.scmp[w * 32 + bpos])) hits[k] &= ~(1u << bpos);
                }
              }
            }
            // lane = candidate, bit = row   ->   lane = row, bit = candidate
            const int32_t rg = rr + w * 32 + lane;  // this lane's row inside the cell
            const int32_t cself = self_base + rg;   // ... and its own position in the candidate list (FULL: j != i)
            const int32_t rid = sm.sid[min(w * 32 + lane, nrows - 1)];
            const bool owned = lane < cnt && rid < a.n_owned;
            int32_t pc = 0;
#pragma unroll
            for (int k = 0; k < RJ; k++) {
              if (k < nblk) {
                uint32_t t = transpose32(hits[k], tc);
                if (HALFMODE == 0 && (cself >> 5) == kblk0 + k) t &= ~(1u << (cself & 31));
                pc += __popc(t);
                if (owned && store) mcell[(size_t)(kblk0 + k) * (size_t)n_a + (size_t)rg] = t;
              }
            }
            if (owned && pc) atomicAdd(a.counts + rid, pc);
          }
        }
        fetched = false;
      }
    }
    // shift the pipeline: the ticket requested an iteration ago names position n+3
    const unsigned int t = __shfl_sync(0xffffffffu, q_pending, 0);
    if (lane == 0) q_pending = atomicAdd(a.queue, 1u);
    const unsigned long long nxt = 3ull * n_warps + t;
    u_cur = u_nx1;
    u_nx1 = u_nx2;
    u_nx2 = nxt < n_units ? (unsigned int)nxt : n_units;
    pf_x = pf2_x;
    pf_y = pf2_y;
    __syncwarp();
  }
  if (cand_local) atomicAdd(&a.st->candidates, cand_local);
  if (band_local) atomicAdd(&a.st->band_tests, band_local);
}

// ---------------------------------------------------------------------------------------------------------------
// emission from row masks
// ---------------------------------------------------------------------------------------------------------------
struct Emit3Args {
  const int32_t* cell_start;
  const int32_t* sorted_ids;
  const int32_t* slot_cell;
  const int32_t* slot_pid;  // partner id reported for a slot: sorted_ids, or the global ids in slot order
  const CellRec* cellrec;
  const uint32_t* mask;
  int32_t n_total, n_owned, n_cells;
  const int64_t* offsets;
  int32_t* partners;
  long long capacity;
  const DeviceStatus* st;
};

#ifndef NLB_EM3_WARPS
#define NLB_EM3_WARPS 2
#endif
constexpr int EM3_WARPS = NLB_EM3_WARPS;
#ifndef NLB_EM3_MINB
#define NLB_EM3_MINB 14
#endif

__global__ void __launch_bounds__(EM3_WARPS * 32, NLB_EM3_MINB) emit3_kernel(Emit3Args a) {
  pdl_enter();
  extern __shared__ __align__(16) int32_t em3_smem[];
  if (a.offsets[a.n_owned] > a.capacity) return;         // overflow already flagged by the offsets scan
  if (a.st->flags & FLAG_MASK_WORDS) return;             // masks incomplete: the build fails
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  int32_t* line = em3_smem + (warp * 32 + lane) * EM_LINE;
  const uint32_t line_sa = (uint32_t)__cvta_generic_to_shared(line);
  const int32_t slot = (blockIdx.x * EM3_WARPS + warp) * 32 + lane;
  int32_t id = 0x7fffffff;
  if (slot < a.n_total && slot < __ldg(a.cell_start + a.n_cells)) id = __ldg(a.sorted_ids + slot);
  const bool owned = id < a.n_owned;
  int32_t K = 0, n_a = 1;
  const uint32_t* mp = a.mask;
  const int2* runs = a.cellrec[0].run;
  long long dst = 0;
  if (owned) {
    const int32_t cell = __ldg(a.slot_cell + slot);
    const CellRec* cr = a.cellrec + cell;
    const int32_t a0 = __ldg(a.cell_start + cell);
    n_a = __ldg(a.cell_start + cell + 1) - a0;
    K = (cr->nj + 31) >> 5;
    mp = a.mask + cr->mask_base + (slot - a0);
    runs = cr->run;
    dst = a.offsets[id];
  }
  int2 cur = owned ? __ldg(runs) : make_int2(0x7fffffff, 0);  // run cursor: candidates below cur.x map to slot c + cur.y
  int32_t run = 0;
  int32_t fill = 0;  // entries staged in this lane's line
  int32_t done = 0;  // entries of this row already written

  // every lane copies ITS OWN line to its row: staged slots -> partner ids (gather), scalar stores until the row
  // position is 16-byte aligned, then 16-byte vector stores; what does not fill a vector stays in the line
  auto flush = [&](bool final) {
    int32_t k = 0;
    int32_t* out = a.partners + dst + done;
    while (k < fill && ((reinterpret_cast<uintptr_t>(out + k) & 15) != 0)) {
      out[k] = __ldg(a.slot_pid + line[k]);
      k++;
    }
    while (k + 4 <= fill) {
      int4 v;
      v.x = __ldg(a.slot_pid + line[k]);
      v.y = __ldg(a.slot_pid + line[k + 1]);
      v.z = __ldg(a.slot_pid + line[k + 2]);
      v.w = __ldg(a.slot_pid + line[k + 3]);
      *reinterpret_cast<int4*>(out + k) = v;
      k += 4;
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

  int32_t kmax = K;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, d));
  for (int32_t k0 = 0; k0 < kmax; k0 += 4) {
    uint32_t wv[4];
#pragma unroll
    for (int u = 0; u < 4; u++) wv[u] = (k0 + u < K) ? __ldg(mp + (size_t)(k0 + u) * (size_t)n_a) : 0u;
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const uint32_t word = wv[u];
      if (!__any_sync(0xffffffffu, word != 0u)) continue;
      if (__any_sync(0xffffffffu, fill + __popc(word) > EM_TILE)) flush(false);  // leaves fill <= 3
      uint32_t wr = __brev(word);  // candidates ascending = bits descending: one FLO per entry
      const int32_t cb = (k0 + u) * 32 + 31;
      uint32_t wa = line_sa + 4u * (uint32_t)fill;
      fill += __popc(word);
      while (wr) {
        uint32_t p;
        asm("bfind.u32 %0, %1;" : "=r"(p) : "r"(wr));
        wr ^= 1u << p;
        const int32_t c = cb - (int32_t)p;
        while (c >= cur.x) cur = __ldg(runs + (++run));  // c < nj = end of run 8
        asm volatile("st.shared.s32 [%0], %1;" ::"r"(wa), "r"(c + cur.y) : "memory");
        wa += 4u;
      }
    }
  }
  flush(true);
}





}  // namespace nlb
