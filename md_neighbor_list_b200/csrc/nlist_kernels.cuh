// nlist_kernels.cuh — hand-written sm_100a kernels of the Verlet-list build.
//
// Pipeline of one build (one stream, replayed as a CUDA graph; the per-build state is left zeroed by the previous build):
//   bin_kernel        cell index + histogram; its last CTA scans the histogram into cell_start (scan_kernel on grids
//                     of more than 8192 cells); a slab rank's launch also SENDS its face records to the neighbours
//                                                                (reference: make_mesh, neighlist_gpu.hpp:26-41;
//                                                                 MakeMeshidOfPtcl / MakeNextDest, neighlist_cpu.hpp:134-152;
//                                                                 thrust::inclusive_scan, neighlist_gpu.hpp:173-175)
//   -> scatter_kernel    counting-sort scatter of ids            (thrust::sort_by_key, neighlist_gpu.hpp:190-199;
//                                                                 neighlist_cpu.hpp:154-160)
//   -> cellsort_kernel   ids ascending inside a cell + cell-sorted FP32 position records
//                                                                (the SortPtclData / CopyGather the reference stubbed
//                                                                 out: neighlist_cpu.hpp:176-180, neighlist_gpu.hpp:144-151)
//   -> runmask_kernel    (nlist_runmask.cuh) pair search: every ordered pair tested once, verdicts kept as bit masks,
//                        row lengths by RED                      (kernel_impl.cuh:3-436, neighlist_cpu.hpp:239-359)
//   -> scan_kernel       row lengths -> CSR offsets              (MakeNeighListForEachPtcl, neighlist_cpu.hpp:361-367)
//   -> emitrun_kernel / emitwin_kernel   bits -> partner ids, rows written with 32-byte stores
//                                                                (neighlist_cpu.hpp:369-372; replaces the row-major
//                                                                 buffer + cublasSgeam transpose, kernel_impl.cuh:217-239)
//      || finalize_kernel   status block to mapped host memory, per-build state re-zeroed (a second graph branch)
//   -> [sort_rows_kernel] [ell_kernel]
//   Other searches in this file: pairmask_kernel + rowcount_kernel + emit_kernel (round 1's pair masks, kept as
//   NLB200_OPT_KERNEL_VARIANT = 2: the independent implementation the tests compare the run masks with);
//   search_kernel (one CTA per cell, test evaluated in a count and a fill pass: variant 1 and NLB200_OPT_EXACT_ONLY);
//   emit_direct_kernel only with -DNLB_ABLATIONS.  Crowded cells: nlist_rowmask.cuh.
//
// Not a port: the reference searches with one thread/warp per particle gathering unsorted positions by id and writes
// an ELL matrix.  Here positions are physically cell-sorted as 16-byte FP32 records relative to their cell corner; the
// distance test is a dot-form FP32 pre-filter
//   |xi-xj|^2 <= SL^2  <=>  xi.xj - |xj|^2/2 - (|xi|^2 - SL^2)/2 >= 0          (3 FFMA2 + 1 FADD2 per two tests)
// whose rigorous error band (E) separates definite hits and misses; the few candidates inside the
// band are re-tested exactly in the caller's precision with the reference's rounding order, so verdicts are
// bit-identical to the reference while the hot loop runs on the packed FP32 pipe.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace nlb {

enum : uint32_t {
  FLAG_OUT_OF_BOX = 1u,
  FLAG_CAPACITY = 2u,
  FLAG_ELL_ROWS = 4u,
  FLAG_OFFSETS32 = 8u,  // total exceeds INT32_MAX: the int32 offsets view is invalid
};

// Device-resident status block, copied to pinned host memory at the end of every build.
struct DeviceStatus {
  unsigned long long total_entries;
  unsigned long long candidates;
  unsigned long long band_tests;
  uint32_t flags;
  int32_t max_partners;
  int32_t max_in_cell;
  int32_t pad;
  unsigned long long mask_words;  // row-mask words requested by the cells (cursor of the per-cell blocks)
};

constexpr uint32_t FLAG_MASK_WORDS = 32u;  // the row-mask buffer is too small (status.mask_words tells the need)

// What the row-mask search and the emission need to know about a cell (nlist_rowmask.cuh).  The candidates of a cell
// are the <= 9 contiguous x-runs of its stencil in the cell-sorted arrays, concatenated: its "candidate list".
struct alignas(8) CellRec {
  int2 run[9];  // .x: end (exclusive) of run r in the cell's candidate list; .y: first slot of run r minus its start
                //     in the list, so that slot = candidate index + .y.  Absent runs: .x = nj.
  int32_t nj;         // candidates of the cell (length of the list)
  int32_t self_base;  // candidate index of the cell's own first particle
  unsigned long long mask_base;  // first word of the cell's block: word (k, i) at mask_base + k * n_A + i
};

template <typename T>
struct GridParams {
  int32_t mesh[3];   // cells per axis of the handle's grid
  int32_t n_cells;
  // Cell window (nlb200_set_cell_window, multi-GPU): the handle's grid is the cells [coff, coff + mesh) of the global
  // grid of gmesh cells per axis.  A particle's cell is computed on the GLOBAL grid — same assignment as a single-GPU
  // build — and shifted by coff.  Without a window gmesh = mesh, coff = 0.
  int32_t gmesh[3];
  int32_t coff[3];
  T ims[3];  // 1/ms, rounded as the reference does (neighlist_gpu.hpp:250-252)
  T ms[3];   // cell edge (neighlist_gpu.hpp:246-248)
  T sl2;     // SL*SL rounded once in T (neighlist_gpu.hpp:254)
  float msf[3];
  float sl2f;
  float band;  // E: half-width of the FP32 pre-filter's uncertainty band in units of (SL2 - r2)/2
};

// ---------------------------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// Division by a run-time invariant (mesh extents, items per cell) as multiply-high + shift — a hardware integer
// division is ~20 instructions and the per-item set-up of the pair-mask kernel did seven of them.  Exact for
// 0 <= n < 2^31 (the round-up method; same construction as CUTLASS's FastDivmod).
struct FastDiv {
  uint32_t d, mul, shr;
};
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  f.mul = 0;
  f.shr = 0;
  if (d > 1) {
    uint32_t lg = 0;
    while ((1ull << lg) < d) lg++;  // ceil(log2 d)
    const uint32_t p = 31 + lg;
    f.mul = (uint32_t)(((1ull << p) + d - 1) / d);
    f.shr = p - 32;
  }
  return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) {
  return f.d > 1 ? (__umulhi(n, f.mul) >> f.shr) : n;
}

// Programmatic dependent launch (PDL): every kernel of the build chain first lets its successor's CTAs be scheduled
// (they occupy free slots only: the trigger fires once ALL CTAs of this grid have started) and then waits until its
// predecessor has completed and its writes are visible.  The chain is transitive because no kernel passes the wait
// before its own predecessor is done.  Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

template <typename T>
struct Vec3 {
  T x, y, z;
};

// 128-bit vectorised loads of one AoS position record (Guideline 13).
template <typename T, int STRIDE>
__device__ __forceinline__ Vec3<T> load_pos(const T* __restrict__ q, int64_t i);

template <>
__device__ __forceinline__ Vec3<double> load_pos<double, 4>(const double* __restrict__ q, int64_t i) {
  const double2* p = reinterpret_cast<const double2*>(q + 4 * i);
  const double2 a = __ldg(p), b = __ldg(p + 1);
  return {a.x, a.y, b.x};
}
template <>
__device__ __forceinline__ Vec3<double> load_pos<double, 3>(const double* __restrict__ q, int64_t i) {
  const double* p = q + 3 * i;
  return {__ldg(p), __ldg(p + 1), __ldg(p + 2)};
}
template <>
__device__ __forceinline__ Vec3<float> load_pos<float, 4>(const float* __restrict__ q, int64_t i) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(q + 4 * i));
  return {a.x, a.y, a.z};
}
template <>
__device__ __forceinline__ Vec3<float> load_pos<float, 3>(const float* __restrict__ q, int64_t i) {
  const float* p = q + 3 * i;
  return {__ldg(p), __ldg(p + 1), __ldg(p + 2)};
}

// The exact verdict in the caller's precision, with the rounding order nvcc/g++ give the reference expression
// `drx*drx + dry*dry + drz*drz` (kernel_impl.cuh:25-29, neighlist_cpu.hpp:219-223): fma(dz,dz, fma(dy,dy, dx*dx)).
// Explicit intrinsics so that no other contraction can be chosen by the compiler.
__device__ __forceinline__ bool exact_within(const Vec3<double>& a, const Vec3<double>& b, double sl2) {
  const double dx = __dsub_rn(a.x, b.x), dy = __dsub_rn(a.y, b.y), dz = __dsub_rn(a.z, b.z);
  const double r2 = __fma_rn(dz, dz, __fma_rn(dy, dy, __dmul_rn(dx, dx)));
  return !(r2 > sl2);
}
__device__ __forceinline__ bool exact_within(const Vec3<float>& a, const Vec3<float>& b, float sl2) {
  const float dx = __fsub_rn(a.x, b.x), dy = __fsub_rn(a.y, b.y), dz = __fsub_rn(a.z, b.z);
  const float r2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
  return !(r2 > sl2);
}

// ---- halo exchange by peer stores: control block and flags (protocol: see pack_faces_p2p_kernel below) ----
struct HaloCtrl {
  unsigned long long step;          // steps completed by this rank
  unsigned long long ready[2];      // [0]: ghosts from the lower neighbour are complete for step ready[0]; [1]: upper
  unsigned long long free_from[2];  // [0]: the lower neighbour has finished reading what this rank sent it; [1]: upper
  unsigned long long error;         // a bounded wait gave up
  unsigned long long pad[2];
};
constexpr unsigned int HALO_SPIN_LIMIT = 1u << 26;  // ~1 s of polling, then give up

__device__ __forceinline__ unsigned long long ld_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// values only THIS rank writes (its step counter, its error word): a volatile load from the L2 — no acquire needed
__device__ __forceinline__ unsigned long long ld_own(const unsigned long long* p) {
  return *reinterpret_cast<const volatile unsigned long long*>(p);
}
// a flag store that follows a __threadfence_system() of the same thread: the fence already ordered the data before it
__device__ __forceinline__ void st_sys_relaxed(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// returns false if the flag did not reach `want` within the spin limit
__device__ __forceinline__ bool halo_wait_flag(const unsigned long long* flag, unsigned long long want) {
  for (unsigned int spins = 0; spins < HALO_SPIN_LIMIT; spins++) {
    if (ld_sys(flag) >= want) return true;
    __nanosleep(64);
  }
  return false;
}

// What a slab rank sends and where (nlb200_pack_faces_p2p / nlb200_set_halo_pack).
struct HaloPackArgs {
  int axis;
  double cut_lo, cut_hi;      // a record with q[axis] < cut_lo goes to the lower neighbour, >= cut_hi to the upper
  void* out_q_lo;             // the neighbours' ghost regions (peer pointers); nullptr: no such neighbour
  int32_t* out_gid_lo;
  void* out_q_hi;
  int32_t* out_gid_hi;
  long long capacity;         // records per face region
  unsigned long long* state;  // [0..1] cursors, [2] ticket, [3..4] previous counts
  long long* out_counts;      // true counts of the step (> capacity = overflow)
  HaloCtrl* ctrl;
  unsigned long long* peer_ready_lo;
  unsigned long long* peer_ready_hi;
  int32_t* send_idx_lo;       // optional [capacity]: the local index of the record packed into slot k of the face —
  int32_t* send_idx_hi;       // the recorded face set nlb200_halo_refresh re-sends between two builds
};

template <typename T, bool WAIT_OWN>
__device__ __forceinline__ void halo_pack_cta(const T* __restrict__ q, const int32_t* __restrict__ gids, int64_t i,
                                              int64_t n, int stride, const HaloPackArgs& hp);

// ---------------------------------------------------------------------------------------------------------------
// 1. cell index + histogram
// ---------------------------------------------------------------------------------------------------------------
// idx = int(q * ims): the reference's reciprocal multiply + truncation (neighlist_gpu.hpp:31-35,
// neighlist_cpu.hpp:52-56).  The reference GPU kernel clamps idx == mesh_size to mesh_size-1 and the CPU class wraps
// one period; because distances are not periodic both give the same pair set for inputs in [0,L].
// Here every index is clamped into [0, mesh-1], which additionally keeps particles slightly outside the box next to
// their true neighbours; a particle more than one cell outside (or NaN) raises FLAG_OUT_OF_BOX because the FP32
// pre-filter's error bound assumes |x - cell corner| <= 2 cells.
// SCAN_HERE: the LAST CTA to finish (ticket counter) also scans the histogram into cell_start and records the most
// crowded cell — one kernel boundary less for grids of up to BIN_SCAN_MAX_CELLS cells (the separate look-back scan
// serves larger grids: one CTA needs ~1 us per thousand cells).  For the default system (3375 cells) one CTA scans in ~2 us what cost a launch of its own.
constexpr int BIN_SCAN_MAX_CELLS = 1 << 13;
// HALO (a slab rank whose exchange is folded into its build, nlb200_set_halo_pack): the build bins in two launches.
//   HALO = 1, records [0, n_owned): the CTA that bins an owned record also SENDS it if it lies within the search
//             length of a face (halo_pack_cta: peer stores into the neighbour's ghost region; the last CTA raises the
//             neighbours' `ready` flags) — the separate packing kernel, its pass over the positions and its launch
//             are gone, and the flight of the flags over NVLink overlaps the kernel boundary;
//   HALO = 2, records [n_owned, n): every CTA first waits for this rank's own `ready` flags, then bins the ghosts.
template <typename T, int STRIDE, bool SCAN_HERE, int HALO = 0>
__global__ void __launch_bounds__(256) bin_kernel(const T* __restrict__ q, int32_t i0, int32_t n, int32_t n_owned,
                                                  GridParams<T> gp, int32_t* __restrict__ cell_count,
                                                  int2* __restrict__ cell_rank, DeviceStatus* __restrict__ st,
                                                  int32_t* __restrict__ cell_start, unsigned int* __restrict__ ticket,
                                                  const int32_t* __restrict__ gids, HaloPackArgs hp) {
  pdl_enter();
  const int32_t i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (HALO == 2) {
    if (threadIdx.x < 2) {  // lane f waits for face f: both acquire loads in flight together
      const unsigned long long want = ld_own(&hp.ctrl->step) + 1ull;
      const bool face = threadIdx.x == 0 ? hp.out_q_lo != nullptr : hp.out_q_hi != nullptr;
      if (face && !halo_wait_flag(&hp.ctrl->ready[threadIdx.x], want)) hp.ctrl->error = 1ull;
    }
    __syncthreads();  // orders every thread's ghost loads after the two acquires
  }
  if (i < n) {
    const Vec3<T> p = load_pos<T, STRIDE>(q, i);
    if (i >= n_owned && p.x != p.x) {
      // an ABSENT ghost: the fixed-capacity halo buffers of the multi-GPU build are padded with NaN records
      // (parallel.py); they take part in nothing — no cell, no slot, no row
      cell_rank[i] = make_int2(-1, 0);
    } else {
      int32_t cx = static_cast<int32_t>(p.x * gp.ims[0]);
      int32_t cy = static_cast<int32_t>(p.y * gp.ims[1]);
      int32_t cz = static_cast<int32_t>(p.z * gp.ims[2]);
      cx = min(max(cx, 0), gp.gmesh[0] - 1);
      cy = min(max(cy, 0), gp.gmesh[1] - 1);
      cz = min(max(cz, 0), gp.gmesh[2] - 1);
      const double rx = (double)p.x - (double)cx * (double)gp.ms[0];
      const double ry = (double)p.y - (double)cy * (double)gp.ms[1];
      const double rz = (double)p.z - (double)cz * (double)gp.ms[2];
      bool ok = (rx >= -(double)gp.ms[0]) && (rx <= 2.0 * (double)gp.ms[0]) && (ry >= -(double)gp.ms[1]) &&
                (ry <= 2.0 * (double)gp.ms[1]) && (rz >= -(double)gp.ms[2]) && (rz <= 2.0 * (double)gp.ms[2]);
      // into the handle's window of the global grid; a particle outside it was given to the wrong handle
      cx -= gp.coff[0];
      cy -= gp.coff[1];
      cz -= gp.coff[2];
      ok = ok && cx >= 0 && cx < gp.mesh[0] && cy >= 0 && cy < gp.mesh[1] && cz >= 0 && cz < gp.mesh[2];
      cx = min(max(cx, 0), gp.mesh[0] - 1);
      cy = min(max(cy, 0), gp.mesh[1] - 1);
      cz = min(max(cz, 0), gp.mesh[2] - 1);
      if (!ok) atomicOr(&st->flags, FLAG_OUT_OF_BOX);
      const int32_t cell = cx + (cy + cz * gp.mesh[1]) * gp.mesh[0];
      const int32_t rank = atomicAdd(&cell_count[cell], 1);
      cell_rank[i] = make_int2(cell, rank);
    }
  }
  if (HALO == 1) halo_pack_cta<T, false>(q, gids, (int64_t)i, (int64_t)n, STRIDE, hp);
  if (!SCAN_HERE) return;
  // ---- the last CTA scans the histogram ----
  __shared__ bool is_last;
  __shared__ int32_t wsum[8], wmax[8];
  __syncthreads();
  if (threadIdx.x == 0) {
    // one fence, after the barrier: it is cumulative over the CTA's histogram updates (a MEMBAR.SC by every thread
    // was 30 % of the kernel's stall samples, profiles/r02_ncu_bin.txt)
    __threadfence();
    is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const int32_t M = gp.n_cells;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int32_t carry = 0, cmax = 0;
  constexpr int CPT = 16;  // consecutive cells per thread: the default system's 3375 cells take ONE round
  for (int32_t b0 = 0; b0 < M; b0 += 256 * CPT) {
    const int32_t c0 = b0 + threadIdx.x * CPT;
    int32_t v[CPT];
    if (c0 + CPT <= M) {
#pragma unroll
      for (int k = 0; k < CPT; k += 4) {
        const int4 t4 = __ldcg(reinterpret_cast<const int4*>(cell_count + c0 + k));
        v[k] = t4.x;
        v[k + 1] = t4.y;
        v[k + 2] = t4.z;
        v[k + 3] = t4.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < CPT; k++) v[k] = (c0 + k < M) ? __ldcg(cell_count + c0 + k) : 0;
    }
    int32_t tsum = 0, tm = 0;
#pragma unroll
    for (int k = 0; k < CPT; k++) {
      tsum += v[k];
      tm = max(tm, v[k]);
    }
    int32_t incl = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int32_t o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) tm = max(tm, __shfl_xor_sync(0xffffffffu, tm, d));
    __syncthreads();  // the previous round's wsum readers are done
    if (lane == 31) wsum[w] = incl;
    if (lane == 0) wmax[w] = tm;
    __syncthreads();
    int32_t wpre = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int32_t s2 = wsum[k];
      if (k < w) wpre += s2;
      tot += s2;
      cmax = max(cmax, wmax[k]);
    }
    int32_t run = carry + wpre + incl - tsum;
    if (c0 + CPT <= M) {
#pragma unroll
      for (int k = 0; k < CPT; k += 4) {
        int4 o4;
        o4.x = run;
        o4.y = run + v[k];
        o4.z = o4.y + v[k + 1];
        o4.w = o4.z + v[k + 2];
        run = o4.w + v[k + 3];
        *reinterpret_cast<int4*>(cell_start + c0 + k) = o4;
      }
    } else {
#pragma unroll
      for (int k = 0; k < CPT; k++) {
        if (c0 + k < M) cell_start[c0 + k] = run;
        run += v[k];
      }
    }
    carry += tot;
  }
  if (threadIdx.x == 0) {
    cell_start[M] = carry;
    if (cmax > 0) atomicMax(&st->max_in_cell, cmax);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 2. single-pass exclusive scan (decoupled look-back), int32 in -> TOut out, out has n+1 entries
// ---------------------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// tile status word: bits 63..62 = state (0 invalid, 1 aggregate, 2 inclusive prefix), bits 61..0 = value
constexpr unsigned long long SCAN_AGG = 1ull << 62;
constexpr unsigned long long SCAN_PFX = 2ull << 62;
constexpr unsigned long long SCAN_VAL = (1ull << 62) - 1;

// `state` = [0]: dynamic tile counter, [1..]: tile status words; zeroed before launch.
// If out32 != nullptr the same offsets are also written as int32 (the reference's key_pointer_ width).
// If st != nullptr: the grand total goes to st->total_entries, the maximum input to *max_out, and FLAG_CAPACITY /
// FLAG_OFFSETS32 are raised against `capacity`.
template <typename TOut>
__global__ void __launch_bounds__(SCAN_THREADS) scan_kernel(const int32_t* __restrict__ in, int64_t n,
                                                            TOut* __restrict__ out, int32_t* __restrict__ out32,
                                                            unsigned long long* __restrict__ state,
                                                            DeviceStatus* __restrict__ st, int32_t* max_out,
                                                            long long capacity) {
  pdl_enter();
  __shared__ long long warp_sums[SCAN_THREADS / 32];
  __shared__ long long tile_prefix_s;
  __shared__ int tile_s;
  __shared__ int warp_max[SCAN_THREADS / 32];
  if (threadIdx.x == 0) tile_s = (int)atomicAdd(&state[0], 1ull);
  __syncthreads();
  const int tile = tile_s;
  const int64_t base = (int64_t)tile * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int32_t v[SCAN_ITEMS];
  if (base + SCAN_ITEMS <= n && ((reinterpret_cast<uintptr_t>(in + base) & 15) == 0)) {
    const int4 a = *reinterpret_cast<const int4*>(in + base);
    const int4 b = *reinterpret_cast<const int4*>(in + base + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) v[k] = (base + k < n) ? in[base + k] : 0;
  }
  long long tsum = 0;
  int tmax = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    tsum += v[k];
    tmax = max(tmax, v[k]);
  }
  // block-wide inclusive scan of the per-thread sums
  long long incl = tsum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const long long o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane_id() >= d) incl += o;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) tmax = max(tmax, __shfl_xor_sync(0xffffffffu, tmax, d));
  const int w = threadIdx.x >> 5;
  if (lane_id() == 31) warp_sums[w] = incl;
  if (lane_id() == 0) warp_max[w] = tmax;
  __syncthreads();
  long long wpre = 0, tile_total = 0;
#pragma unroll
  for (int k = 0; k < SCAN_THREADS / 32; k++) {
    const long long s = warp_sums[k];
    if (k < w) wpre += s;
    tile_total += s;
  }
  // publish + look back (first warp)
  if (w == 0) {
    if (max_out != nullptr && lane_id() == 0) {
      int m = 0;
#pragma unroll
      for (int k = 0; k < SCAN_THREADS / 32; k++) m = max(m, warp_max[k]);
      if (m > 0) atomicMax(max_out, m);
    }
    long long prefix = 0;
    if (tile == 0) {
      if (lane_id() == 0) {
        atomicExch(&state[1], SCAN_PFX | ((unsigned long long)tile_total & SCAN_VAL));
      }
    } else {
      if (lane_id() == 0) {
        atomicExch(&state[1 + tile], SCAN_AGG | ((unsigned long long)tile_total & SCAN_VAL));
      }
      int look = tile - 1;
      while (true) {
        const int idx = look - lane_id();
        unsigned long long s = SCAN_PFX;  // lanes before tile 0 act as a zero prefix
        if (idx >= 0) {
          do {
            s = *reinterpret_cast<volatile unsigned long long*>(&state[1 + idx]);
          } while ((s >> 62) == 0);
        }
        const unsigned pfx_mask = __ballot_sync(0xffffffffu, (s >> 62) == 2);
        const int first_pfx = pfx_mask ? (__ffs(pfx_mask) - 1) : 32;
        long long contrib = (lane_id() <= first_pfx) ? (long long)(s & SCAN_VAL) : 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, d);
        prefix += contrib;
        if (pfx_mask) break;
        look -= 32;
      }
      if (lane_id() == 0) {
        atomicExch(&state[1 + tile], SCAN_PFX | ((unsigned long long)(prefix + tile_total) & SCAN_VAL));
      }
    }
    if (lane_id() == 0) tile_prefix_s = prefix;
  }
  __syncthreads();
  long long run = tile_prefix_s + wpre + (incl - tsum);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    if (base + k < n) {
      out[base + k] = (TOut)run;
      if (out32 != nullptr) out32[base + k] = (int32_t)run;
    }
    run += v[k];
  }
  // the thread that owns the last element writes the grand total
  if (n > 0 && base <= n - 1 && n - 1 < base + SCAN_ITEMS) {
    out[n] = (TOut)run;
    if (out32 != nullptr) out32[n] = (int32_t)run;
    if (st != nullptr) {
      st->total_entries = (unsigned long long)run;
      uint32_t f = 0;
      if (run > capacity) f |= FLAG_CAPACITY;
      if (run > 2147483647ll) f |= FLAG_OFFSETS32;
      if (f) atomicOr(&st->flags, f);
    }
  }
  if (n == 0 && tile == 0 && threadIdx.x == 0) {
    out[0] = (TOut)0;
    if (out32 != nullptr) out32[0] = 0;
    if (st != nullptr) st->total_entries = 0;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 3. counting-sort scatter of ids (arrival order inside a cell; made deterministic by cellsort_kernel)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) scatter_kernel(const int2* __restrict__ cell_rank, int32_t n,
                                                      const int32_t* __restrict__ cell_start,
                                                      int32_t* __restrict__ perm) {
  pdl_enter();
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int2 cr = cell_rank[i];
  if (cr.x < 0) return;  // absent ghost
  perm[__ldg(cell_start + cr.x) + cr.y] = i;
}

// ---------------------------------------------------------------------------------------------------------------
// 4. per-cell id sort (stable counting-sort order = ids ascending, neighlist_cpu.hpp:154-160) + physical reorder:
//    rec[slot] = { float(x - cx*ms), float(y - cy*ms), float(z - cz*ms), id }   (cell-corner-relative FP32)
//    sorted_ids[slot] = id, slot_cell[slot] = cell
//    One warp per cell; rank sort with warp shuffles (O(n^2/32) per cell, n ~ 35).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void axis_range(int c, int m, int& lo, int& hi);

// ABS: records hold the ABSOLUTE coordinates rounded to FP32 (the row-mask search shifts them into a cell frame with
// three subtractions; its band E accounts for the rounding, nlist_api.cu), else coordinates relative to the
// particle's own cell corner (round-1 search kernels).
// cellrec != nullptr (row-mask path): the warp also writes the cell's CellRec — run table of its candidate list, its
// length, the cell's own position in it — takes the cell's block of the row-mask buffer from a cursor (one atomicAdd
// per cell: the buffer holds one bit per test whatever the density profile) and zeroes the row lengths of its owned
// particles, which the search accumulates with atomics.
template <typename T, int STRIDE, bool ABS>
__global__ void __launch_bounds__(128) cellsort_kernel(const T* __restrict__ q, GridParams<T> gp,
                                                       const int32_t* __restrict__ cell_start,
                                                       const int32_t* __restrict__ perm,
                                                       int32_t* __restrict__ sorted_ids, float4* __restrict__ rec,
                                                       int32_t* __restrict__ slot_cell,
                                                       const int32_t* __restrict__ global_ids,
                                                       int32_t* __restrict__ slot_pid, CellRec* __restrict__ cellrec,
                                                       unsigned long long mask_cap, int32_t* __restrict__ counts,
                                                       int32_t n_owned, DeviceStatus* __restrict__ st, int32_t batch,
                                                       int32_t big_thr) {
  pdl_enter();
  // big_thr > 0: the particles of cells of more than big_thr particles are ranked and written by cellsort_big_kernel
  // (one warp per 32 slots) — the shuffle ranking below is O(cnt^2 / 32) steps of ONE warp per cell: 11.7 ms of a
  // 30 ms build of 2^20 clustered particles (cells of 4000) before the split.
  // warps stride over batches of `batch` consecutive cells (a slab rank bins on the global grid: most of its cells
  // are empty).  batch > 1 on large grids: the blocks of a batch's cells are taken from the row-mask cursor with ONE
  // atomicAdd (lane = cell computes its need first) — one atomic per cell on a single address costs ~2 ns each,
  // 0.9 ms at 456 k cells.
  const int lane = lane_id();
  const int32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (int32_t c0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * batch; c0 < gp.n_cells; c0 += warps * batch) {
  unsigned long long batch_base = 0;
  if (cellrec != nullptr && batch > 1) {
    // lane = cell c0 + lane: candidates of its stencil (sum over the <= 9 runs) -> words it needs
    unsigned long long need = 0;
    const int32_t cl = c0 + lane;
    if (lane < batch && cl < gp.n_cells) {
      const int32_t na = __ldg(cell_start + cl + 1) - __ldg(cell_start + cl);
      if (na > 0) {
        const int32_t mx = gp.mesh[0], my = gp.mesh[1], mz = gp.mesh[2];
        const int32_t lx = cl % mx, ly = (cl / mx) % my, lzc = cl / (mx * my);
        int xlo, xhi, ylo, yhi, zlo, zhi;
        axis_range(lx, mx, xlo, xhi);
        axis_range(ly, my, ylo, yhi);
        axis_range(lzc, mz, zlo, zhi);
        int32_t nj = 0;
        for (int z = zlo; z <= zhi; z++)
          for (int y = ylo; y <= yhi; y++) {
            const int32_t* cs = cell_start + (y + z * my) * mx;
            nj += __ldg(cs + xhi + 1) - __ldg(cs + xlo);
          }
        need = (unsigned long long)na * (unsigned long long)((nj + 31) >> 5);
      }
    }
    unsigned long long incl = need;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) {
      const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, dd);
      if (lane >= dd) incl += v;
    }
    const unsigned long long total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned long long base = 0;
    if (lane == 0 && total > 0) {
      base = atomicAdd(&st->mask_words, total);
      if (base + total > mask_cap) atomicOr(&st->flags, FLAG_MASK_WORDS);
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    batch_base = base + incl - need;  // lane's cell
  }
  for (int32_t cb = 0; cb < batch; cb++) {
  const int32_t cell = c0 + cb;
  if (cell >= gp.n_cells) break;
  const unsigned long long my_base = __shfl_sync(0xffffffffu, batch_base, cb);
  const int32_t beg = __ldg(cell_start + cell);
  const int32_t cnt = __ldg(cell_start + cell + 1) - beg;
  if (cnt == 0) continue;
  const int32_t cx = cell % gp.mesh[0];
  const int32_t cy = (cell / gp.mesh[0]) % gp.mesh[1];
  const int32_t cz = cell / (gp.mesh[0] * gp.mesh[1]);
  unsigned long long mbase = 0;
  int32_t cr_ce = 0, cr_delta = 0, cr_nj = 0, cr_self = 0;
  if (cellrec != nullptr) {
    // run r = (z, y) of the stencil: the cells [xlo, xhi] of that row are contiguous in the cell-sorted arrays
    const int32_t mx = gp.mesh[0], my = gp.mesh[1], mz = gp.mesh[2];
    int xlo, xhi, ylo, yhi, zlo, zhi;
    axis_range(cx, mx, xlo, xhi);
    axis_range(cy, my, ylo, yhi);
    axis_range(cz, mz, zlo, zhi);
    const int32_t ny = yhi - ylo + 1, nruns = ny * (zhi - zlo + 1);
    int32_t s0 = 0, len = 0;
    if (lane < nruns) {
      const int lz = lane / ny;
      const int z = zlo + lz, y = ylo + lane - lz * ny;
      const int32_t* cs = cell_start + (y + z * my) * mx;
      s0 = __ldg(cs + xlo);
      len = __ldg(cs + xhi + 1) - s0;
    }
    int32_t incl = len;
#pragma unroll
    for (int dd = 1; dd < 16; dd <<= 1) {
      const int32_t v = __shfl_up_sync(0xffffffffu, incl, dd);
      if (lane >= dd) incl += v;
    }
    cr_nj = __shfl_sync(0xffffffffu, incl, 8);
    cr_ce = incl;
    cr_delta = s0 - (incl - len);
    const int r_own = (cz - zlo) * ny + (cy - ylo);
    cr_self = __shfl_sync(0xffffffffu, incl - len, r_own) + (beg - __shfl_sync(0xffffffffu, s0, r_own));
    if (batch > 1) {
      mbase = my_base;
    } else if (lane == 0) {
      const unsigned long long need = (unsigned long long)cnt * (unsigned long long)((cr_nj + 31) >> 5);
      mbase = atomicAdd(&st->mask_words, need);  // consumed after the id sort below
      if (mbase + need > mask_cap) atomicOr(&st->flags, FLAG_MASK_WORDS);
    }
  }
  const double ox = (double)(cx + gp.coff[0]) * (double)gp.ms[0], oy = (double)(cy + gp.coff[1]) * (double)gp.ms[1],
               oz = (double)(cz + gp.coff[2]) * (double)gp.ms[2];
  for (int32_t eb = 0; eb < ((big_thr > 0 && cnt > big_thr) ? 0 : cnt); eb += 32) {
    const int32_t e = eb + lane;
    const bool valid = e < cnt;
    const int32_t id = valid ? __ldg(perm + beg + e) : 0x7fffffff;
    // sort key: the id the rows report — with a local -> global map the GLOBAL id, so that a cell's particles come
    // out in the order a single-GPU build of the whole system gives them, whatever order the ghosts arrived in
    const int32_t key = (valid && global_ids != nullptr) ? __ldg(global_ids + id) : id;
    // the record is requested before the ranking loop: its (random, L2) latency overlaps the shuffles
    const Vec3<T> p = load_pos<T, STRIDE>(q, valid ? id : 0);
    int32_t rank = 0;
    for (int32_t cb = 0; cb < cnt; cb += 32) {
      int32_t other = (cb + lane < cnt) ? __ldg(perm + beg + cb + lane) : 0x7fffffff;
      if (global_ids != nullptr && cb + lane < cnt) other = __ldg(global_ids + other);
      const int lim = min(32, cnt - cb);
      for (int t = 0; t < lim; t++) rank += (__shfl_sync(0xffffffffu, other, t) < key) ? 1 : 0;
    }
    if (valid) {
      float4 r;
      r.x = ABS ? (float)p.x : (float)((double)p.x - ox);
      r.y = ABS ? (float)p.y : (float)((double)p.y - oy);
      r.z = ABS ? (float)p.z : (float)((double)p.z - oz);
      r.w = __int_as_float(id);
      sorted_ids[beg + rank] = id;
      rec[beg + rank] = r;
      slot_cell[beg + rank] = cell;
      // the id a row reports for this particle: with a local -> global map (multi-GPU) the emission gathers the
      // global id straight from the slot instead of slot -> local id -> global id
      if (global_ids != nullptr) slot_pid[beg + rank] = key;
      if (counts != nullptr && id < n_owned) counts[id] = 0;
    }
  }
  if (cellrec != nullptr) {
    CellRec& cr = cellrec[cell];
    if (lane < 9) cr.run[lane] = make_int2(cr_ce, cr_delta);
    if (lane == 0) {
      cr.nj = cr_nj;
      cr.self_base = cr_self;
      cr.mask_base = mbase;
    }
  }
  }
  }
}

// Crowded cells (clustered inputs; only launched beside cellsort_kernel with big_thr > 0): thread = slot of the
// arrival-ordered array.  A particle of a cell of more than CS_BIG particles finds its rank by counting the smaller
// keys of its cell — cnt iterations per thread, but every 32 slots of the cell have a warp of their own, and the lanes
// of a warp read the same addresses (one wavefront per load).  Same outputs as cellsort_kernel.
constexpr int CS_BIG = 256;
template <typename T, int STRIDE, bool ABS>
__global__ void __launch_bounds__(128) cellsort_big_kernel(const T* __restrict__ q, GridParams<T> gp,
                                                           const int32_t* __restrict__ cell_start,
                                                           const int2* __restrict__ cell_rank,
                                                           const int32_t* __restrict__ perm,
                                                           int32_t* __restrict__ sorted_ids, float4* __restrict__ rec,
                                                           int32_t* __restrict__ slot_cell,
                                                           const int32_t* __restrict__ global_ids,
                                                           int32_t* __restrict__ slot_pid, int32_t* __restrict__ counts,
                                                           int32_t n_owned, int32_t big_thr) {
  pdl_enter();
  const int32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  const bool present = slot < __ldg(cell_start + gp.n_cells);
  const int32_t id = present ? __ldg(perm + slot) : 0;
  const int32_t cell = present ? cell_rank[id].x : 0;
  const int32_t beg = __ldg(cell_start + cell);
  const int32_t cnt = __ldg(cell_start + cell + 1) - beg;
  const bool big = present && cnt > big_thr;
  if (!__any_sync(0xffffffffu, big)) return;
  if (!big) return;
  const int32_t key = global_ids != nullptr ? __ldg(global_ids + id) : id;
  const Vec3<T> p = load_pos<T, STRIDE>(q, id);
  int32_t rank = 0;
  const int32_t* cp = perm + beg;
  int32_t t = 0;
  if (global_ids != nullptr) {
    for (; t + 4 <= cnt; t += 4) {
      const int32_t o0 = __ldg(cp + t), o1 = __ldg(cp + t + 1), o2 = __ldg(cp + t + 2), o3 = __ldg(cp + t + 3);
      const int32_t k0 = __ldg(global_ids + o0), k1 = __ldg(global_ids + o1), k2 = __ldg(global_ids + o2),
                    k3 = __ldg(global_ids + o3);
      rank += (k0 < key) + (k1 < key) + (k2 < key) + (k3 < key);
    }
    for (; t < cnt; t++) rank += __ldg(global_ids + __ldg(cp + t)) < key ? 1 : 0;
  } else {
    for (; t + 4 <= cnt; t += 4) {
      const int32_t o0 = __ldg(cp + t), o1 = __ldg(cp + t + 1), o2 = __ldg(cp + t + 2), o3 = __ldg(cp + t + 3);
      rank += (o0 < key) + (o1 < key) + (o2 < key) + (o3 < key);
    }
    for (; t < cnt; t++) rank += __ldg(cp + t) < key ? 1 : 0;
  }
  const int32_t cx = cell % gp.mesh[0];
  const int32_t cy = (cell / gp.mesh[0]) % gp.mesh[1];
  const int32_t cz = cell / (gp.mesh[0] * gp.mesh[1]);
  const double ox = (double)(cx + gp.coff[0]) * (double)gp.ms[0], oy = (double)(cy + gp.coff[1]) * (double)gp.ms[1],
               oz = (double)(cz + gp.coff[2]) * (double)gp.ms[2];
  float4 r;
  r.x = ABS ? (float)p.x : (float)((double)p.x - ox);
  r.y = ABS ? (float)p.y : (float)((double)p.y - oy);
  r.z = ABS ? (float)p.z : (float)((double)p.z - oz);
  r.w = __int_as_float(id);
  sorted_ids[beg + rank] = id;
  rec[beg + rank] = r;
  slot_cell[beg + rank] = cell;
  if (global_ids != nullptr) slot_pid[beg + rank] = key;
  if (counts != nullptr && id < n_owned) counts[id] = 0;
}

// ---------------------------------------------------------------------------------------------------------------
// 5/7. pair search.  One CTA per cell; thread t owns the t-th particle of the cell.
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
struct SearchArgs {
  const T* q;                 // caller's positions (exact re-test only)
  GridParams<T> gp;
  const int32_t* cell_start;  // [M+1]
  const float4* rec;          // [n] cell-sorted records
  const int32_t* global_ids;  // optional local -> global id map (multi-GPU), else nullptr
  int32_t n_owned;            // rows are produced for local ids < n_owned
  int32_t* counts;            // [n_owned]
  const int64_t* offsets;     // [n_owned+1]   (FILL)
  int32_t* partners;          // [capacity]    (FILL)
  long long capacity;
  DeviceStatus* st;
  int32_t jt;                 // staged j records per shared-memory tile
};

// Stencil range along one axis: [c-1, c+1] clamped to the box.  The reference wraps the stencil periodically
// (neighlist_gpu.hpp:125-142) but measures distances without minimum image, so a wrapped cell can only contribute
// when it is also a direct neighbour, i.e. when the axis has exactly 3 cells — then all 3 cells are visited.
__device__ __forceinline__ void axis_range(int c, int m, int& lo, int& hi) {
  lo = max(c - 1, 0);
  hi = min(c + 1, m - 1);
  if (m == 3) {
    lo = 0;
    hi = 2;
  }
}

template <typename T, int STRIDE, bool HALF, bool FILL, bool EXACT_ONLY>
__global__ void __launch_bounds__(128) search_kernel(SearchArgs<T> a) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* sj = reinterpret_cast<float4*>(smem_raw);
  int32_t* sid = reinterpret_cast<int32_t*>(smem_raw + (size_t)a.jt * sizeof(float4));
  __shared__ int32_t r_start[9], r_b1[9], r_b2[9], r_pre[10], r_dy[9], r_dz[9];
  __shared__ int32_t s_xlo;

  const GridParams<T>& gp = a.gp;
  const int32_t cell = blockIdx.x;
  const int32_t ibeg = __ldg(a.cell_start + cell);
  const int32_t ni = __ldg(a.cell_start + cell + 1) - ibeg;
  if (ni == 0) return;
  if (FILL) {
    // capacity overflow was detected by the offsets scan: emit nothing (status already flagged)
    if (a.offsets[a.n_owned] > a.capacity) return;
  }
  const int32_t cx = cell % gp.mesh[0];
  const int32_t cy = (cell / gp.mesh[0]) % gp.mesh[1];
  const int32_t cz = cell / (gp.mesh[0] * gp.mesh[1]);

  if (threadIdx.x == 0) {
    int xlo, xhi, ylo, yhi, zlo, zhi;
    axis_range(cx, gp.mesh[0], xlo, xhi);
    axis_range(cy, gp.mesh[1], ylo, yhi);
    axis_range(cz, gp.mesh[2], zlo, zhi);
    s_xlo = xlo;
    int r = 0, pre = 0;
    for (int z = zlo; z <= zhi; z++)
      for (int y = ylo; y <= yhi; y++) {
        const int32_t row = (y + z * gp.mesh[1]) * gp.mesh[0];
        const int32_t s0 = a.cell_start[row + xlo];
        const int32_t s3 = a.cell_start[row + xhi + 1];
        r_start[r] = s0;
        r_b1[r] = (xlo + 1 <= xhi) ? a.cell_start[row + xlo + 1] : 0x7fffffff;
        r_b2[r] = (xlo + 2 <= xhi) ? a.cell_start[row + xlo + 2] : 0x7fffffff;
        r_dy[r] = y - cy;
        r_dz[r] = z - cz;
        r_pre[r] = pre;
        pre += s3 - s0;
        r++;
      }
    for (; r < 9; r++) {
      r_start[r] = 0;
      r_b1[r] = r_b2[r] = 0x7fffffff;
      r_dy[r] = r_dz[r] = 0;
      r_pre[r] = pre;
    }
    r_pre[9] = pre;
  }
  __syncthreads();
  const int32_t nj = r_pre[9];
  const int32_t xlo = s_xlo;
  const float hx = 0.5f * gp.msf[0], hy = 0.5f * gp.msf[1], hz = 0.5f * gp.msf[2];

  unsigned long long band_local = 0;

  for (int32_t ib = 0; ib < ni; ib += blockDim.x) {
    // ---- this thread's i particle ----
    const int32_t il = ib + threadIdx.x;
    bool active = il < ni;
    float xi = 0.f, yi = 0.f, zi = 0.f;
    int32_t iid = -1;
    if (active) {
      const float4 r = __ldg(a.rec + ibeg + il);
      xi = r.x - hx;
      yi = r.y - hy;
      zi = r.z - hz;
      iid = __float_as_int(r.w);
      active = iid < a.n_owned;
    }
    const int32_t icmp = (a.global_ids != nullptr && iid >= 0) ? __ldg(a.global_ids + iid) : iid;
    const float ai = 0.5f * (fmaf(xi, xi, fmaf(yi, yi, zi * zi)) - gp.sl2f);
    const float a_lo = active ? (ai - gp.band) : __int_as_float(0x7f800000);  // +inf: never hits
    const float a_hi = ai + gp.band;
    Vec3<T> qi_exact = {0, 0, 0};
    if (EXACT_ONLY && active) qi_exact = load_pos<T, STRIDE>(a.q, iid);

    int32_t cnt = 0;
    int32_t* wptr = nullptr;
    if (FILL && active) wptr = a.partners + a.offsets[iid];

    for (int32_t jb = 0; jb < nj; jb += a.jt) {
      const int32_t jn = min(a.jt, nj - jb);
      __syncthreads();  // previous tile fully consumed
      // ---- stage j records of this tile in CTA-local coordinates (origin = centre of the i cell) ----
      for (int32_t k = threadIdx.x; k < jn; k += blockDim.x) {
        const int32_t g = jb + k;
        int r = 0;
#pragma unroll
        for (int t = 1; t < 9; t++) r += (g >= r_pre[t]) ? 1 : 0;
        const int32_t slot = r_start[r] + (g - r_pre[r]);
        const int32_t dxc = xlo + ((slot >= r_b1[r]) ? 1 : 0) + ((slot >= r_b2[r]) ? 1 : 0) - cx;
        const float4 rj = __ldg(a.rec + slot);
        const float x = fmaf((float)dxc - 0.5f, gp.msf[0], rj.x);
        const float y = fmaf((float)r_dy[r] - 0.5f, gp.msf[1], rj.y);
        const float z = fmaf((float)r_dz[r] - 0.5f, gp.msf[2], rj.z);
        const float nb = -0.5f * fmaf(x, x, fmaf(y, y, z * z));
        sj[k] = make_float4(x, y, z, nb);
        sid[k] = __float_as_int(rj.w);
      }
      __syncthreads();
      // ---- test ----
#pragma unroll 4
      for (int32_t k = 0; k < jn; k++) {
        const float4 j = sj[k];
        const float t = fmaf(xi, j.x, fmaf(yi, j.y, fmaf(zi, j.z, j.w)));
        bool hit = EXACT_ONLY ? active : (t >= a_lo);
        if (hit) {
          const int32_t jid = sid[k];
          if (EXACT_ONLY) {
            hit = exact_within(qi_exact, load_pos<T, STRIDE>(a.q, jid), gp.sl2);
          } else if (t < a_hi) {
            // inside the pre-filter's uncertainty band: decide exactly, in the caller's precision
            hit = exact_within(load_pos<T, STRIDE>(a.q, iid), load_pos<T, STRIDE>(a.q, jid), gp.sl2);
            band_local++;
          }
          const int32_t jcmp = (a.global_ids != nullptr) ? __ldg(a.global_ids + jid) : jid;
          hit = hit && (HALF ? (jcmp > icmp) : (jid != iid));
          if (hit) {
            if (FILL) *wptr++ = jcmp;
            cnt++;
          }
        }
      }
    }
    if (!FILL && il < ni && iid >= 0 && iid < a.n_owned) a.counts[iid] = cnt;
  }
  if (!FILL) {
    if (threadIdx.x == 0) atomicAdd(&a.st->candidates, (unsigned long long)ni * (unsigned long long)nj);
    if (band_local) atomicAdd(&a.st->band_tests, band_local);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 5-7 (default path).  "pair masks": every distance test is evaluated ONCE per ordered pair, its verdict kept as one
// bit, and the CSR rows are expanded from the bits after the offsets are known.
//
//   pairmask_kernel   persistent warps draw (cell A, part) items from a queue.  The particles i of A sit in the warp's
//                     shared memory, pre-duplicated for packed math ({xi,xi,yi,yi}, {zi,zi,-ai,-ai},
//                     ai = (|xi|^2 - SL^2)/2, frame centred on A).  The candidates j — the <= 9 contiguous x-runs of A's
//                     stencil in the cell-sorted array — are spread over the LANES, PM_RJ per lane in registers,
//                     translated once into A's frame ({x, y, z, -|x|^2/2}).  Lane utilisation follows the ~950-long
//                     candidate list (>= 93 %) instead of the ~35 particles of a cell (55 % of two warps), and two
//                     broadcast LDS.128 feed PM_RJ tests (the shared-memory return path bounded the thread-per-i
//                     form, profiles/r01_microbench_issue_rates.txt).
//                     Test, dot form:  d = xi.xj - |xj|^2/2 - ai = (SL^2 - r^2)/2: 3 FFMA2 + 1 FADD2 per TWO tests;
//                     the sign bit is funnel-shifted into the lane's 32-bit word (1 SHF), min|d| tracked (FMNMX3).
//                     After 32 particles of A the lane holds, for ITS candidate j, the word "which i of A are within
//                     SL of j" — by symmetry a piece of ROW j.  Only a word whose min|d| fell inside the uncertainty
//                     band E is revisited (by the whole warp, lane = particle) with the exact input-precision test.
//                     mask[o][w][slot_j]: o = ordinal of A inside the stencil of j's cell, w = word (32 i's each).
//                     Candidate set-up is table-driven: per item the warp builds 9 int4 run entries and 27 float4
//                     (run, column) entries in shared memory; a lane keeps a cursor into the run table in registers,
//                     issues its PM_RJ record loads together and takes translation + mask plane from one 16-byte
//                     entry.  Integer divisions by mesh extents / parts are multiply-high (FastDiv).
//                     HALFIDS: the id filter of HALF lists is a suffix cut per (candidate, cell), found by a
//                     branch-free upper-bound search over the staged ids in shared memory, see the kernel.
//   rowcount_kernel   thread = row: popcount of the row's <= 27*WI words.
//   scan_kernel       counts -> offsets (as before).
//   emit_kernel       thread = row: expands set bits MSB-first (FLO) into its line of a per-warp shared-memory tile;
//                     when a line could overflow every lane flushes its own line to its row with 16-byte stores
//                     (scattered 4-byte stores cost more L2 sector writes than the whole test phase, and at 16 M
//                     particles they turned into DRAM read-modify-writes).
//   Rows come out in stencil order: cells ascending, ids ascending inside a cell — the discovery order of the
//   reference kernels (kernel_impl.cuh:17-33).
// ---------------------------------------------------------------------------------------------------------------
#ifndef NLB_PM_RJ
#define NLB_PM_RJ 8
#endif
constexpr int PM_RJ = NLB_PM_RJ;  // candidates per lane (packed in pairs)
#ifndef NLB_PM_THREADS
#define NLB_PM_THREADS 128
#endif
constexpr int PM_THREADS = NLB_PM_THREADS;  // warps are independent; the CTA only groups them
constexpr uint32_t FLAG_CELL_WORDS = 16u;  // a cell holds more than 32*WI particles: mask words too narrow

__device__ __forceinline__ int axis_lo(int c, int m) { return m == 3 ? 0 : max(c - 1, 0); }

// (SL^2 - r^2)/2 in FP32, fixed evaluation order: d = fma(xi, xj, fma(yi, yj, fma(zi, zj, wj))) + nai, nai = -ai.
// The hot loop evaluates two candidates per instruction with the packed forms (FFMA2 / FADD2: IEEE per element, so
// the scalar form below — used by the band re-test — reproduces the same bits).
__device__ __forceinline__ float pre_d(float xi, float yi, float zi, float nai, float xj, float yj, float zj,
                                       float wj) {
  return __fadd_rn(__fmaf_rn(xi, xj, __fmaf_rn(yi, yj, __fmaf_rn(zi, zj, wj))), nai);
}
typedef unsigned long long f32x2;  // two floats in one 64-bit register pair
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

template <typename T>
struct PairMaskArgs {
  const T* q;  // caller's positions (band re-test only)
  GridParams<T> gp;
  const int32_t* cell_start;
  const float4* rec;
  const int32_t* sorted_ids;
  int32_t n_owned;
  uint32_t* mask;      // [27][wi][n_cap]
  long long n_cap;
  int32_t wi;
  int32_t fits32;  // 27 * wi * n_cap < 2^32: mask element indices fit 32 bits
  FastDiv d_parts, d_mx, d_my;  // item -> (cell, part), cell -> (cx, cy, cz)
  float band;
  unsigned long long* queue;  // item counter (zeroed per build): warps draw (cell, part) items from it
  int32_t parts;   // items per cell: part p takes the candidate chunks p, p + parts, ...  (small systems: more
                   // items than resident warps, so that the queue can balance them)
  int32_t grab;    // items drawn per atomic (large systems: the single counter would otherwise serialise the warps)
  DeviceStatus* st;
};

// per-warp shared memory: the cell's particles + its run table
// per-warp run table: 9 int4 entries, one per x-run of the stencil, {end of the run in the candidate list, first slot
// minus start in the candidate list, first slot of the 2nd cell, first slot of the 3rd cell}; then 27 float4 entries,
// one per (run, x column): {tx, ty, tz} = translation of that cell into A's frame in cell units, .w = mask plane of A
// in the stencil of that cell (o * wi, as int bits)
constexpr int PM_TAB = 9 * 4 + 27 * 4;
// particles are stored pre-duplicated for the packed FMAs: {xi, xi, yi, yi}, {zi, zi, -ai, -ai}; at most PM_WC words
// (256 particles) of a cell are staged at a time, denser cells are walked in several rounds
constexpr int PM_WC = 8;
__host__ __device__ inline int pm_staged_words(int wi) { return wi < PM_WC ? wi : PM_WC; }
__host__ __device__ inline size_t pm_warp_bytes(int wi) {
  // staged particles {xi,xi,yi,yi},{zi,zi,-ai,-ai} + the tables + the staged particles' ids (HALF lists)
  return (size_t)pm_staged_words(wi) * 32 * (2 * sizeof(float4) + sizeof(int32_t)) + PM_TAB * 4;
}

#ifndef NLB_PM_MINB
#define NLB_PM_MINB 4
#endif
// HALFIDS: HALF lists without a global-id map.  Row j keeps the partners with a larger id (neighlist_cpu.hpp:225-236).
// The ids of a cell ascend with the slot, so inside one cell these partners are a SUFFIX: the first kept particle is
// found once per (candidate, cell) by a binary search and every word is cut with one AND — the filter costs nothing
// per test, the popcount pass and the emission then see HALF rows directly.
template <typename T, int STRIDE, bool HALFIDS>
__global__ void __launch_bounds__(PM_THREADS, NLB_PM_MINB) pairmask_kernel(PairMaskArgs<T> a) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  unsigned char* wbase = smem_raw + (size_t)warp * pm_warp_bytes(a.wi);
  float4* si = reinterpret_cast<float4*>(wbase);  // [32 * min(wi, PM_WC)][2]
  int4* t_run = reinterpret_cast<int4*>(wbase + (size_t)pm_staged_words(a.wi) * 32 * 2 * sizeof(float4));  // [9]
  float4* t_rc = reinterpret_cast<float4*>(t_run + 9);  // [27], index run * 3 + column
  int32_t* sid = reinterpret_cast<int32_t*>(t_rc + 27);  // [32 * staged words]: ids of the staged particles, ascending

  const GridParams<T>& gp = a.gp;
  const int32_t mx = gp.mesh[0], my = gp.mesh[1], mz = gp.mesh[2];
  const float msx = gp.msf[0], msy = gp.msf[1], msz = gp.msf[2];
  const float hx = 0.5f * msx, hy = 0.5f * msy, hz = 0.5f * msz;
  const uint32_t ncap32 = (uint32_t)a.n_cap;
  unsigned long long band_local = 0, cand_local = 0;

  const long long n_items = (long long)gp.n_cells * a.parts;
  // the first batch of every warp is static (no storm of atomics on one address at start-up); later batches come from
  // the queue, which therefore starts behind the static ones
  const long long n_warps = (long long)gridDim.x * (blockDim.x >> 5);
  const long long first_dyn = n_warps * a.grab;
  long long base = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * a.grab;
  while (base < n_items) {
    long long next = 0;
    if (lane == 0) next = first_dyn + (long long)atomicAdd(a.queue, (unsigned long long)a.grab);  // in flight meanwhile
    for (long long item = base; item < base + a.grab && item < n_items; item++) {
    // n_items < 2^31 (checked on the host): 32-bit fast division
    const int32_t cell = (int32_t)fdiv((uint32_t)item, a.d_parts), part = (int32_t)item - cell * a.parts;
    const int32_t ibeg = __ldg(a.cell_start + cell);
    int32_t ni = __ldg(a.cell_start + cell + 1) - ibeg;
    if (ni > 0) {
      if (ni > 32 * a.wi) {
        if (lane == 0 && part == 0) atomicOr(&a.st->flags, FLAG_CELL_WORDS);  // the build fails; stay in range
        ni = 32 * a.wi;
      }
      const int32_t cyz = (int32_t)fdiv((uint32_t)cell, a.d_mx);
      const int32_t cx = cell - cyz * mx;
      const int32_t cz = (int32_t)fdiv((uint32_t)cyz, a.d_my);
      const int32_t cy = cyz - cz * my;
      int xlo, xhi, ylo, yhi, zlo, zhi;
      axis_range(cx, mx, xlo, xhi);
      axis_range(cy, my, ylo, yhi);
      axis_range(cz, mz, zlo, zhi);
      const int32_t ny = yhi - ylo + 1, nruns = ny * (zhi - zlo + 1);
      __syncwarp();  // the previous cell's readers are done
      int32_t nj;  // candidates of this cell = particles of its stencil
      // the first two words of the cell's own particles are requested now, next to the run table's cell starts: one
      // round trip instead of two before the first candidate can be set up
      float4 pre_i[2];
#pragma unroll
      for (int u = 0; u < 2; u++) pre_i[u] = __ldg(a.rec + ibeg + min(u * 32 + lane, ni - 1));
      {
        // run table: lane r describes run r = (z, y); prefix of the run lengths by warp scan
        int32_t len = 0, s0 = 0, b1 = 0x7fffffff, b2 = 0x7fffffff, o = 0;
        float ty = 0.f, tz = 0.f;
        if (lane < nruns) {
          const int lz = ny == 3 ? (lane * 11) >> 5 : (ny == 2 ? lane >> 1 : lane);  // lane / ny for lane < 9
          const int z = zlo + lz, y = ylo + lane - lz * ny;
          const int32_t* cs = a.cell_start + (y + z * my) * mx;
          s0 = __ldg(cs + xlo);
          if (xlo + 1 <= xhi) b1 = __ldg(cs + xlo + 1);
          if (xlo + 2 <= xhi) b2 = __ldg(cs + xlo + 2);
          len = __ldg(cs + xhi + 1) - s0;
          ty = (float)(y - cy) - 0.5f;
          tz = (float)(z - cz) - 0.5f;
          // ordinal of this cell inside the stencil of the candidate's cell (x part added per candidate)
          o = ((cz - axis_lo(z, mz)) * 3 + (cy - axis_lo(y, my))) * 3;
        }
        int32_t incl = len;
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) {
          const int32_t v = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += v;
        }
        if (lane < 9) t_run[lane] = make_int4(incl, s0 - (incl - len), b1, b2);  // absent runs: len 0, end = total
        nj = __shfl_sync(0xffffffffu, incl, 8);
        // lane e = run * 3 + column writes that stencil cell's entry: everything a candidate's set-up needs from its
        // cell comes from ONE 16-byte shared-memory load instead of six look-ups and the selects between them
        const int er = lane / 3, ecol = lane - er * 3;
        const float ty_e = __shfl_sync(0xffffffffu, ty, er & 15), tz_e = __shfl_sync(0xffffffffu, tz, er & 15);
        const int32_t o_e = __shfl_sync(0xffffffffu, o, er & 15);
        if (lane < 27) {
          // ordinal (x part) of A inside the stencil of a candidate in column xlo + ecol
          const int32_t ox = cx - axis_lo(min(xlo + ecol, mx - 1), mx);
          t_rc[lane] = make_float4((float)(xlo + ecol - cx) - 0.5f, ty_e, tz_e, __int_as_float((o_e + ox) * a.wi));
        }
      }
      __syncwarp();
      if (lane == 0 && part == 0) cand_local += (unsigned long long)ni * (unsigned long long)nj;

      for (int32_t iw0 = 0; iw0 * 32 < ni; iw0 += PM_WC) {  // one round for cells of up to 256 particles
      const int32_t iw1 = min(iw0 + PM_WC, (ni + 31) >> 5);
      if (iw0 > 0) __syncwarp();  // the previous round's readers are done
      for (int32_t k = iw0 * 32 + lane; k < min(ni, iw1 * 32); k += 32) {
        const float4 r = k < 32 ? pre_i[0] : (k < 64 ? pre_i[1] : __ldg(a.rec + ibeg + k));
        const float x = r.x - hx, y = r.y - hy, z = r.z - hz;
        const float nai = -0.5f * (fmaf(x, x, fmaf(y, y, z * z)) - gp.sl2f);
        si[2 * (k - iw0 * 32)] = make_float4(x, x, y, y);
        si[2 * (k - iw0 * 32) + 1] = make_float4(z, z, nai, nai);
        if (HALFIDS) sid[k - iw0 * 32] = __float_as_int(r.w);
      }
      __syncwarp();
      // cursor into the run table: a lane's candidates ascend over the chunks, so the run only moves forward and the
      // table is read again (one 16-byte load) only when a candidate crosses into the next run
      int32_t run = 0;
      int4 cur = t_run[0];
      for (int32_t c0 = part * (32 * PM_RJ); c0 < nj; c0 += a.parts * (32 * PM_RJ)) {
        float xj[PM_RJ], yj[PM_RJ], zj[PM_RJ], wj[PM_RJ];
        int32_t sj[PM_RJ];  // candidate's slot
        int32_t oj[PM_RJ];  // its mask plane for this cell (o * wi); -1: tail lane or ghost row, nothing to store
        int32_t pj[PM_RJ];  // HALFIDS: particles of this cell with an id <= the candidate's (they are not kept)
        // three passes so that the PM_RJ record loads are in flight together: slots first (run table only), then
        // every load, then the translation — the run look-up between two loads used to serialise them into PM_RJ
        // round trips per chunk
        int32_t rcol[PM_RJ];  // run * 3 + column; -1: tail lane
#pragma unroll
        for (int k = 0; k < PM_RJ; k++) {
          const int32_t c = c0 + k * 32 + lane;
          sj[k] = 0;
          rcol[k] = -1;
          if (c < nj) {
            while (c >= cur.x) cur = t_run[++run];  // c < nj = end of run 8: stops at run <= 8
            const int32_t s = cur.y + c;
            sj[k] = s;
            rcol[k] = run * 3 + ((s >= cur.z) ? 1 : 0) + ((s >= cur.w) ? 1 : 0);
          }
        }
        float4 rjv[PM_RJ];
#pragma unroll
        for (int k = 0; k < PM_RJ; k++) rjv[k] = __ldg(a.rec + sj[k]);  // tail lanes read slot 0 (present: ni > 0)
#pragma unroll
        for (int k = 0; k < PM_RJ; k++) {
          xj[k] = yj[k] = zj[k] = 0.f;
          wj[k] = -1.0e30f;  // d = -1e30: a miss, far from the band
          oj[k] = -1;
          pj[k] = 0;
          if (rcol[k] >= 0) {
            const float4 rj = rjv[k];
            const float4 tr = t_rc[rcol[k]];
            xj[k] = fmaf(tr.x, msx, rj.x);
            yj[k] = fmaf(tr.y, msy, rj.y);
            zj[k] = fmaf(tr.z, msz, rj.z);
            wj[k] = -0.5f * fmaf(xj[k], xj[k], fmaf(yj[k], yj[k], zj[k] * zj[k]));
            // rec.w carries the particle's local id (cellsort_kernel): rows of ghosts (id >= n_owned) are not stored
            const int32_t idj = __float_as_int(rj.w);
            if (idj < a.n_owned) oj[k] = __float_as_int(tr.w);
          }
        }
        {
          // a chunk whose candidates are all ghosts (the outer cell layers of a slab rank) produces no row: skip it
          bool row_needed = false;
#pragma unroll
          for (int k = 0; k < PM_RJ; k++) row_needed = row_needed || oj[k] >= 0;
          if (!__any_sync(0xffffffffu, row_needed)) continue;
        }
        if (HALFIDS) {
          // pj = number of this cell's particles with an id <= the candidate's: upper bound in the ascending ids of the
          // staged particles (shared memory), branch-free and with the PM_RJ searches of a lane interleaved step by
          // step — a binary search per candidate in global memory cost a third of the kernel
          const int32_t nwin = min(ni, iw1 * 32) - iw0 * 32;
          int32_t pos[PM_RJ];
#pragma unroll
          for (int k = 0; k < PM_RJ; k++) pos[k] = 0;
          for (int32_t step = 1 << (31 - __clz(nwin)); step > 0; step >>= 1) {
#pragma unroll
            for (int k = 0; k < PM_RJ; k++) {
              const int32_t t = pos[k] + step;
              if (t <= nwin && sid[min(t, nwin) - 1] <= __float_as_int(rjv[k].w)) pos[k] = t;
            }
          }
#pragma unroll
          for (int k = 0; k < PM_RJ; k++) pj[k] = iw0 * 32 + pos[k];
        }
        f32x2 X[PM_RJ / 2], Y[PM_RJ / 2], Z[PM_RJ / 2], W[PM_RJ / 2];
#pragma unroll
        for (int h = 0; h < PM_RJ / 2; h++) {
          X[h] = pack2(xj[2 * h], xj[2 * h + 1]);
          Y[h] = pack2(yj[2 * h], yj[2 * h + 1]);
          Z[h] = pack2(zj[2 * h], zj[2 * h + 1]);
          W[h] = pack2(wj[2 * h], wj[2 * h + 1]);
        }
        for (int32_t w = iw0; w < iw1; w++) {
          const int32_t cnt = min(32, ni - w * 32);
          const ulonglong2* sp = reinterpret_cast<const ulonglong2*>(si + (w - iw0) * 64);
          uint32_t miss[PM_RJ];
#pragma unroll
          for (int k = 0; k < PM_RJ; k++) miss[k] = 0u;
          float mh[PM_RJ / 2];  // min |d| per candidate pair
#pragma unroll
          for (int h = 0; h < PM_RJ / 2; h++) mh[h] = 3.0e38f;
#pragma unroll 2
          for (int32_t ii = 0; ii < cnt; ii++) {
            const ulonglong2 p0 = sp[2 * ii];      // {xi, xi}, {yi, yi}
            const ulonglong2 p1 = sp[2 * ii + 1];  // {zi, zi}, {-ai, -ai}
#pragma unroll
            for (int h = 0; h < PM_RJ / 2; h++) {
              const f32x2 d2 = add2(fma2(p0.x, X[h], fma2(p0.y, Y[h], fma2(p1.x, Z[h], W[h]))), p1.y);
              float d0, d1;
              unpack2(d2, d0, d1);
              miss[2 * h] = __funnelshift_l(__float_as_uint(d0), miss[2 * h], 1);  // shift the sign bit in
              miss[2 * h + 1] = __funnelshift_l(__float_as_uint(d1), miss[2 * h + 1], 1);
              mh[h] = fminf(mh[h], fminf(fabsf(d0), fabsf(d1)));
            }
          }
          uint32_t hits[PM_RJ];
#pragma unroll
          for (int k = 0; k < PM_RJ; k++) hits[k] = (~miss[k]) << (32 - cnt);  // bit (31 - ii) <-> particle w*32+ii
          // Tests that fell inside the pre-filter's uncertainty band are decided exactly, in the caller's precision.
          // Rare per test (~2e-5) but not per word (8192 tests): the re-check is done by the whole warp — lane = particle
          // ii, the triggering lane's candidate broadcast by shuffles — instead of one lane looping alone.
          float mall = mh[0];
#pragma unroll
          for (int h = 1; h < PM_RJ / 2; h++) mall = fminf(mall, mh[h]);
          unsigned trig = __ballot_sync(0xffffffffu, mall < a.band);
          while (trig) {
            const int src = __ffs(trig) - 1;
            trig &= trig - 1;
#pragma unroll
            for (int h = 0; h < PM_RJ / 2; h++) {
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
                const int32_t s_src = __shfl_sync(0xffffffffu, sj[k], src);
                const int32_t o_src = __shfl_sync(0xffffffffu, oj[k], src);
                bool fix = false, hit = false;
                if (lane < cnt && o_src >= 0) {
                  const float4 q0 = si[(w - iw0) * 64 + 2 * lane], q1 = si[(w - iw0) * 64 + 2 * lane + 1];
                  const float d = pre_d(q0.x, q0.z, q1.x, q1.z, cx2[e], cy2[e], cz2[e], cw2[e]);
                  if (fabsf(d) < a.band) {
                    const int32_t iid = __ldg(a.sorted_ids + ibeg + w * 32 + lane);
                    const int32_t jid = __ldg(a.sorted_ids + s_src);
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
          if (HALFIDS) {
#pragma unroll
            for (int k = 0; k < PM_RJ; k++) {
              const int32_t cut = pj[k] - w * 32;  // particles w*32 .. w*32 + cut - 1 have an id <= the candidate's
              hits[k] = cut <= 0 ? hits[k] : (cut >= 32 ? 0u : (hits[k] & (0xffffffffu >> cut)));
            }
          }
          // element index (plane * n_cap + slot): one 32 x 32 -> 64-bit multiply-add (n_cap < 2^31), or plain 32-bit
          // arithmetic when the whole mask has fewer than 2^32 words
          if (a.fits32) {
#pragma unroll
            for (int k = 0; k < PM_RJ; k++)
              if (oj[k] >= 0) a.mask[(uint32_t)(oj[k] + w) * ncap32 + (uint32_t)sj[k]] = hits[k];
          } else {
#pragma unroll
            for (int k = 0; k < PM_RJ; k++)
              if (oj[k] >= 0)
                a.mask[(unsigned long long)(uint32_t)(oj[k] + w) * ncap32 + (uint32_t)sj[k]] = hits[k];
          }
        }
      }
      }
    }
    }
    base = __shfl_sync(0xffffffffu, next, 0);
  }
  if (cand_local) atomicAdd(&a.st->candidates, cand_local);
  if (band_local) atomicAdd(&a.st->band_tests, band_local);
}

struct EmitArgs {
  const int32_t* cell_start;
  const int32_t* sorted_ids;
  const int32_t* slot_cell;
  const int32_t* global_ids;  // optional local -> global id map (HALF rule of the row's own particle)
  const int32_t* slot_pid;    // partner id reported for a slot: sorted_ids, or the global ids in slot order
  int32_t mesh[3];
  FastDiv d_mx, d_my;  // cell -> (bx, by, bz) without hardware divisions
  int32_t n_total, n_owned;
  int32_t n_cells;  // cell_start[n_cells] = particles present (n_total minus absent ghosts) = slots in use
  int32_t clear_self;  // 1: a row's own bit is set in the masks and has to be dropped (FULL lists)
  const uint32_t* mask;
  long long n_cap;
  int32_t wi;
  int32_t* counts;
  const int64_t* offsets;
  int32_t* partners;
  long long capacity;
};

// Visits the mask words of the row held in `slot` (a particle of cell `cell`) in stencil order — cells ascending,
// words ascending inside a cell — and calls f(word, first_slot): bit (31 - b) of `word` set <=> the particle in
// cell-sorted slot first_slot + b is within SL.  The first two words of the three cells of an x-run (all of them
// unless a cell holds more than 64 particles) are fetched by six independent loads before any is used.
// CLEAR_SELF removes the row's own bit (FULL lists: j != i, kernel_impl.cuh:29).
template <bool CLEAR_SELF, typename F>
__device__ __forceinline__ void walk_words(const EmitArgs& a, int32_t slot, int32_t cell, F&& f) {
  const int32_t mx = a.mesh[0], my = a.mesh[1], mz = a.mesh[2];
  const int32_t bx = cell % mx, by = (cell / mx) % my, bz = cell / (mx * my);
  int xlo, xhi, ylo, yhi, zlo, zhi;
  axis_range(bx, mx, xlo, xhi);
  axis_range(by, my, ylo, yhi);
  axis_range(bz, mz, zlo, zhi);
  const int32_t nx = xhi - xlo + 1;
  const int32_t own = slot - __ldg(a.cell_start + cell);  // index inside the own cell
  const uint32_t* mrow = a.mask + slot;
  for (int z = zlo; z <= zhi; z++)
    for (int y = ylo; y <= yhi; y++) {
      const int32_t* cs = a.cell_start + (y + z * my) * mx + xlo;
      int32_t cb[4];
#pragma unroll
      for (int k = 0; k < 4; k++) cb[k] = (k <= nx) ? __ldg(cs + k) : 0;
      const int32_t o0 = ((z - zlo) * 3 + (y - ylo)) * 3 * a.wi;
      int32_t nw[3];
      uint32_t pre[3][2];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        nw[k] = (k < nx) ? min((cb[k + 1] - cb[k] + 31) >> 5, a.wi) : 0;  // > wi only after FLAG_CELL_WORDS
#pragma unroll
        for (int u = 0; u < 2; u++)
          pre[k][u] = (u < nw[k]) ? __ldg(mrow + (long long)(o0 + k * a.wi + u) * a.n_cap) : 0u;
      }
      const bool own_row = CLEAR_SELF && (z == bz) && (y == by);
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const bool own_cell = own_row && (xlo + k == bx);
#pragma unroll
        for (int u = 0; u < 2; u++) {
          uint32_t word = pre[k][u];
          if (own_cell && (own >> 5) == u) word &= ~(0x80000000u >> (own & 31));
          if (word) f(word, cb[k] + 32 * u);
        }
        for (int32_t w = 2; w < nw[k]; w++) {
          uint32_t word = __ldg(mrow + (long long)(o0 + k * a.wi + w) * a.n_cap);
          if (own_cell && (own >> 5) == w) word &= ~(0x80000000u >> (own & 31));
          if (word) f(word, cb[k] + 32 * w);
        }
      }
    }
}

// FULL lists: the row length is a popcount of the row's mask words (without its own bit).  Thread = row.  All loads
// of a stencil plane (12 cell starts + 18 words) are issued before any is used and the word loads do not wait for the
// cell starts.  The kernel is bound by its own index arithmetic (it issues on half of the cycles for 26 MB of L2
// reads), which is why the loads are unconditional and the plane pointer is a running 64-bit add.
__global__ void __launch_bounds__(128) rowcount_kernel(EmitArgs a) {
  pdl_enter();
  const int32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= a.n_total || slot >= __ldg(a.cell_start + a.n_cells)) return;
  const int32_t id = __ldg(a.sorted_ids + slot);
  if (id >= a.n_owned) return;
  const int32_t cell = __ldg(a.slot_cell + slot);
  const int32_t mx = a.mesh[0], my = a.mesh[1], mz = a.mesh[2];
  const int32_t byz = (int32_t)fdiv((uint32_t)cell, a.d_mx), bx = cell - byz * mx;
  const int32_t bz = (int32_t)fdiv((uint32_t)byz, a.d_my), by = byz - bz * my;
  int xlo, xhi, ylo, yhi, zlo, zhi;
  axis_range(bx, mx, xlo, xhi);
  axis_range(by, my, ylo, yhi);
  axis_range(bz, mz, zlo, zhi);
  const int32_t nx = xhi - xlo + 1, ny = yhi - ylo + 1, nz = zhi - zlo + 1;
  const uint32_t* mrow = a.mask + slot;
  const long long cell_stride = (long long)a.wi * a.n_cap;  // words between the planes of two stencil cells
  const long long word1 = a.wi >= 2 ? a.n_cap : 0;           // second word of a cell (wi = 1: the first one again)
  const uint32_t* mcell = mrow;  // walks the stencil cells' planes in order
  int32_t cnt = 0;
  for (int oz = 0; oz < nz; oz++) {
    // unconditional loads at clamped addresses (as in emit_kernel): cells the row does not have are dropped by nw = 0
    int32_t cbp[3][4];
    uint32_t mp[3][3][2];
#pragma unroll
    for (int oy = 0; oy < 3; oy++) {
      const int32_t* cs = a.cell_start + (min(ylo + oy, my - 1) + (zlo + oz) * my) * mx + xlo;
#pragma unroll
      for (int k = 0; k < 4; k++) cbp[oy][k] = __ldg(cs + min(k, nx));
#pragma unroll
      for (int k = 0; k < 3; k++) {
        mp[oy][k][0] = __ldg(mcell);
        mp[oy][k][1] = __ldg(mcell + word1);
        mcell += cell_stride;
      }
    }
#pragma unroll
    for (int oy = 0; oy < 3; oy++) {
      const int32_t o0 = (oz * 3 + oy) * 3 * a.wi;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const int32_t nw = (oy < ny && k < nx) ? min((cbp[oy][k + 1] - cbp[oy][k] + 31) >> 5, a.wi) : 0;
        if (nw > 0) cnt += __popc(mp[oy][k][0]);
        if (nw > 1) cnt += __popc(mp[oy][k][1]);
        for (int32_t w = 2; w < nw; w++) cnt += __popc(__ldg(mrow + (long long)(o0 + k * a.wi + w) * a.n_cap));
      }
    }
  }
  if (a.clear_self) {
    // the row's own bit (r2 = 0 passes the test unless the record is NaN): FULL rows hold j != i (kernel_impl.cuh:29)
    const int32_t own = slot - __ldg(a.cell_start + cell);
    if ((own >> 5) < a.wi) {
      const int32_t o = ((bz - zlo) * 3 + (by - ylo)) * 3 + (bx - xlo);
      const uint32_t word = __ldg(mrow + (long long)(o * a.wi + (own >> 5)) * a.n_cap);
      cnt -= (int32_t)((word >> (31 - (own & 31))) & 1u);
    }
  }
  a.counts[id] = cnt;
}

#ifdef NLB_ABLATIONS
// emit_direct_kernel (ablation, NLB200_OPT_KERNEL_VARIANT = 3 of a -DNLB_ABLATIONS build): thread = row, every hit stored straight to
// partners[offsets[id] + k] — a warp store touches 32 rows, 32 single-word partial-sector writes.  Measured on B200:
// 144 us vs 101 us staged on the default system, 5.9 ms vs 1.7 ms at 2 M uniform particles.
template <bool HALF, bool GID, bool COUNT>
__global__ void __launch_bounds__(128) emit_direct_kernel(EmitArgs a) {
  pdl_enter();
  if (!COUNT) {
    if (a.offsets[a.n_owned] > a.capacity) return;
  }
  const int32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= a.n_total || slot >= __ldg(a.cell_start + a.n_cells)) return;
  const int32_t id = __ldg(a.sorted_ids + slot);
  if (id >= a.n_owned) return;
  const int32_t cell = __ldg(a.slot_cell + slot);
  const int32_t mycmp = GID ? __ldg(a.global_ids + id) : id;
  int32_t* wp = COUNT ? nullptr : a.partners + a.offsets[id];
  int32_t cnt = 0;
  walk_words<!HALF>(a, slot, cell, [&](uint32_t word, int32_t first) {
    const int32_t* ids = a.slot_pid + first;
    while (word) {
      const int b = __clz(word);
      word &= ~(0x80000000u >> b);
      const int32_t pid = __ldg(ids + b);
      if (HALF && !(pid > mycmp)) continue;
      if (!COUNT) wp[cnt] = pid;
      cnt++;
    }
  });
  if (COUNT) a.counts[id] = cnt;
}

#endif  // NLB_ABLATIONS

// emit_kernel: thread = row, warp = 32 consecutive cell-sorted slots; warps are independent (no CTA barrier).
//   Each lane expands the set bits of its row's words, MSB first (one FLO per bit), into its line of a
//   [32][EM_TILE] shared-memory tile: the cell-sorted SLOTS of its partners, cell by cell in stencil order.  When a
//   line could overflow (checked per cell), every lane flushes ITS OWN line to its row: slots -> partner ids (gather
//   from sorted_ids / the per-slot global ids), scalar stores until the row position is 16-byte aligned, then 16-byte
//   vector stores at partners[offsets[id] + done ...]; the <= 3 entries that do not fill a vector stay in the line.
//   7 KB of tile per warp instead of whole rows keeps 28 warps per SM resident (72 registers), which is all the
//   25 warps of rows an SM gets on the default system.
// Rejected (measured): one warp per (32 rows, stencil plane) — three times the warps, a third of the serial chain per
// lane — emits the default system in 77.7 us vs 75.9 us: the kernel is not short of parallelism, its issue slots, XU
// (FLO) and LSU wavefronts are each ~40-50 % busy.
// HALF:  rows keep the partners with a larger (global) id (neighlist_cpu.hpp:225-236); the filter runs in the flush
//        (scalar stores of the kept ids).  COUNT: write counts[id] instead of partners (HALF lists need the ids to count).
#ifndef NLB_EM_WARPS
#define NLB_EM_WARPS 2
#endif
constexpr int EM_WARPS = NLB_EM_WARPS;
#ifndef NLB_EM_TILE
#define NLB_EM_TILE 56
#endif
constexpr int EM_TILE = NLB_EM_TILE;
constexpr int EM_LINE = EM_TILE + 1;  // +1: lanes with equal fill hit different banks
#ifndef NLB_EM_MINB
#define NLB_EM_MINB 14
#endif

template <bool HALF, bool GID, bool COUNT>
__global__ void __launch_bounds__(EM_WARPS * 32, NLB_EM_MINB) emit_kernel(EmitArgs a) {
  pdl_enter();
  extern __shared__ __align__(16) int32_t em_smem[];
  if (!COUNT) {
    if (a.offsets[a.n_owned] > a.capacity) return;  // overflow already flagged by the offsets scan
  }
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  int32_t* tile = em_smem + warp * 32 * EM_LINE;
  int32_t* line = tile + lane * EM_LINE;
  const uint32_t line_sa = (uint32_t)__cvta_generic_to_shared(line);
  const int32_t slot = (blockIdx.x * EM_WARPS + warp) * 32 + lane;
  int32_t id = 0x7fffffff;
  if (slot < a.n_total && slot < __ldg(a.cell_start + a.n_cells)) id = __ldg(a.sorted_ids + slot);
  const bool owned = id < a.n_owned;
  const int32_t cell = owned ? __ldg(a.slot_cell + slot) : 0;
  const int32_t rcmp = (HALF && owned) ? (GID ? __ldg(a.global_ids + id) : id) : 0;
  const long long dst = (!COUNT && owned) ? (long long)a.offsets[id] : 0;
  int32_t fill = 0;  // entries staged in this lane's line
  int32_t done = 0;  // entries of this row already written (HALF: after the id filter)

  // flush(final): every lane copies ITS OWN line to its row.  Staged slots -> partner ids (gather), then 16-byte
  // vector stores once the row position is 16-byte aligned: a warp store writes 32 half sectors instead of 32 single
  // words, and there is no per-row serial loop (all 32 rows move in parallel).  Entries that do not fill a vector
  // stay at the front of the line for the next flush; `final` writes them out.
  auto flush = [&](bool final) {
    int32_t k = 0;
    if (!HALF) {
      int32_t* out = a.partners + dst + done;
      // head: scalar stores until the row position is 16-byte aligned
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
    } else {
      // HALF: the j > i filter compacts the stream, so the kept ids are stored one by one
      int32_t* out = a.partners + dst;
      for (; k < fill; k++) {
        const int32_t pid = __ldg(a.slot_pid + line[k]);
        if (pid > rcmp) {
          if (!COUNT) out[done] = pid;
          done++;
        }
      }
    }
    // keep the (at most 3) entries that did not fill a vector
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
  const int32_t nx = owned ? xhi - xlo + 1 : 0, ny = yhi - ylo + 1, nz = zhi - zlo + 1;
  const int32_t own = owned ? slot - __ldg(a.cell_start + cell) : 0;
  const uint32_t* mrow = a.mask + min((long long)slot, a.n_cap - 1);  // lanes past the last slot load in range
  const long long cell_stride = (long long)a.wi * a.n_cap;  // words between the planes of two stencil cells
  const long long word1 = a.wi >= 2 ? a.n_cap : 0;           // second word of a cell (wi = 1: the first one again)

  // the plane / run loops are warp-uniform (3 x 3 stencil ordinals); lanes without that run see empty words
  const uint32_t* mcell = mrow;  // walks the 27 stencil cells' planes in order
  for (int oz = 0; oz < 3; oz++) {
    // all loads of the plane (3 runs x (4 cell starts + 6 words)) are issued before any is used; the word loads do
    // not wait for the cell starts (words beyond a cell's population are discarded afterwards).  Every load is
    // unconditional with an address clamped into the arrays — stencil cells a boundary row does not have read
    // planes nobody wrote, and their words are dropped by nw = 0 below — and the mask pointer advances by one
    // 64-bit add per cell: the predicated loads with a 64-bit multiply each were a tenth of the kernel's instructions
    int32_t cbp[3][4];
    uint32_t mp[3][3][2];
    const int32_t zz = min(zlo + oz, mz - 1);
#pragma unroll
    for (int oy = 0; oy < 3; oy++) {
      const int32_t* cs = a.cell_start + (min(ylo + oy, my - 1) + zz * my) * mx + xlo;
#pragma unroll
      for (int k = 0; k < 4; k++) cbp[oy][k] = __ldg(cs + min(k, nx));
#pragma unroll
      for (int k = 0; k < 3; k++) {
        mp[oy][k][0] = __ldg(mcell);
        mp[oy][k][1] = __ldg(mcell + word1);
        mcell += cell_stride;
      }
    }
#pragma unroll
    for (int oy = 0; oy < 3; oy++) {
      const bool rv = (nx > 0) && (oz < nz) && (oy < ny);
      const int32_t o0 = (oz * 3 + oy) * 3 * a.wi;
      int32_t cb[4], nw[3];
      uint32_t m[3][2];
#pragma unroll
      for (int k = 0; k < 4; k++) cb[k] = cbp[oy][k];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        nw[k] = (rv && k < nx) ? min((cb[k + 1] - cb[k] + 31) >> 5, a.wi) : 0;  // > wi only after FLAG_CELL_WORDS
        m[k][0] = nw[k] > 0 ? mp[oy][k][0] : 0u;
        m[k][1] = nw[k] > 1 ? mp[oy][k][1] : 0u;
      }
      const bool own_run = !HALF && a.clear_self && rv && (zlo + oz == bz) && (ylo + oy == by);  // FULL: j != i
      if (own_run) {
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
          for (int u = 0; u < 2; u++)
            if (xlo + k == bx && (own >> 5) == u) m[k][u] &= ~(0x80000000u >> (own & 31));
      }
      int32_t cell_hits[3];
#pragma unroll
      for (int k = 0; k < 3; k++) cell_hits[k] = __popc(m[k][0]) + __popc(m[k][1]);
      const bool slow_lane = max(nw[0], max(nw[1], nw[2])) > 2 ||
                             max(cell_hits[0], max(cell_hits[1], cell_hits[2])) > EM_TILE - 3;
      if (!__any_sync(0xffffffffu, slow_lane)) {
        // stencil order: cell 0 (words 0, 1), cell 1, cell 2; lanes of a warp sit in the same or adjacent cells, so
        // their popcounts of one word are alike and the per-word loops stay reasonably full.  The flush check runs
        // per cell: the central run of a row holds up to ~65 partners, more than a line, but no single cell does.
#pragma unroll
        for (int k = 0; k < 3; k++) {
          if (__any_sync(0xffffffffu, fill + cell_hits[k] > EM_TILE)) flush(false);  // leaves fill <= 3
          uint32_t wa = line_sa + 4u * (uint32_t)fill;  // shared-memory byte address of the next free entry
          fill += cell_hits[k];
#pragma unroll
          for (int u = 0; u < 2; u++) {
            uint32_t word = m[k][u];
            const int32_t last = cb[k] + 32 * u + 31;  // slot of bit 0
            while (word) {
              uint32_t p;  // position of the highest set bit: one FLO, no 31 - clz
              asm("bfind.u32 %0, %1;" : "=r"(p) : "r"(word));
              word ^= 1u << p;
              asm volatile("st.shared.s32 [%0], %1;" ::"r"(wa), "r"(last - (int32_t)p) : "memory");
              wa += 4u;
            }
          }
        }
      } else {
        // a cell of this run holds more than 64 particles (or one cell alone overflows a line): word by word,
        // warp-uniform loop bounds, a flush check before every word
        for (int k = 0; k < 3; k++)
          for (int32_t w = 0; w < a.wi; w++) {
            const int32_t nwk = k == 0 ? nw[0] : (k == 1 ? nw[1] : nw[2]);
            uint32_t word = 0u;
            if (w < nwk) {
              word = __ldg(mrow + (long long)(o0 + k * a.wi + w) * a.n_cap);
              if (own_run && xlo + k == bx && (own >> 5) == w) word &= ~(0x80000000u >> (own & 31));
            }
            if (!__any_sync(0xffffffffu, word != 0u)) continue;
            if (__any_sync(0xffffffffu, fill + __popc(word) > EM_TILE)) flush(false);
            const int32_t first = (k == 0 ? cb[0] : (k == 1 ? cb[1] : cb[2])) + 32 * w;
            while (word) {
              const int b = __clz(word);
              word &= ~(0x80000000u >> b);
              line[fill++] = first + b;
            }
          }
      }
    }
  }
  flush(true);
  if (COUNT && owned) a.counts[id] = done;
}

// ---------------------------------------------------------------------------------------------------------------
// 8. optional: rows ascending (what the reference tests do on the host before comparing, make_list.cpp:120-128).
//    One warp per row; bitonic sort in shared memory for rows <= SORT_SMEM, in global memory beyond.
// ---------------------------------------------------------------------------------------------------------------
constexpr int SORT_SMEM = 1024;
constexpr int SORT_WARPS = 4;

// Ascending-only bitonic network ("flip" formulation): every comparator leaves the minimum at the lower index, so a
// virtual +inf padding above `len` never moves and comparators that touch it can simply be skipped — any length
// works in place, in shared or global memory.
__device__ __forceinline__ void bitonic_warp(int32_t* buf, int len, int lane) {
  int p2 = 1;
  while (p2 < len) p2 <<= 1;
  for (int k = 2; k <= p2; k <<= 1) {
    for (int t = lane; t < len; t += 32) {
      const int p = t ^ (k - 1);
      if (p > t && p < len) {
        const int32_t x = buf[t], y = buf[p];
        if (x > y) {
          buf[t] = y;
          buf[p] = x;
        }
      }
    }
    __syncwarp();
    for (int j = k >> 2; j > 0; j >>= 1) {
      for (int t = lane; t < len; t += 32) {
        const int p = t ^ j;
        if (p > t && p < len) {
          const int32_t x = buf[t], y = buf[p];
          if (x > y) {
            buf[t] = y;
            buf[p] = x;
          }
        }
      }
      __syncwarp();
    }
  }
}

__global__ void __launch_bounds__(SORT_WARPS * 32) sort_rows_kernel(const int64_t* __restrict__ offsets,
                                                                    int32_t n_rows, int32_t* partners,
                                                                    long long capacity) {
  pdl_enter();
  __shared__ int32_t sbuf[SORT_WARPS][SORT_SMEM];
  if (offsets[n_rows] > capacity) return;
  const int lane = lane_id();
  const int w = threadIdx.x >> 5;
  const int32_t row = blockIdx.x * SORT_WARPS + w;
  if (row >= n_rows) return;
  const int64_t beg = offsets[row];
  const int32_t len = (int32_t)(offsets[row + 1] - beg);
  if (len <= 1) return;
  int32_t* g = partners + beg;
  if (len <= SORT_SMEM) {
    int32_t* buf = sbuf[w];
    for (int t = lane; t < len; t += 32) buf[t] = g[t];
    __syncwarp();
    bitonic_warp(buf, len, lane);
    for (int t = lane; t < len; t += 32) g[t] = buf[t];
  } else {
    // rows this long only occur in the clustered stress configuration: same network, in place in global memory
    // (__syncwarp orders the lanes' global accesses between stages)
    bitonic_warp(g, len, lane);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 9. optional: the reference GPU layout list[k*n + i], -1 padded (kernel_impl.cuh:30, neighlist_gpu.hpp:271-274)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ell_kernel(const int64_t* __restrict__ offsets,
                                                  const int32_t* __restrict__ partners, int32_t n, int32_t rows,
                                                  int32_t* __restrict__ ell, int32_t* __restrict__ prev_count,
                                                  long long capacity, DeviceStatus* st) {
  pdl_enter();
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (offsets[n] > capacity) return;
  const int64_t beg = offsets[i];
  int32_t cnt = (int32_t)(offsets[i + 1] - beg);
  if (cnt > rows) {
    atomicOr(&st->flags, FLAG_ELL_ROWS);
    cnt = rows;
  }
  for (int32_t k = 0; k < cnt; k++) ell[(int64_t)k * n + i] = partners[beg + k];
  const int32_t prev = prev_count[i];
  for (int32_t k = cnt; k < prev; k++) ell[(int64_t)k * n + i] = -1;
  prev_count[i] = cnt;
}

// Last node of every build: the status block goes to pinned host memory (the host reads it after synchronising the
// stream; a zero-copy store instead of a copy-engine node) and every piece of per-build state — histogram, look-back
// words of both scans, queue, ticket, the status block itself — is cleared for the NEXT build, so that a build starts
// without a memset node.
struct HaloCtrl;
__device__ void halo_done_device(HaloCtrl* ctrl, unsigned long long* peer_free_lo, unsigned long long* peer_free_hi);
__global__ void __launch_bounds__(256) finalize_kernel(DeviceStatus* __restrict__ st, DeviceStatus* __restrict__ host,
                                                       uint4* __restrict__ zero_region, size_t zero_vecs,
                                                       size_t status_vec0, size_t status_vecs, HaloCtrl* halo_ctrl,
                                                       unsigned long long* peer_free_lo,
                                                       unsigned long long* peer_free_hi) {
  pdl_enter();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *host = *st;
    // a slab rank (nlb200_set_halo_sync): this build has finished reading its ghosts — the neighbours may overwrite
    // them, the step counter advances
    if (halo_ctrl != nullptr) halo_done_device(halo_ctrl, peer_free_lo, peer_free_hi);
  }
  // the status block is cleared by the thread that copied it (block 0 clears all of it after the copy)
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < zero_vecs; i += (size_t)gridDim.x * blockDim.x) {
    if (i >= status_vec0 && i < status_vec0 + status_vecs) continue;
    zero_region[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (blockIdx.x == 0) {
    __syncthreads();
    for (size_t i = threadIdx.x; i < status_vecs; i += blockDim.x) zero_region[status_vec0 + i] = make_uint4(0u, 0u, 0u, 0u);
  }
}

__global__ void fill_i32_kernel(int32_t* p, int64_t n, int32_t v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = v;
}

// ---------------------------------------------------------------------------------------------------------------
// adjacent utilities: slab selection (deterministic, ascending) and record gather for the halo exchange
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) slab_flag_kernel(const T* __restrict__ q, int64_t n, int stride, int axis,
                                                        double lo, double hi, int32_t* __restrict__ flags) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = (double)q[i * stride + axis];
  flags[i] = (v >= lo && v < hi) ? 1 : 0;
}

__global__ void __launch_bounds__(256) slab_compact_kernel(const int32_t* __restrict__ flags,
                                                           const int64_t* __restrict__ pos, int64_t n,
                                                           int32_t* __restrict__ out, int64_t capacity,
                                                           int64_t* __restrict__ out_count) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i == 0) *out_count = pos[n];
  if (i >= n) return;
  if (flags[i] && pos[i] < capacity) out[pos[i]] = (int32_t)i;
}

// both faces of a slab in one pass over the positions: flags_lo[i] = q[i][axis] < cut_lo, flags_hi[i] = q[i][axis] >= cut_hi
template <typename T>
__global__ void __launch_bounds__(256) slab_flag2_kernel(const T* __restrict__ q, int64_t n, int stride, int axis,
                                                         double cut_lo, double cut_hi, int32_t* __restrict__ flags_lo,
                                                         int32_t* __restrict__ flags_hi) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = (double)q[i * stride + axis];
  flags_lo[i] = (v < cut_lo) ? 1 : 0;
  flags_hi[i] = (v >= cut_hi) ? 1 : 0;
}

template <typename T>
__global__ void __launch_bounds__(256) slab_pack2_kernel(const T* __restrict__ q, const int32_t* __restrict__ gids,
                                                         const int32_t* __restrict__ flags_lo,
                                                         const int64_t* __restrict__ pos_lo,
                                                         const int32_t* __restrict__ flags_hi,
                                                         const int64_t* __restrict__ pos_hi, int64_t n, int stride,
                                                         T* __restrict__ out_q_lo, int32_t* __restrict__ out_gid_lo,
                                                         T* __restrict__ out_q_hi, int32_t* __restrict__ out_gid_hi,
                                                         int64_t capacity, int64_t* __restrict__ out_counts) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i == 0) {
    out_counts[0] = pos_lo[n];
    out_counts[1] = pos_hi[n];
  }
  // slots behind the selected records become ABSENT ghosts (NaN x); the grid covers max(n, capacity) threads
  if (i < capacity) {
    const T nan = (T)__longlong_as_double(0x7ff8000000000000ll);
    // every component: an absent record must fail any later coordinate test (periodic images are built axis by axis)
    for (int c = 0; c < stride; c++) {
      if (out_q_lo != nullptr && i >= pos_lo[n]) out_q_lo[i * stride + c] = nan;
      if (out_q_hi != nullptr && i >= pos_hi[n]) out_q_hi[i * stride + c] = nan;
    }
  }
  if (i >= n) return;
  const bool lo = flags_lo[i] != 0 && out_q_lo != nullptr, hi = flags_hi[i] != 0 && out_q_hi != nullptr;
  if (!lo && !hi) return;
  const int32_t g = gids != nullptr ? gids[i] : (int32_t)i;
  if (lo) {
    const int64_t p = pos_lo[i];
    if (p < capacity) {
      for (int c = 0; c < stride; c++) out_q_lo[p * stride + c] = q[i * stride + c];
      out_gid_lo[p] = g;
    }
  }
  if (hi) {
    const int64_t p = pos_hi[i];
    if (p < capacity) {
      for (int c = 0; c < stride; c++) out_q_hi[p * stride + c] = q[i * stride + c];
      out_gid_hi[p] = g;
    }
  }
}

// Both faces of a slab in ONE kernel (nlb200_pack_faces): records with q[axis] < cut_lo go to the lo buffers, records
// with q[axis] >= cut_hi to the hi buffers, positions from warp-aggregated atomics on two cursors (the order of the
// ghosts is whatever the warps arrive in: the build sorts a cell's particles by global id, so its rows do not depend
// on it).  The last CTA to finish publishes the two counts, pads the unused slots up to `capacity` with NaN records
// (absent ghosts) and zeroes the cursors and the ticket for the next call — no memset, no scan, no second launch.
template <typename T>
__global__ void __launch_bounds__(256) pack_faces_kernel(const T* __restrict__ q, const int32_t* __restrict__ gids,
                                                         int64_t n, int stride, int axis, double cut_lo, double cut_hi,
                                                         T* __restrict__ out_q_lo, int32_t* __restrict__ out_gid_lo,
                                                         T* __restrict__ out_q_hi, int32_t* __restrict__ out_gid_hi,
                                                         int64_t capacity, unsigned long long* __restrict__ state,
                                                         int64_t* __restrict__ out_counts) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  bool f_lo = false, f_hi = false;
  if (i < n) {
    const double v = (double)q[i * stride + axis];
    f_lo = out_q_lo != nullptr && v < cut_lo;
    f_hi = out_q_hi != nullptr && v >= cut_hi;
  }
#pragma unroll
  for (int f = 0; f < 2; f++) {
    const bool flag = f == 0 ? f_lo : f_hi;
    const unsigned m = __ballot_sync(0xffffffffu, flag);
    if (m == 0u) continue;
    const int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(&state[f], (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    const int64_t pos = (int64_t)base + __popc(m & ((1u << lane) - 1u));
    if (flag && pos < capacity) {
      T* oq = f == 0 ? out_q_lo : out_q_hi;
      int32_t* og = f == 0 ? out_gid_lo : out_gid_hi;
      for (int c = 0; c < stride; c++) oq[pos * stride + c] = q[i * stride + c];
      og[pos] = gids != nullptr ? gids[i] : (int32_t)i;
    }
  }
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(&state[2], 1ull) == (unsigned long long)gridDim.x - 1ull;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const long long c_lo = (long long)*reinterpret_cast<volatile unsigned long long*>(&state[0]);
  const long long c_hi = (long long)*reinterpret_cast<volatile unsigned long long*>(&state[1]);
  const T nan = (T)__longlong_as_double(0x7ff8000000000000ll);
  // every component: an absent record must fail any later coordinate test (periodic images are built axis by axis)
  if (out_q_lo != nullptr)
    for (int64_t t = c_lo * stride + threadIdx.x; t < capacity * stride; t += blockDim.x) out_q_lo[t] = nan;
  if (out_q_hi != nullptr)
    for (int64_t t = c_hi * stride + threadIdx.x; t < capacity * stride; t += blockDim.x) out_q_hi[t] = nan;
  __syncthreads();
  if (threadIdx.x == 0) {
    out_counts[0] = c_lo;
    out_counts[1] = c_hi;
    state[0] = state[1] = state[2] = 0ull;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Halo exchange by peer stores (no NCCL call on the step): a slab rank's packing kernel writes the ghost records
// STRAIGHT INTO ITS NEIGHBOURS' assembly buffers over NVLink (peer pointers from CUDA IPC) and then raises a flag in
// the neighbour's control block; the neighbour's build waits for the flags of both faces before it reads its ghosts
// and tells the senders when it is done with them.  All flags carry the step number of a device-side counter, so an
// exchange + build replays as a CUDA graph.  Every wait is bounded: a lost peer raises `error` instead of hanging.
//   sender  (pack_faces_p2p_kernel, step n = ctrl->step + 1):
//            wait  free_from[f] >= n - 1     the neighbour has finished reading what step n-1 wrote into it
//            write ghosts + NaN padding into the neighbour's region, __threadfence_system
//            set   neighbour.ready[f'] = n   (last CTA)
//   receiver (halo_wait_kernel before the build):   wait ready[f] >= n for both faces
//            (finalize_kernel after the build):     step = n; set neighbour.free_from[f'] = n
// ---------------------------------------------------------------------------------------------------------------
// The packing step of one CTA of 256 threads (thread <-> record i; i >= n: no record).  Called by every thread of
// every CTA of the launch exactly once; the last CTA to arrive pads, signals the neighbours and — WAIT_OWN — waits for
// this rank's own ghosts.
template <typename T, bool WAIT_OWN>
__device__ __forceinline__ void halo_pack_cta(const T* __restrict__ q, const int32_t* __restrict__ gids, int64_t i,
                                              int64_t n, int stride, const HaloPackArgs& hp) {
  const int axis = hp.axis;
  const double cut_lo = hp.cut_lo, cut_hi = hp.cut_hi;
  T* out_q_lo = reinterpret_cast<T*>(hp.out_q_lo);
  T* out_q_hi = reinterpret_cast<T*>(hp.out_q_hi);
  int32_t* out_gid_lo = hp.out_gid_lo;
  int32_t* out_gid_hi = hp.out_gid_hi;
  const int64_t capacity = hp.capacity;
  unsigned long long* state = hp.state;
  long long* out_counts = hp.out_counts;
  HaloCtrl* ctrl = hp.ctrl;
  unsigned long long* peer_ready_lo = hp.peer_ready_lo;
  unsigned long long* peer_ready_hi = hp.peer_ready_hi;
  const int lane = threadIdx.x & 31;
  __shared__ bool is_last, go;
  bool f_lo = false, f_hi = false;
  if (i < n) {
    const double v = (double)q[i * stride + axis];
    f_lo = out_q_lo != nullptr && v < cut_lo;
    f_hi = out_q_hi != nullptr && v >= cut_hi;
  }
  // only a CTA that has something to send touches the neighbours: it first makes sure that they have finished
  // reading what the previous step wrote into them
  const bool sends = __syncthreads_or(f_lo || f_hi) != 0;
  if (threadIdx.x < 32) {
    // lanes 0..2 read {step, free_from[0], free_from[1]} together: three system-scope loads in flight at once instead
    // of one after the other (each is a round trip to the L2)
    bool ok = true;
    if (sends) {
      unsigned long long v = 0ull;
      if (lane < 3) v = lane == 0 ? ld_own(&ctrl->step) : ld_sys(&ctrl->free_from[lane - 1]);
      const unsigned long long step = __shfl_sync(0xffffffffu, v, 0);
      const bool face = (lane == 1 && out_q_lo != nullptr) || (lane == 2 && out_q_hi != nullptr);
      if (face && v < step) ok = halo_wait_flag(&ctrl->free_from[lane - 1], step);
      ok = __all_sync(0xffffffffu, ok);
      if (!ok && lane == 0) ctrl->error = 1ull;
    }
    if (lane == 0) go = ok;
  }
  __syncthreads();
  if (!go) f_lo = f_hi = false;
#pragma unroll
  for (int f = 0; f < 2; f++) {
    const bool flag = f == 0 ? f_lo : f_hi;
    const unsigned m = __ballot_sync(0xffffffffu, flag);
    if (m == 0u) continue;
    const int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(&state[f], (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    const int64_t pos = (int64_t)base + __popc(m & ((1u << lane) - 1u));
    if (flag && pos < capacity) {
      T* oq = f == 0 ? out_q_lo : out_q_hi;
      int32_t* og = f == 0 ? out_gid_lo : out_gid_hi;
      if (stride == 4 && sizeof(T) == 8) {
        // a {x, y, z, w} double record: two 16-byte peer stores
        const double2* src = reinterpret_cast<const double2*>(q + i * 4);
        double2* dst = reinterpret_cast<double2*>(oq + pos * 4);
        dst[0] = src[0];
        dst[1] = src[1];
      } else {
        for (int c = 0; c < stride; c++) oq[pos * stride + c] = q[i * stride + c];
      }
      og[pos] = gids != nullptr ? gids[i] : (int32_t)i;
      int32_t* sidx = f == 0 ? hp.send_idx_lo : hp.send_idx_hi;
      if (sidx != nullptr) sidx[pos] = (int32_t)i;
    }
  }
  // ONE fence per CTA, after the barrier: it is cumulative over the stores of the CTA's threads (the barrier orders
  // them before it), system-wide if the CTA wrote into a neighbour — a fence by every thread was most of the kernel
  __syncthreads();
  if (threadIdx.x == 0) {
    if (sends)
      __threadfence_system();  // this CTA's peer stores are visible system-wide before its ticket
    else
      __threadfence();
    is_last = atomicAdd(&state[2], 1ull) == (unsigned long long)gridDim.x - 1ull;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const long long c_lo = (long long)*reinterpret_cast<volatile unsigned long long*>(&state[0]);
  const long long c_hi = (long long)*reinterpret_cast<volatile unsigned long long*>(&state[1]);
  const T nan = (T)__longlong_as_double(0x7ff8000000000000ll);
  // absent slots: the regions start out all-NaN (the owner fills them once), so only the slots that held a record of
  // the PREVIOUS step and hold none now have to be cleared (state[3], state[4] remember the previous counts)
  const long long p_lo = (long long)state[3], p_hi = (long long)state[4];
  if (threadIdx.x == 0) go = ld_own(&ctrl->error) == 0ull;  // some CTA gave up waiting: nothing more is sent
  __syncthreads();
  if (go && (c_lo < p_lo || c_hi < p_hi) && threadIdx.x == 0) {
    // slots are about to be cleared in the neighbours: they must have finished the previous step (a CTA that sent
    // records has checked this already; this one may not have sent any)
    const unsigned long long step = ld_own(&ctrl->step);
    bool ok = true;
    if (out_q_lo != nullptr) ok = halo_wait_flag(&ctrl->free_from[0], step) && ok;
    if (out_q_hi != nullptr) ok = halo_wait_flag(&ctrl->free_from[1], step) && ok;
    if (!ok) ctrl->error = 1ull;
    go = ok;
  }
  __syncthreads();
  if (go && out_q_lo != nullptr)
    for (int64_t t = min(c_lo, (long long)capacity) * stride + threadIdx.x; t < min(p_lo, (long long)capacity) * stride;
         t += blockDim.x)
      out_q_lo[t] = nan;
  if (go && out_q_hi != nullptr)
    for (int64_t t = min(c_hi, (long long)capacity) * stride + threadIdx.x; t < min(p_hi, (long long)capacity) * stride;
         t += blockDim.x)
      out_q_hi[t] = nan;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();  // cumulative over the CTA's clearing stores (ordered before it by the barrier)
    out_counts[0] = c_lo;
    out_counts[1] = c_hi;
    state[0] = state[1] = state[2] = 0ull;
    if (go) {
      state[3] = (unsigned long long)c_lo;
      state[4] = (unsigned long long)c_hi;
    }
    const unsigned long long nstep = ld_own(&ctrl->step) + 1ull;
    // (the system-scope fence above ordered every record and padding store of the grid before these two flags)
    if (go && peer_ready_lo != nullptr) st_sys_relaxed(peer_ready_lo, nstep);
    if (go && peer_ready_hi != nullptr) st_sys_relaxed(peer_ready_hi, nstep);
    if (WAIT_OWN) {
      // ... and this rank's own ghosts: wait here for the neighbours' flags, so that the build can follow directly
      bool ok = go;
      if (ok && out_q_lo != nullptr) ok = halo_wait_flag(&ctrl->ready[0], nstep);
      if (ok && out_q_hi != nullptr) ok = halo_wait_flag(&ctrl->ready[1], nstep) && ok;
      if (!ok) ctrl->error = 1ull;
      __threadfence_system();
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) pack_faces_p2p_kernel(const T* __restrict__ q, const int32_t* __restrict__ gids,
                                                             int64_t n, int stride, HaloPackArgs hp) {
  halo_pack_cta<T, true>(q, gids, blockIdx.x * (int64_t)blockDim.x + threadIdx.x, n, stride, hp);
}

// Incremental halo refresh: between two builds a Verlet list stays valid while no particle has
// moved more than margin / 2, but a consumer of the list needs the CURRENT positions of the ghosts.  The face set of
// the last build is recorded (send_idx: slot k of a face <- local record send_idx[k]); this kernel re-sends exactly
// those records into the same slots of the neighbours' ghost regions — no selection, no compaction, ids untouched —
// under the same flag protocol as a packing step: wait until the neighbour has finished with the previous contents,
// peer stores, one system fence per CTA, the last CTA raises the neighbours' `ready` flags and waits for this rank's
// own.  The caller consumes the ghosts and ends the step with nlb200_halo_done.
template <typename T>
__global__ void __launch_bounds__(256) halo_refresh_kernel(const T* __restrict__ q, int stride, HaloPackArgs hp) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  __shared__ bool is_last, go;
  HaloCtrl* ctrl = hp.ctrl;
  const long long n_lo = hp.out_q_lo != nullptr ? min(hp.out_counts[0], hp.capacity) : 0;
  const long long n_hi = hp.out_q_hi != nullptr ? min(hp.out_counts[1], hp.capacity) : 0;
  const bool f_lo = t < n_lo, f_hi = t < n_hi;
  const bool sends = __syncthreads_or(f_lo || f_hi) != 0;
  if (threadIdx.x < 32) {
    bool ok = true;
    if (sends) {
      unsigned long long v = 0ull;
      if (lane < 3) v = lane == 0 ? ld_own(&ctrl->step) : ld_sys(&ctrl->free_from[lane - 1]);
      const unsigned long long step = __shfl_sync(0xffffffffu, v, 0);
      const bool face = (lane == 1 && hp.out_q_lo != nullptr) || (lane == 2 && hp.out_q_hi != nullptr);
      if (face && v < step) ok = halo_wait_flag(&ctrl->free_from[lane - 1], step);
      ok = __all_sync(0xffffffffu, ok);
      if (!ok && lane == 0) ctrl->error = 1ull;
    }
    if (lane == 0) go = ok;
  }
  __syncthreads();
#pragma unroll
  for (int f = 0; f < 2; f++) {
    if (!(f == 0 ? f_lo : f_hi) || !go) continue;
    const int64_t i = (f == 0 ? hp.send_idx_lo : hp.send_idx_hi)[t];
    T* oq = reinterpret_cast<T*>(f == 0 ? hp.out_q_lo : hp.out_q_hi);
    if (stride == 4 && sizeof(T) == 8) {
      const double2* src = reinterpret_cast<const double2*>(q + i * 4);
      double2* dst = reinterpret_cast<double2*>(oq + t * 4);
      dst[0] = src[0];
      dst[1] = src[1];
    } else {
      for (int c = 0; c < stride; c++) oq[t * stride + c] = q[i * stride + c];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (sends)
      __threadfence_system();
    else
      __threadfence();
    is_last = atomicAdd(&hp.state[2], 1ull) == (unsigned long long)gridDim.x - 1ull;
  }
  __syncthreads();
  if (!is_last || threadIdx.x != 0) return;
  __threadfence();
  hp.state[2] = 0ull;
  const bool good = ld_own(&ctrl->error) == 0ull;
  __threadfence_system();
  const unsigned long long nstep = ld_own(&ctrl->step) + 1ull;
  if (good && hp.peer_ready_lo != nullptr) st_sys_relaxed(hp.peer_ready_lo, nstep);
  if (good && hp.peer_ready_hi != nullptr) st_sys_relaxed(hp.peer_ready_hi, nstep);
  bool ok = good;
  if (ok && hp.out_q_lo != nullptr) ok = halo_wait_flag(&ctrl->ready[0], nstep);
  if (ok && hp.out_q_hi != nullptr) ok = halo_wait_flag(&ctrl->ready[1], nstep) && ok;
  if (!ok) ctrl->error = 1ull;
  __threadfence_system();
}

// before the build of a slab rank: the ghosts of both faces have arrived (faces: bit 0 lower, bit 1 upper)
__global__ void halo_wait_kernel(HaloCtrl* ctrl, int faces) {
  if (threadIdx.x != 0) return;
  const unsigned long long want = ld_sys(&ctrl->step) + 1ull;
  bool ok = true;
  if (faces & 1) ok = halo_wait_flag(&ctrl->ready[0], want) && ok;
  if (faces & 2) ok = halo_wait_flag(&ctrl->ready[1], want) && ok;
  if (!ok) ctrl->error = 1ull;
  __threadfence_system();
}

// after the build: this rank has finished reading its ghosts of the step — the neighbours may overwrite them
__device__ void halo_done_device(HaloCtrl* ctrl, unsigned long long* peer_free_lo, unsigned long long* peer_free_hi) {
  const unsigned long long nstep = ld_sys(&ctrl->step) + 1ull;
  st_sys(&ctrl->step, nstep);
  if (peer_free_lo != nullptr) st_sys(peer_free_lo, nstep);
  if (peer_free_hi != nullptr) st_sys(peer_free_hi, nstep);
}
__global__ void halo_done_kernel(HaloCtrl* ctrl, unsigned long long* peer_free_lo, unsigned long long* peer_free_hi) {
  if (threadIdx.x == 0) halo_done_device(ctrl, peer_free_lo, peer_free_hi);
}

// halo packing: the selected records (flags/pos from slab_flag_kernel + scan) go to out_q[pos], their global ids to
// out_gid[pos]; positions >= capacity are dropped (the count tells).  out_q is pre-filled with NaN = absent.
template <typename T>
__global__ void __launch_bounds__(256) slab_pack_kernel(const T* __restrict__ q, const int32_t* __restrict__ gids,
                                                        int32_t gid_base, const int32_t* __restrict__ flags,
                                                        const int64_t* __restrict__ pos, int64_t n, int stride,
                                                        T* __restrict__ out_q, int32_t* __restrict__ out_gid,
                                                        int64_t capacity, int64_t* __restrict__ out_count) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i == 0) *out_count = pos[n];
  if (i >= n || !flags[i]) return;
  const int64_t p = pos[i];
  if (p >= capacity) return;
  for (int c = 0; c < stride; c++) out_q[p * stride + c] = q[i * stride + c];
  out_gid[p] = gids != nullptr ? gids[i] : gid_base + (int32_t)i;
}

// ---------------------------------------------------------------------------------------------------------------
// callers either side of the build
// ---------------------------------------------------------------------------------------------------------------
// f2, Verlet-list lifetime: the search length includes a margin (make_list.cpp:23: 3.0 + 0.3) so that a list stays
// valid until some particle has moved more than margin/2 since the build.  max |q_now - q_ref|^2 over the particles;
// the block maxima are combined with an atomicMax on the bit pattern (non-negative doubles order like integers).
template <typename T>
__global__ void __launch_bounds__(256) max_disp2_kernel(const T* __restrict__ q, const T* __restrict__ qref,
                                                        int64_t n, int stride,
                                                        unsigned long long* __restrict__ out_bits) {
  double m = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double dx = (double)q[i * stride] - (double)qref[i * stride];
    const double dy = (double)q[i * stride + 1] - (double)qref[i * stride + 1];
    const double dz = (double)q[i * stride + 2] - (double)qref[i * stride + 2];
    const double d2 = dx * dx + dy * dy + dz * dz;
    m = d2 > m ? d2 : m;  // NaN never wins; bin_kernel reports NaN positions at the next build
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const double o = __shfl_xor_sync(0xffffffffu, m, d);
    m = o > m ? o : m;
  }
  __shared__ double wm[8];
  if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; k++) m = wm[k] > m ? wm[k] : m;
    atomicMax(out_bits, (unsigned long long)__double_as_longlong(m));
  }
}

// f4, a consumer of the list: Lennard-Jones forces and energy over the FULL CSR rows (the reference allocates a
// momentum array `p` it never uses, make_list.cpp:135-140).  One warp per row: lanes stride over the partners (coalesced
// reads of the row, gathered partner positions), warp-reduce, lane 0 writes.  Pairs beyond rc contribute nothing —
// the list is built with rc + margin.  Plain Euclidean distance, like the list itself.
template <typename T>
__global__ void __launch_bounds__(128) lj_forces_kernel(const T* __restrict__ q, int stride, int32_t n,
                                                        const int64_t* __restrict__ offsets,
                                                        const int32_t* __restrict__ partners, double rc2, double eps,
                                                        double sigma2, double* __restrict__ f,
                                                        double* __restrict__ energy) {
  const int32_t i = (int32_t)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  if (i >= n) return;
  const int lane = lane_id();
  const double xi = (double)q[(int64_t)i * stride], yi = (double)q[(int64_t)i * stride + 1],
               zi = (double)q[(int64_t)i * stride + 2];
  double fx = 0, fy = 0, fz = 0, e = 0;
  const int64_t beg = offsets[i], end = offsets[i + 1];
  for (int64_t k = beg + lane; k < end; k += 32) {
    const int32_t j = partners[k];
    const double dx = xi - (double)q[(int64_t)j * stride], dy = yi - (double)q[(int64_t)j * stride + 1],
                 dz = zi - (double)q[(int64_t)j * stride + 2];
    const double r2 = dx * dx + dy * dy + dz * dz;
    if (r2 < rc2 && r2 > 0.0) {
      const double s2 = sigma2 / r2, s6 = s2 * s2 * s2;
      const double fr = 24.0 * eps * s6 * (2.0 * s6 - 1.0) / r2;  // -(dU/dr)/r
      fx += fr * dx;
      fy += fr * dy;
      fz += fr * dz;
      e += 2.0 * eps * s6 * (s6 - 1.0);  // half of 4 eps (s12 - s6): every pair appears in two rows
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    fx += __shfl_xor_sync(0xffffffffu, fx, d);
    fy += __shfl_xor_sync(0xffffffffu, fy, d);
    fz += __shfl_xor_sync(0xffffffffu, fz, d);
    e += __shfl_xor_sync(0xffffffffu, e, d);
  }
  if (lane == 0) {
    f[(int64_t)i * 3] = fx;
    f[(int64_t)i * 3 + 1] = fy;
    f[(int64_t)i * 3 + 2] = fz;
    if (energy != nullptr) energy[i] = e;
  }
}

// f1, the physical reorder the reference stubbed out (SortPtclData, neighlist_cpu.hpp:176-180; CopyGather,
// neighlist_gpu.hpp:144-151): out[slot] = src[sorted_ids[slot]] for any per-particle array of `width` elements.
template <typename T>
__global__ void __launch_bounds__(256) gather_sorted_kernel(const T* __restrict__ src,
                                                            const int32_t* __restrict__ sorted_ids,
                                                            const int32_t* __restrict__ n_present, int width,
                                                            T* __restrict__ dst) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t slot = t / width;
  if (slot >= *n_present) return;
  const int c = (int)(t - slot * width);
  dst[t] = src[(int64_t)sorted_ids[slot] * width + c];
}

// f3 helper: q[i][axis] += delta for `count` records (periodic images are copies shifted by one box length; absent
// NaN records stay NaN)
template <typename T>
__global__ void __launch_bounds__(256) shift_axis_kernel(T* __restrict__ q, int64_t count, int stride, int axis,
                                                         T delta) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < count) q[i * stride + axis] += delta;
}

template <typename T>
__global__ void __launch_bounds__(256) gather_records_kernel(const T* __restrict__ src,
                                                             const int32_t* __restrict__ idx, int64_t count,
                                                             int stride, T* __restrict__ dst) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= count * stride) return;
  const int64_t k = t / stride;
  const int c = (int)(t - k * stride);
  dst[t] = src[(int64_t)idx[k] * stride + c];
}

}  // namespace nlb
