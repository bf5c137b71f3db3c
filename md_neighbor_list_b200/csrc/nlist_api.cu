// nlist_api.cu — host side of libnlist_b200.so: the C ABI declared in include/nlist_b200.h.
//
// The reference's host logic lives in NeighListGPU (neighlist_gpu.hpp:43-488): constructor (236-255), Allocate
// (94-112), Initialize (268-287), MakeNeighList (289-466), accessors (468-487).  This file is its B200-native
// counterpart: per-handle state instead of globals/statics, one stream-ordered kernel chain with no host
// synchronisation inside a build, the chain replayed as a CUDA graph, capacities checked on the device and reported.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "../../include/nlist_b200.h"
#include "nlist_kernels.cuh"
#include "nlist_rowmask.cuh"
#include "nlist_runmask.cuh"

#ifndef NLB_RM4_RJ
#define NLB_RM4_RJ 8  // candidates per lane of the row-mask search (packed in pairs)
#endif

using namespace nlb;

struct nlb200_context {
  // configuration
  double sl = 0, L[3] = {0, 0, 0};
  int dtype = NLB200_F64, mode = NLB200_FULL_CSR;
  int stride = 4, sort_rows = 0, ell_rows = 200, exact_only = 0, use_graph = 1, variant = 0, profile = 0;
  GridParams<double> gp64;
  GridParams<float> gp32;
  int32_t mesh[3] = {0, 0, 0};   // cells per axis of the handle's grid (the window, if one is set)
  int32_t gmesh[3] = {0, 0, 0};  // ... of the global grid
  int64_t n_cells = 0;
  bool initialized = false;
  int device = 0;

  // capacities
  int64_t max_n = 0, cap_entries = 0;
  int64_t tiles_cells = 0, tiles_n = 0;

  // device buffers
  unsigned char* zero_region = nullptr;  // [cell_count | scan state (cells) | scan state (counts) | status]
  size_t zero_bytes = 0;
  int32_t* cell_count = nullptr;
  unsigned long long* scan_state_cells = nullptr;
  unsigned long long* scan_state_counts = nullptr;
  DeviceStatus* status_dev = nullptr;
  unsigned long long* queue = nullptr;  // work counter of the persistent pair-mask kernel (inside zero_region)
  unsigned int* ticket = nullptr;       // CTA ticket of bin_kernel's last-CTA scan (inside zero_region)
  size_t status_off = 0;                // byte offset of the status block inside zero_region
  void* halo_ctrl = nullptr;            // nlb200_set_halo_sync: control block + the neighbours' free flags
  void* halo_free_lo = nullptr;
  void* halo_free_hi = nullptr;
  HaloPackArgs halo_pack{};              // nlb200_set_halo_pack: the exchange folded into the binning kernels
  bool halo_pack_on = false;
  int path = 0;                         // PATH_*: which search / emission pair the handle runs (pick_path)
  int64_t mic_alloc = 0;                // cell capacity the search buffers were last sized for
  bool mic_alloc_bound = false;
  bool state_clean = false;             // the zero region is all zero (left so by the last build's finalize_kernel)
  int sm_count = 148;
  int64_t l2_bytes = 0;
  int32_t* cell_start = nullptr;
  int2* cell_rank = nullptr;
  int32_t* perm = nullptr;
  int32_t* sorted_ids = nullptr;
  int32_t* slot_cell = nullptr;
  int32_t* slot_gid = nullptr;  // global id per slot (only written when a local -> global map is given)
  float4* rec = nullptr;
  uint32_t* mask = nullptr;  // [27][mask_wi][mask_ncap] pair-mask words
  int32_t mask_wi = 0;       // words per (row, stencil cell): cells may hold up to 32*mask_wi particles
  int64_t mask_ncap = 0;
  int32_t mask_wr = 0;       // run-mask path: words per (row, run); the buffer is [9][mask_wr][mask_ncap]
  CellRec* cellrec = nullptr;    // [M] per-cell records of the row-mask path
  uint32_t* rmask = nullptr;     // row-mask words (per-cell blocks handed out by a cursor)
  int64_t rmask_cap = 0;         // ... capacity in words
  int64_t rmask_need = 0;        // words the last synchronized build asked for
  int32_t win_cap = 0;           // candidates staged per window round
  float band_v3 = 0.f;           // pre-filter band of the row-mask path (absolute FP32 records)
  bool pdl = false;             // NLB200_OPT_PDL
  int64_t max_in_cell_opt = 0;  // NLB200_OPT_MAX_IN_CELL (0 = estimate from the density)
  int32_t* counts = nullptr;
  int64_t* offsets = nullptr;
  int32_t* offsets32 = nullptr;
  int32_t* partners = nullptr;
  int32_t* ell = nullptr;
  int32_t* ell_prev = nullptr;
  int64_t ell_last_n = -1;  // row stride of the ELL view written by the previous build
  void* q_stage = nullptr;  // device staging for nlb200_build_host
  void* q_ref = nullptr;    // positions remembered by nlb200_track_reference (Verlet-list lifetime)
  int64_t q_ref_n = 0;
  unsigned long long* disp_dev = nullptr;
  unsigned long long* disp_host = nullptr;  // pinned
  cudaStream_t own_stream = nullptr;
  cudaStream_t side_stream = nullptr;  // second branch of a build's graph (finalize_kernel beside the emission)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;

  // host mirrors
  DeviceStatus* status_host = nullptr;  // pinned
  cudaStream_t last_stream = nullptr;
  bool build_pending = false;
  bool have_result = false;
  bool have_build = false;  // some build has been enqueued since initialize (nlb200_mark_enqueued needs one)
  int64_t last_n = 0, last_owned = 0;
  nlb200_stats stats{};

  // graph cache (one entry: the reference's usage is LOOP identical builds, make_list.cu:124-127)
  cudaGraphExec_t graph_exec = nullptr;
  const void* g_q = nullptr;
  const int32_t* g_gids = nullptr;
  int64_t g_n = -1, g_owned = -1;

  // per-stage CUDA events (NLB200_OPT_PROFILE): ev[k] is recorded before stage k, ev[n_stages] after the last one
  static constexpr int MAX_STAGES = 16;
  cudaEvent_t ev[MAX_STAGES + 1] = {};
  int stage_id[MAX_STAGES] = {};
  int n_stages = 0;

  std::string err;
};

enum StageId { ST_ZERO = 0, ST_BIN, ST_SCAN_CELLS, ST_SCATTER, ST_CELLSORT, ST_COUNT, ST_SCAN_COUNTS, ST_FILL,
               ST_SORT_ROWS, ST_ELL, ST_STATUS, ST_PAIRMASK, ST_ROWCOUNT, ST_EMIT, ST_ROWMASK, ST_EMIT3, ST_RUNMASK,
               ST_EMITRUN, ST_NUM };
static const char* const kStageNames[ST_NUM] = {"zero", "bin", "scan_cells", "scatter", "cellsort", "search_count",
                                                "scan_counts", "search_fill", "sort_rows", "ell", "status_copy",
                                                "pairmask", "row_count", "emit", "rowmask", "emit3", "runmask",
                                                "emit_run"};

namespace {

int fail(nlb200_context* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  return code;
}

#define CK(h, call)                                                                                     \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess)                                                                              \
      return fail(h, NLB200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                  __LINE__);                                                                            \
  } while (0)

template <typename T>
bool make_grid(double sl, const double* L, GridParams<T>* g) {
  // neighlist_gpu.hpp:240-254 evaluated in T (the reference's Dtype)
  const T slt = (T)sl;
  T msmax = 0;
  int64_t cells = 1;
  for (int d = 0; d < 3; d++) {
    const T l = (T)L[d];
    const int32_t m = (int32_t)(l / slt);
    if (m < 3) return false;
    g->mesh[d] = m;
    g->gmesh[d] = m;
    g->coff[d] = 0;
    g->ms[d] = l / (T)m;
    g->ims[d] = (T)(1.0 / (double)g->ms[d]);
    g->msf[d] = (float)g->ms[d];
    if (g->ms[d] > msmax) msmax = g->ms[d];
    cells *= m;
  }
  if (cells > 2000000000ll) return false;
  g->n_cells = (int32_t)cells;
  g->sl2 = slt * slt;
  g->sl2f = (float)g->sl2;
  // Half-width E of the pre-filter band, in units of (SL2 - r2)/2.  DESIGN.md derives
  //   E <= u * ms_max^2 * ~216 (FP32 evaluation + coordinate rounding) + 16 u SL2 (FP32 reference evaluation)
  // with u = 2^-24; 384 u ms_max^2 leaves > 1.5x slack.
  g->band = (float)(384.0 * std::ldexp(1.0, -24) * (double)msmax * (double)msmax);
  return true;
}

void free_buffers(nlb200_context* h) {
  auto F = [](auto*& p) {
    if (p) cudaFree(p);
    p = nullptr;
  };
  if (h->graph_exec) {
    cudaGraphExecDestroy(h->graph_exec);
    h->graph_exec = nullptr;
  }
  F(h->zero_region);
  F(h->cell_start);
  F(h->cell_rank);
  F(h->perm);
  F(h->sorted_ids);
  F(h->slot_cell);
  F(h->slot_gid);
  F(h->mask);
  F(h->rmask);
  F(h->cellrec);
  F(h->rec);
  F(h->counts);
  F(h->offsets);
  F(h->offsets32);
  F(h->partners);
  F(h->ell);
  F(h->ell_prev);
  F(h->q_stage);
  F(h->q_ref);
  F(h->disp_dev);
  if (h->disp_host) cudaFreeHost(h->disp_host);
  h->disp_host = nullptr;
  if (h->status_host) cudaFreeHost(h->status_host);
  h->status_host = nullptr;
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  h->own_stream = nullptr;
  if (h->side_stream) cudaStreamDestroy(h->side_stream);
  h->side_stream = nullptr;
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  h->ev_fork = h->ev_join = nullptr;
  for (int k = 0; k <= nlb200_context::MAX_STAGES; k++) {
    if (h->ev[k]) cudaEventDestroy(h->ev[k]);
    h->ev[k] = nullptr;
  }
  h->initialized = false;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

void drop_graph(nlb200_context* h) {
  if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
  h->graph_exec = nullptr;
  h->g_q = nullptr;
  h->g_n = -1;
}

// ---- launches of the build chain --------------------------------------------------------------------------------
// With programmatic dependent launch (NLB200_OPT_PDL, default OFF: measured slower) a kernel's CTAs are
// scheduled while its predecessor drains and block in pdl_enter() until the predecessor's writes are visible, inside
// a captured graph (programmatic edges) as well as on a plain stream.
thread_local bool t_pdl = false;

template <typename... KArgs, typename... Args>
cudaError_t launch_chain(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = t_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// ---- search-kernel dispatch -------------------------------------------------------------------------------------
template <typename T, int STRIDE, bool HALF, bool FILL, bool EXACT>
cudaError_t launch_search_t(const SearchArgs<T>& a, int grid, int block, size_t smem, cudaStream_t s) {
  return launch_chain(search_kernel<T, STRIDE, HALF, FILL, EXACT>, dim3(grid), dim3(block), smem, s, a);
}

constexpr int MAX_SEARCH_SMEM = 200 * 1024;

// opt every search-kernel instantiation into > 48 KB of dynamic shared memory (done once, outside graph capture)
template <typename T, int STRIDE, bool HALF, bool FILL>
cudaError_t set_attr_e() {
  cudaError_t e = cudaFuncSetAttribute(search_kernel<T, STRIDE, HALF, FILL, false>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SEARCH_SMEM);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(search_kernel<T, STRIDE, HALF, FILL, true>,
                              cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SEARCH_SMEM);
}
template <typename T, int STRIDE>
cudaError_t set_attr_ts() {
  cudaError_t e;
  if ((e = set_attr_e<T, STRIDE, false, false>()) != cudaSuccess) return e;
  if ((e = set_attr_e<T, STRIDE, false, true>()) != cudaSuccess) return e;
  if ((e = set_attr_e<T, STRIDE, true, false>()) != cudaSuccess) return e;
  return set_attr_e<T, STRIDE, true, true>();
}
constexpr int MAX_EMIT_SMEM = 200 * 1024;

template <bool HALF, bool GID, bool COUNT>
cudaError_t launch_emit_t(bool direct, const EmitArgs& a, cudaStream_t s) {
#ifdef NLB_ABLATIONS
  if (direct)
    return launch_chain(emit_direct_kernel<HALF, GID, COUNT>, dim3((unsigned)((a.n_total + 127) / 128)), dim3(128), 0,
                        s, a);
#else
  (void)direct;  // the direct-store emission is an ablation: built only with -DNLB_ABLATIONS
#endif
  constexpr int rows = EM_WARPS * 32;
  return launch_chain(emit_kernel<HALF, GID, COUNT>, dim3((unsigned)((a.n_total + rows - 1) / rows)), dim3(rows),
                      (size_t)EM_WARPS * 32 * EM_LINE * sizeof(int32_t), s, a);
}
cudaError_t launch_emit(bool half, bool count, bool direct, const EmitArgs& a, cudaStream_t s) {
  const bool gid = a.global_ids != nullptr;
  if (!half) {
    if (count) return cudaSuccess;  // FULL: the popcount pass already wrote counts[]
    return gid ? launch_emit_t<false, true, false>(direct, a, s) : launch_emit_t<false, false, false>(direct, a, s);
  }
  if (count) return gid ? launch_emit_t<true, true, true>(direct, a, s) : launch_emit_t<true, false, true>(direct, a, s);
  return gid ? launch_emit_t<true, true, false>(direct, a, s) : launch_emit_t<true, false, false>(direct, a, s);
}

template <bool HALF, bool GID, bool COUNT>
cudaError_t set_emit_attr() {
  return cudaFuncSetAttribute(emit_kernel<HALF, GID, COUNT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              MAX_EMIT_SMEM);
}

template <typename T, int STRIDE>
cudaError_t set_pm_attr() {
  cudaError_t e = cudaFuncSetAttribute(pairmask_kernel<T, STRIDE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       MAX_EMIT_SMEM);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(pairmask_kernel<T, STRIDE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              MAX_EMIT_SMEM);
}

cudaError_t set_runmask_attrs() {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(runmask_kernel<double, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_EMIT_SMEM)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(runmask_kernel<double, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_EMIT_SMEM)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(runmask_kernel<float, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_EMIT_SMEM)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(runmask_kernel<float, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_EMIT_SMEM)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(runmask_kernel<double, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_EMIT_SMEM)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(runmask_kernel<double, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_EMIT_SMEM)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(runmask_kernel<float, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_EMIT_SMEM)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(runmask_kernel<float, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_EMIT_SMEM)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(emitwin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_EMIT_SMEM)) != cudaSuccess) return e;
  return cudaFuncSetAttribute(emitrun_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_EMIT_SMEM);
}

cudaError_t set_search_attrs() {
  cudaError_t e;
  if ((e = set_pm_attr<double, 4>()) != cudaSuccess) return e;
  if ((e = set_pm_attr<double, 3>()) != cudaSuccess) return e;
  if ((e = set_pm_attr<float, 4>()) != cudaSuccess) return e;
  if ((e = set_pm_attr<float, 3>()) != cudaSuccess) return e;
  if ((e = set_emit_attr<false, false, false>()) != cudaSuccess) return e;
  if ((e = set_emit_attr<false, true, false>()) != cudaSuccess) return e;
  if ((e = set_emit_attr<true, false, false>()) != cudaSuccess) return e;
  if ((e = set_emit_attr<true, true, false>()) != cudaSuccess) return e;
  if ((e = set_emit_attr<true, false, true>()) != cudaSuccess) return e;
  if ((e = set_emit_attr<true, true, true>()) != cudaSuccess) return e;
  if ((e = set_attr_ts<double, 4>()) != cudaSuccess) return e;
  if ((e = set_attr_ts<double, 3>()) != cudaSuccess) return e;
  if ((e = set_attr_ts<float, 4>()) != cudaSuccess) return e;
  return set_attr_ts<float, 3>();
}

template <typename T, int STRIDE, bool HALF, bool FILL>
cudaError_t launch_search_e(bool exact, const SearchArgs<T>& a, int grid, int block, size_t smem, cudaStream_t s) {
  return exact ? launch_search_t<T, STRIDE, HALF, FILL, true>(a, grid, block, smem, s)
               : launch_search_t<T, STRIDE, HALF, FILL, false>(a, grid, block, smem, s);
}

template <typename T, int STRIDE>
cudaError_t launch_search(bool half, bool fill, bool exact, const SearchArgs<T>& a, int grid, int block, size_t smem,
                          cudaStream_t s) {
  if (half)
    return fill ? launch_search_e<T, STRIDE, true, true>(exact, a, grid, block, smem, s)
                : launch_search_e<T, STRIDE, true, false>(exact, a, grid, block, smem, s);
  return fill ? launch_search_e<T, STRIDE, false, true>(exact, a, grid, block, smem, s)
              : launch_search_e<T, STRIDE, false, false>(exact, a, grid, block, smem, s);
}


// Which search / emission pair a handle runs.  NLB200_OPT_KERNEL_VARIANT:
//   0 (default): RUN MASKS (runmask_kernel + emitrun_kernel / emitwin_kernel, nlist_runmask.cuh) — FULL and HALF lists.
//                Emission: emitrun_kernel (ids gathered from global memory) below EMITWIN_MIN_PARTICLES particles,
//                emitwin_kernel (ids through a shared-memory window) from there on; 9 / 10 force the one / the other.
//                (a cell is ordered by the id its rows report — local, or global with a map — so the HALF cut is a suffix)
//                Any mode with a cell of more than PAIRMASK_MAX_CELL particles (clustered inputs): ROW MASKS
//                (rowmask4_kernel, emit3_kernel: one bit per test, a block per cell from a cursor);
//                nlb200_reserve_cell_capacity switches.
//   1, or NLB200_OPT_EXACT_ONLY: search_kernel twice (count, fill): every test in the input precision if asked
//   2, 3, 4, 7, 100..:           pair masks and their ablations
//   5: row masks with the CTA-per-cell search (rowmask_kernel)      6, 20..39: row masks (rowmask4_kernel)
//   8: run masks (= default)     200..299: run masks with (variant - 200) units per cell (tuning)
enum { PATH_V1 = 1, PATH_PAIRMASK = 2, PATH_ROWMASK = 3, PATH_RUNMASK = 4 };
constexpr int64_t PAIRMASK_MAX_CELL = 256;
constexpr int32_t EMITWIN_MIN_PARTICLES = 1 << 20;
int pick_path(const nlb200_context* h, int64_t max_in_cell) {
  if (h->exact_only != 0 || h->variant == 1) return PATH_V1;
  if (h->variant == 5 || h->variant == 6 || (h->variant >= 20 && h->variant < 40)) return PATH_ROWMASK;
  const bool rn_tuning = h->variant == 9 || h->variant == 10 || (h->variant >= 200 && h->variant < 300);
  if (h->variant != 0 && h->variant != 8 && !rn_tuning) return PATH_PAIRMASK;
  if (max_in_cell > PAIRMASK_MAX_CELL) return PATH_ROWMASK;
  // HALF lists take the run masks too: the id cut is a per-cell suffix (cellsort_kernel orders a cell by the id its
  // rows report — the local id, or the global id of a local -> global map)
  return PATH_RUNMASK;  // 36 bytes x words-per-run per particle, dense
}
bool uses_v1(const nlb200_context* h) { return h->path == PATH_V1; }
bool uses_rowmask(const nlb200_context* h) { return h->path == PATH_ROWMASK; }
bool uses_runmask(const nlb200_context* h) { return h->path == PATH_RUNMASK; }

template <typename T, int STRIDE, int HALFMODE>
cudaError_t set_rm_attr_h() {
  return cudaFuncSetAttribute(rowmask_kernel<T, STRIDE, HALFMODE, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              MAX_EMIT_SMEM);
}
template <typename T, int STRIDE>
cudaError_t set_rm_attr() {
  cudaError_t e;
  if ((e = set_rm_attr_h<T, STRIDE, 0>()) != cudaSuccess) return e;
  if ((e = set_rm_attr_h<T, STRIDE, 1>()) != cudaSuccess) return e;
  return set_rm_attr_h<T, STRIDE, 2>();
}
cudaError_t set_rowmask_attrs() {
  cudaError_t e;
  if ((e = set_rm_attr<double, 4>()) != cudaSuccess) return e;
  if ((e = set_rm_attr<double, 3>()) != cudaSuccess) return e;
  if ((e = set_rm_attr<float, 4>()) != cudaSuccess) return e;
  if ((e = set_rm_attr<float, 3>()) != cudaSuccess) return e;
  return cudaFuncSetAttribute(emit3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_EMIT_SMEM);
}

template <typename T, int STRIDE, int HALFMODE>
int launch_rowmask_h(nlb200_context* h, const RowMaskArgs<T>& a, cudaStream_t s) {
  const size_t smem = rm_smem_bytes(a.win_cap);
  int per_sm = 0;
  CK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rowmask_kernel<T, STRIDE, HALFMODE, 8>, RM_THREADS,
                                                      smem));
  if (per_sm < 1) return fail(h, NLB200_ERR_CUDA, "row-mask kernel does not fit an SM (%zu bytes of shared memory)", smem);
  int64_t grid = (int64_t)per_sm * h->sm_count;
  if (grid > a.gp.n_cells) grid = a.gp.n_cells;
  CK(h, launch_chain(rowmask_kernel<T, STRIDE, HALFMODE, 8>, dim3((unsigned)grid), dim3(RM_THREADS), smem, s, a));
  return NLB200_OK;
}
template <typename T, int STRIDE, int HALFMODE>
int launch_rowmask4_h(nlb200_context* h, RowMask4Args<T>& a, cudaStream_t s) {
  const size_t smem = sizeof(Rm4Smem<NLB_RM4_RJ>) * RM_WARPS;
  int per_sm = 0;
  CK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rowmask4_kernel<T, STRIDE, HALFMODE, NLB_RM4_RJ>, RM_THREADS,
                                                      smem));
  if (per_sm < 1) return fail(h, NLB200_ERR_CUDA, "row-mask kernel does not fit an SM (%zu bytes of shared memory)", smem);
  int64_t grid = (int64_t)per_sm * h->sm_count;
  const int64_t units = (int64_t)a.gp.n_cells * a.upc;
  if (grid * RM_WARPS > units) grid = (units + RM_WARPS - 1) / RM_WARPS;
  CK(h, launch_chain(rowmask4_kernel<T, STRIDE, HALFMODE, NLB_RM4_RJ>, dim3((unsigned)grid), dim3(RM_THREADS), smem, s, a));
  return NLB200_OK;
}
template <typename T, int STRIDE>
int launch_rowmask4(nlb200_context* h, int halfmode, RowMask4Args<T>& a, cudaStream_t s) {
  if (halfmode == 0) return launch_rowmask4_h<T, STRIDE, 0>(h, a, s);
  if (halfmode == 1) return launch_rowmask4_h<T, STRIDE, 1>(h, a, s);
  return launch_rowmask4_h<T, STRIDE, 2>(h, a, s);
}
template <typename T, int STRIDE, int HALFMODE>
cudaError_t set_rm4_attr_h() {
  return cudaFuncSetAttribute(rowmask4_kernel<T, STRIDE, HALFMODE, NLB_RM4_RJ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              MAX_EMIT_SMEM);
}
template <typename T, int STRIDE>
cudaError_t set_rm4_attr() {
  cudaError_t e;
  if ((e = set_rm4_attr_h<T, STRIDE, 0>()) != cudaSuccess) return e;
  if ((e = set_rm4_attr_h<T, STRIDE, 1>()) != cudaSuccess) return e;
  return set_rm4_attr_h<T, STRIDE, 2>();
}
cudaError_t set_rowmask4_attrs() {
  cudaError_t e;
  if ((e = set_rm4_attr<double, 4>()) != cudaSuccess) return e;
  if ((e = set_rm4_attr<double, 3>()) != cudaSuccess) return e;
  if ((e = set_rm4_attr<float, 4>()) != cudaSuccess) return e;
  return set_rm4_attr<float, 3>();
}

template <typename T, int STRIDE>
int launch_rowmask(nlb200_context* h, int halfmode, const RowMaskArgs<T>& a, cudaStream_t s) {
  if (halfmode == 0) return launch_rowmask_h<T, STRIDE, 0>(h, a, s);
  if (halfmode == 1) return launch_rowmask_h<T, STRIDE, 1>(h, a, s);
  return launch_rowmask_h<T, STRIDE, 2>(h, a, s);
}

template <typename T>
const GridParams<T>& grid_of(const nlb200_context* h);
template <>
const GridParams<double>& grid_of<double>(const nlb200_context* h) {
  return h->gp64;
}
template <>
const GridParams<float>& grid_of<float>(const nlb200_context* h) {
  return h->gp32;
}

int64_t estimate_entries(const nlb200_context* h, int64_t n);

// Enqueue one build on `s` (plain launches; the caller may be capturing them into a graph).
template <typename T, int STRIDE>
int enqueue_build(nlb200_context* h, const T* q, int64_t n_total, int64_t n_owned, const int32_t* gids,
                  cudaStream_t s) {
  const GridParams<T>& gp = grid_of<T>(h);
  const int32_t n = (int32_t)n_total;
  const int32_t M = gp.n_cells;
  h->n_stages = 0;
  t_pdl = h->pdl && !h->profile;  // stage events between the kernels would serialise them anyway
  auto stage = [&](int id) -> cudaError_t {
    if (!h->profile) return cudaSuccess;
    if (h->n_stages >= nlb200_context::MAX_STAGES) return cudaSuccess;
    h->stage_id[h->n_stages] = id;
    return cudaEventRecord(h->ev[h->n_stages++], s);
  };
  CK(h, stage(ST_ZERO));
  // every build leaves the per-build state zeroed (finalize_kernel); only a build that follows a failed enqueue
  // clears it itself
  if (!h->state_clean) CK(h, cudaMemsetAsync(h->zero_region, 0, h->zero_bytes, s));
  h->state_clean = false;
  CK(h, stage(ST_BIN));
  const bool scan_in_bin = n > 0 && M <= BIN_SCAN_MAX_CELLS;
  // a slab rank with nlb200_set_halo_pack: the owned records are binned AND sent by the first launch, the ghosts are
  // binned by a second one that waits for the neighbours' flags (bin_kernel, HALO)
  const bool fused_halo = h->halo_pack_on && gids != nullptr && n_owned > 0 && n_owned < n_total;
  if (n > 0 && fused_halo) {
    const int32_t no = (int32_t)n_owned;
    bin_kernel<T, STRIDE, false, 1><<<(no + 255) / 256, 256, 0, s>>>(q, 0, no, no, gp, h->cell_count, h->cell_rank,
                                                                     h->status_dev, h->cell_start, h->ticket, gids,
                                                                     h->halo_pack);
    CK(h, cudaGetLastError());
    if (scan_in_bin)
      bin_kernel<T, STRIDE, true, 2><<<(n - no + 255) / 256, 256, 0, s>>>(q, no, n, no, gp, h->cell_count, h->cell_rank,
                                                                          h->status_dev, h->cell_start, h->ticket, gids,
                                                                          h->halo_pack);
    else
      bin_kernel<T, STRIDE, false, 2><<<(n - no + 255) / 256, 256, 0, s>>>(q, no, n, no, gp, h->cell_count,
                                                                           h->cell_rank, h->status_dev, h->cell_start,
                                                                           h->ticket, gids, h->halo_pack);
    CK(h, cudaGetLastError());
  } else if (n > 0) {
    // first kernel of the chain: a plain launch
    if (scan_in_bin)
      bin_kernel<T, STRIDE, true><<<(n + 255) / 256, 256, 0, s>>>(q, 0, n, (int32_t)n_owned, gp, h->cell_count,
                                                                  h->cell_rank, h->status_dev, h->cell_start, h->ticket,
                                                                  gids, HaloPackArgs{});
    else
      bin_kernel<T, STRIDE, false><<<(n + 255) / 256, 256, 0, s>>>(q, 0, n, (int32_t)n_owned, gp, h->cell_count,
                                                                   h->cell_rank, h->status_dev, h->cell_start, h->ticket,
                                                                   gids, HaloPackArgs{});
    CK(h, cudaGetLastError());
  }
  CK(h, stage(ST_SCAN_CELLS));
  if (!scan_in_bin) {
    const int tiles = (int)((M + SCAN_TILE - 1) / SCAN_TILE);
    const bool after_kernel = n > 0;  // n == 0: the scan follows the memset directly
    const bool keep = t_pdl;
    t_pdl = keep && after_kernel;
    CK(h, launch_chain(scan_kernel<int32_t>, dim3(tiles), dim3(SCAN_THREADS), 0, s, (const int32_t*)h->cell_count,
                       (int64_t)M, h->cell_start, (int32_t*)nullptr, h->scan_state_cells, (DeviceStatus*)nullptr,
                       &h->status_dev->max_in_cell, 0ll));
    t_pdl = keep;
  }
  CK(h, stage(ST_SCATTER));
  if (n > 0) {
    CK(h, launch_chain(scatter_kernel, dim3((n + 255) / 256), dim3(256), 0, s, (const int2*)h->cell_rank, n,
                       (const int32_t*)h->cell_start, h->perm));
    CK(h, stage(ST_CELLSORT));
    int64_t cs_blocks = ((int64_t)M * 32 + 127) / 128;
    if (cs_blocks > (int64_t)h->sm_count * 64) cs_blocks = (int64_t)h->sm_count * 64;  // warps stride over the cells
    // cells per warp batch: 1 while every warp gets about one cell, up to 32 on large grids (see cellsort_kernel)
    int32_t cs_batch = (int32_t)((int64_t)M / (cs_blocks * 4 * 2));
    if (cs_batch < 1) cs_batch = 1;
    if (cs_batch > 32) cs_batch = 32;
    if (uses_rowmask(h))
      // variant 5 (the CTA-per-cell search) writes the cell records itself
      CK(h, launch_chain(cellsort_kernel<T, STRIDE, true>, dim3((unsigned)cs_blocks), dim3(128), 0, s, q, gp,
                         (const int32_t*)h->cell_start, (const int32_t*)h->perm, h->sorted_ids, h->rec, h->slot_cell,
                         gids, h->slot_gid, h->variant == 5 ? (CellRec*)nullptr : h->cellrec,
                         (unsigned long long)h->rmask_cap, h->counts, (int32_t)n_owned, h->status_dev, cs_batch,
                         (int32_t)CS_BIG));
    else
      CK(h, launch_chain(cellsort_kernel<T, STRIDE, false>, dim3((unsigned)cs_blocks), dim3(128), 0, s, q, gp,
                         (const int32_t*)h->cell_start, (const int32_t*)h->perm, h->sorted_ids, h->rec, h->slot_cell,
                         gids, h->slot_gid, (CellRec*)nullptr, 0ull, uses_runmask(h) ? h->counts : (int32_t*)nullptr,
                         (int32_t)n_owned, h->status_dev, 1, 0));
    if (uses_rowmask(h))
      // the crowded cells' particles: one warp per 32 slots instead of one warp per cell (see cellsort_big_kernel)
      CK(h, launch_chain(cellsort_big_kernel<T, STRIDE, true>, dim3((unsigned)((n + 127) / 128)), dim3(128), 0, s, q, gp,
                         (const int32_t*)h->cell_start, (const int2*)h->cell_rank, (const int32_t*)h->perm,
                         h->sorted_ids, h->rec, h->slot_cell, gids, h->slot_gid, h->counts, (int32_t)n_owned,
                         (int32_t)CS_BIG));
  }
  const bool half = h->mode == NLB200_HALF_CSR;
  const bool use_v1 = uses_v1(h);
  // finalize_kernel: status block to mapped host memory, per-build state re-zeroed, (slab ranks) the neighbours told
  // that their ghosts may be overwritten
  bool finalized_early = false;
  auto launch_finalize = [&](cudaStream_t fs, bool chained) -> int {
    const size_t vecs = h->zero_bytes / 16;
    unsigned fgrid = (unsigned)((vecs + 1023) / 1024);
    if (fgrid < 1) fgrid = 1;
    if (fgrid > (unsigned)h->sm_count * 4) fgrid = (unsigned)h->sm_count * 4;
    if (chained) {
      CK(h, launch_chain(finalize_kernel, dim3(fgrid), dim3(256), 0, fs, h->status_dev, h->status_host,
                         reinterpret_cast<uint4*>(h->zero_region), vecs, h->status_off / 16,
                         (sizeof(DeviceStatus) + 15) / 16, reinterpret_cast<HaloCtrl*>(h->halo_ctrl),
                         reinterpret_cast<unsigned long long*>(h->halo_free_lo),
                         reinterpret_cast<unsigned long long*>(h->halo_free_hi)));
    } else {
      finalize_kernel<<<fgrid, 256, 0, fs>>>(h->status_dev, h->status_host, reinterpret_cast<uint4*>(h->zero_region),
                                             vecs, h->status_off / 16, (sizeof(DeviceStatus) + 15) / 16,
                                             reinterpret_cast<HaloCtrl*>(h->halo_ctrl),
                                             reinterpret_cast<unsigned long long*>(h->halo_free_lo),
                                             reinterpret_cast<unsigned long long*>(h->halo_free_hi));
      CK(h, cudaGetLastError());
    }
    return NLB200_OK;
  };
  // ... as a second branch of the graph beside an emission that writes the list only (called after the offsets scan)
  auto fork_finalize = [&]() -> int {
    if (!(n > 0 && !h->profile && !t_pdl && !h->sort_rows && h->mode != NLB200_FULL_ELL_TRANSPOSED && s != nullptr &&
          s != cudaStreamLegacy && s != cudaStreamPerThread))
      return NLB200_OK;
    CK(h, cudaEventRecord(h->ev_fork, s));
    CK(h, cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
    const int rc_fin = launch_finalize(h->side_stream, false);
    if (rc_fin) return rc_fin;
    CK(h, cudaEventRecord(h->ev_join, h->side_stream));
    finalized_early = true;
    return NLB200_OK;
  };
  if (uses_rowmask(h)) {
    // --- default: row masks.  Search (every test once, verdict blocks transposed to row-major words, row lengths
    //     on the way) -> offsets -> emission. ---
    RowMaskArgs<T> rm;
    rm.q = q;
    rm.gp = gp;
    rm.cell_start = h->cell_start;
    rm.rec = h->rec;
    rm.global_ids = gids;
    rm.n_owned = (int32_t)n_owned;
    rm.mask = h->rmask;
    rm.mask_cap = (unsigned long long)h->rmask_cap;
    rm.cellrec = h->cellrec;
    rm.counts = h->counts;
    rm.d_mx = make_fastdiv((uint32_t)gp.mesh[0]);
    rm.d_my = make_fastdiv((uint32_t)gp.mesh[1]);
    rm.band = h->band_v3;
    rm.win_cap = h->win_cap;
    rm.queue = h->queue;
    rm.st = h->status_dev;
    CK(h, stage(ST_ROWMASK));
    if (n > 0 && h->variant == 5) {
      const int rc = launch_rowmask<T, STRIDE>(h, !half ? 0 : (gids == nullptr ? 1 : 2), rm, s);
      if (rc) return rc;
    } else if (n > 0) {
      RowMask4Args<T> r4;
      r4.q = q;
      r4.gp = gp;
      r4.cell_start = h->cell_start;
      r4.rec = h->rec;
      r4.global_ids = gids;
      r4.n_owned = (int32_t)n_owned;
      r4.mask = h->rmask;
      r4.mask_cap = (unsigned long long)h->rmask_cap;
      r4.cellrec = h->cellrec;
      r4.counts = h->counts;
      // parts per cell: one warp-sized chunk of 256 candidates each for a mean window; crowded cells loop inside
      int64_t upc = (int64_t)(27.0 * ((double)n / (double)M) / (32.0 * NLB_RM4_RJ) + 0.999);
      if (h->variant >= 20 && h->variant < 40) upc = h->variant - 20;  // tuning override
      if (upc < 1) upc = 1;
      if (upc > 8) upc = 8;
      if ((int64_t)M * upc >= (1ll << 31)) return fail(h, NLB200_ERR_INVALID, "cells x parts exceeds 2^31 units");
      r4.upc = (int32_t)upc;
      r4.d_mx = make_fastdiv((uint32_t)gp.mesh[0]);
      r4.d_my = make_fastdiv((uint32_t)gp.mesh[1]);
      r4.d_upc = make_fastdiv((uint32_t)upc);
      r4.band = h->band_v3;
      r4.queue = reinterpret_cast<unsigned int*>(h->queue);
      r4.st = h->status_dev;
      const int rc = launch_rowmask4<T, STRIDE>(h, !half ? 0 : (gids == nullptr ? 1 : 2), r4, s);
      if (rc) return rc;
    }
    CK(h, stage(ST_SCAN_COUNTS));
    {
      const int tiles = (int)((n_owned + SCAN_TILE - 1) / SCAN_TILE);
      CK(h, launch_chain(scan_kernel<int64_t>, dim3(tiles > 0 ? tiles : 1), dim3(SCAN_THREADS), 0, s,
                         (const int32_t*)h->counts, (int64_t)n_owned, h->offsets, h->offsets32, h->scan_state_counts,
                         h->status_dev, &h->status_dev->max_partners, (long long)h->cap_entries));
    }
    CK(h, stage(ST_EMIT3));
    if (n > 0) {
      Emit3Args em;
      em.cell_start = h->cell_start;
      em.sorted_ids = h->sorted_ids;
      em.slot_cell = h->slot_cell;
      em.slot_pid = gids != nullptr ? h->slot_gid : h->sorted_ids;
      em.cellrec = h->cellrec;
      em.mask = h->rmask;
      em.n_total = n;
      em.n_owned = (int32_t)n_owned;
      em.n_cells = M;
      em.offsets = h->offsets;
      em.partners = h->partners;
      em.capacity = h->cap_entries;
      em.st = h->status_dev;
      constexpr int rows = EM3_WARPS * 32;
      CK(h, launch_chain(emit3_kernel, dim3((unsigned)((n + rows - 1) / rows)), dim3(rows),
                         (size_t)rows * EM_LINE * sizeof(int32_t), s, em));
    }
  } else if (uses_runmask(h)) {
    // --- FULL lists: run masks.  Search (every ordered pair once; bits = the particles of an x-run, row lengths
    //     by RED) -> offsets -> emission. ---
    CK(h, stage(ST_RUNMASK));
    if (n > 0) {
      RunMaskArgs<T> rn;
      rn.q = q;
      rn.gp = gp;
      rn.cell_start = h->cell_start;
      rn.rec = h->rec;
      rn.sorted_ids = h->sorted_ids;
      rn.cut_ids = gids != nullptr ? h->slot_gid : h->sorted_ids;
      rn.cut_by_slot_table = gids != nullptr ? 1 : 0;
      rn.n_owned = (int32_t)n_owned;
      rn.mask = h->mask;
      rn.n_cap = h->mask_ncap;
      rn.wr = h->mask_wr;
      rn.fits32 = (9ll * h->mask_wr * h->mask_ncap) < (1ll << 32) ? 1 : 0;
      rn.band = gp.band * 1.25f;  // rows reach 1.5 cells from the centre in x here
      rn.queue = h->queue;
      rn.counts = h->counts;
      rn.st = h->status_dev;
      const size_t rn_smem = rn_warp_bytes(h->mask_wr) * (RN_THREADS / 32);
      int per_sm = 0;
      if (half)
        CK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, runmask_kernel<T, STRIDE, true>, RN_THREADS,
                                                            rn_smem));
      else
        CK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, runmask_kernel<T, STRIDE>, RN_THREADS, rn_smem));
      if (per_sm < 1) return fail(h, NLB200_ERR_CUDA, "run-mask kernel does not fit an SM (%zu bytes of shared memory)", rn_smem);
      int64_t grid = (int64_t)per_sm * h->sm_count;
      const int64_t resident = grid * (RN_THREADS / 32);
      // units per cell: one chunk of RN_CH candidates each for a mean column (9 cells); longer columns loop
      int64_t parts = (int64_t)(9.0 * ((double)n / (double)M) * 1.1 / (double)RN_CH + 0.999);
      if (parts < 1) parts = 1;
      if (parts > 16) parts = 16;
      if (h->variant >= 200 && h->variant < 300) parts = h->variant - 200;  // tuning override
      if (parts < 1) parts = 1;
      rn.parts = (int32_t)parts;
      int64_t grab = (M * parts) / (resident * 64);
      if (grab < 1) grab = 1;
      if (grab > 64) grab = 64;
      rn.grab = (int32_t)grab;
      if (M * parts >= (1ll << 31)) return fail(h, NLB200_ERR_INVALID, "cells x parts exceeds 2^31 units");
      rn.d_parts = make_fastdiv((uint32_t)parts);
      rn.d_mx = make_fastdiv((uint32_t)gp.mesh[0]);
      rn.d_my = make_fastdiv((uint32_t)gp.mesh[1]);
      const int64_t need = (M * parts + RN_THREADS / 32 - 1) / (RN_THREADS / 32);
      if (grid > need) grid = need;
      if (half)
        CK(h, launch_chain(runmask_kernel<T, STRIDE, true>, dim3((unsigned)grid), dim3(RN_THREADS), rn_smem, s, rn));
      else
        CK(h, launch_chain(runmask_kernel<T, STRIDE>, dim3((unsigned)grid), dim3(RN_THREADS), rn_smem, s, rn));
    }
    CK(h, stage(ST_SCAN_COUNTS));
    {
      const int tiles = (int)((n_owned + SCAN_TILE - 1) / SCAN_TILE);
      CK(h, launch_chain(scan_kernel<int64_t>, dim3(tiles > 0 ? tiles : 1), dim3(SCAN_THREADS), 0, s,
                         (const int32_t*)h->counts, (int64_t)n_owned, h->offsets, h->offsets32, h->scan_state_counts,
                         h->status_dev, &h->status_dev->max_partners, (long long)h->cap_entries));
    }
    CK(h, stage(ST_EMITRUN));
    // After the offsets scan nothing the status block reports can change any more on this path (the emission writes
    // the list only) and every kernel that reads the ghosts has run: finalize_kernel becomes a second branch of the
    // graph BESIDE the emission instead of a node behind it.
    {
      const int rc_fin = fork_finalize();
      if (rc_fin) return rc_fin;
    }
    if (n > 0) {
      EmitRunArgs em;
      em.cell_start = h->cell_start;
      em.sorted_ids = h->sorted_ids;
      em.slot_cell = h->slot_cell;
      em.slot_pid = gids != nullptr ? h->slot_gid : h->sorted_ids;
      for (int d = 0; d < 3; d++) em.mesh[d] = gp.mesh[d];
      em.d_mx = make_fastdiv((uint32_t)gp.mesh[0]);
      em.d_my = make_fastdiv((uint32_t)gp.mesh[1]);
      em.n_total = n;
      em.n_owned = (int32_t)n_owned;
      em.n_cells = M;
      em.mask = h->mask;
      em.n_cap = h->mask_ncap;
      em.wr = h->mask_wr;
      em.offsets = h->offsets;
      em.partners = h->partners;
      em.capacity = h->cap_entries;
      constexpr int rows = ER_WARPS * 32;
      // partner ids through a shared-memory window once the slot -> id table is far larger than an SM's L1 (the
      // gather of emitrun_kernel then misses: 2 M uniform particles 1.10 -> 1.00 ms of emission); small systems keep
      // the gather (default system: 61.7 us vs 69 us — the window adds an LDS to the expansion loop's chain)
      const bool win = h->variant == 10 || (h->variant != 9 && n >= EMITWIN_MIN_PARTICLES);
      if (!win)
        CK(h, launch_chain(emitrun_kernel, dim3((unsigned)((n + rows - 1) / rows)), dim3(rows),
                           (size_t)rows * EM_LINE * sizeof(int32_t), s, em));
      else
        CK(h, launch_chain(emitwin_kernel, dim3((unsigned)((n + rows - 1) / rows)), dim3(rows), ew_smem_bytes(), s, em));
    }
    if (finalized_early) CK(h, cudaStreamWaitEvent(s, h->ev_join, 0));
  } else if (use_v1) {
    // --- v1: one CTA per cell, thread per particle, test evaluated twice (count, fill).  Kept for the exact-only
    //     validation mode and as an ablation. ---
    const double avg = (double)n_total / (double)M;
    int block = (int)(std::ceil(avg * 1.25 / 32.0) * 32.0);
    if (block < 32) block = 32;
    if (block > 128) block = 128;
    int64_t jt = (int64_t)(27.0 * avg * 1.5);
    jt = (jt + 255) / 256 * 256;
    if (jt < 512) jt = 512;
    if (jt > 8192) jt = 8192;
    const size_t smem = (size_t)jt * (sizeof(float4) + sizeof(int32_t));
    SearchArgs<T> a;
    a.q = q;
    a.gp = gp;
    a.cell_start = h->cell_start;
    a.rec = h->rec;
    a.global_ids = gids;
    a.n_owned = (int32_t)n_owned;
    a.counts = h->counts;
    a.offsets = h->offsets;
    a.partners = h->partners;
    a.capacity = h->cap_entries;
    a.st = h->status_dev;
    a.jt = (int32_t)jt;
    CK(h, stage(ST_COUNT));
    if (n > 0) CK(h, (launch_search<T, STRIDE>(half, false, h->exact_only != 0, a, M, block, smem, s)));
    CK(h, stage(ST_SCAN_COUNTS));
    {
      const int tiles = (int)((n_owned + SCAN_TILE - 1) / SCAN_TILE);
      CK(h, launch_chain(scan_kernel<int64_t>, dim3(tiles > 0 ? tiles : 1), dim3(SCAN_THREADS), 0, s,
                         (const int32_t*)h->counts, (int64_t)n_owned, h->offsets, h->offsets32, h->scan_state_counts,
                         h->status_dev, &h->status_dev->max_partners, (long long)h->cap_entries));
    }
    CK(h, stage(ST_FILL));
    if (n > 0) CK(h, (launch_search<T, STRIDE>(half, true, h->exact_only != 0, a, M, block, smem, s)));
  } else {
    // --- default: pair masks (every test once) -> row counts -> offsets -> staged, coalesced emission ---
    PairMaskArgs<T> pm;
    pm.q = q;
    pm.gp = gp;
    pm.cell_start = h->cell_start;
    pm.rec = h->rec;
    pm.sorted_ids = h->sorted_ids;
    pm.n_owned = (int32_t)n_owned;
    pm.queue = h->queue;
    pm.mask = h->mask;
    pm.n_cap = h->mask_ncap;
    pm.wi = h->mask_wi;
    // variant 7 (test hook): the 64-bit index path that masks of >= 2^32 words (> 53 M particles at 3 words) take
    pm.fits32 = ((27ll * h->mask_wi * h->mask_ncap) < (1ll << 32) && h->variant != 7) ? 1 : 0;
    pm.band = gp.band;
    pm.st = h->status_dev;
    EmitArgs em;
    em.cell_start = h->cell_start;
    em.sorted_ids = h->sorted_ids;
    em.slot_cell = h->slot_cell;
    em.global_ids = gids;
    em.slot_pid = gids != nullptr ? h->slot_gid : h->sorted_ids;
    for (int d = 0; d < 3; d++) em.mesh[d] = gp.mesh[d];
    em.d_mx = make_fastdiv((uint32_t)gp.mesh[0]);
    em.d_my = make_fastdiv((uint32_t)gp.mesh[1]);
    em.n_total = n;
    em.n_owned = (int32_t)n_owned;
    em.n_cells = M;
    em.clear_self = half ? 0 : 1;
    em.mask = h->mask;
    em.n_cap = h->mask_ncap;
    em.wi = h->mask_wi;
    em.counts = h->counts;
    em.offsets = h->offsets;
    em.partners = h->partners;
    em.capacity = h->cap_entries;
    const bool direct = h->variant == 3;  // ablation: per-thread scalar stores instead of the staged emission
    // HALF lists without a global-id map: the id filter is folded into the masks (pairmask_kernel<HALFIDS>), so the
    // popcount pass and the plain emission serve; with a map (multi-GPU) or as variant 4 the emission filters by id
    const bool half_in_mask = half && gids == nullptr && h->variant != 4;
    const bool half_emit = half && !half_in_mask;
    CK(h, stage(ST_PAIRMASK));
    if (n > 0) {
      // persistent warps: as many CTAs as are resident at once, each warp draws cells from the queue
      const size_t pm_smem = pm_warp_bytes(h->mask_wi) * (PM_THREADS / 32);
      int per_sm = 0;
      if (half_in_mask)
        CK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pairmask_kernel<T, STRIDE, true>, PM_THREADS,
                                                            pm_smem));
      else
        CK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pairmask_kernel<T, STRIDE, false>, PM_THREADS,
                                                            pm_smem));
      if (per_sm < 1) per_sm = 1;
      int64_t grid = (int64_t)per_sm * h->sm_count;
      // items per cell: at least ~4 items per resident warp so that the queue can balance a small system
      const int64_t resident = grid * (PM_THREADS / 32);
      int64_t parts = (4 * resident + M - 1) / M;
      if (parts < 1) parts = 1;
      if (parts > 8) parts = 8;
      if (h->variant >= 100 && h->variant < 200) parts = h->variant - 100;  // tuning override
      pm.parts = (int32_t)parts;
      // items per draw from the queue: ~64 draws per resident warp
      int64_t grab = (M * parts) / (resident * 64);
      if (grab < 1) grab = 1;
      if (grab > 64) grab = 64;
      pm.grab = (int32_t)grab;
      if (M * parts >= (1ll << 31)) return fail(h, NLB200_ERR_INVALID, "cells x parts exceeds 2^31 items");
      pm.d_parts = make_fastdiv((uint32_t)parts);
      pm.d_mx = make_fastdiv((uint32_t)gp.mesh[0]);
      pm.d_my = make_fastdiv((uint32_t)gp.mesh[1]);
      const int64_t need = (M * parts + PM_THREADS / 32 - 1) / (PM_THREADS / 32);
      if (grid > need) grid = need;
      if (half_in_mask)
        CK(h, launch_chain(pairmask_kernel<T, STRIDE, true>, dim3((unsigned)grid), dim3(PM_THREADS), pm_smem, s, pm));
      else
        CK(h, launch_chain(pairmask_kernel<T, STRIDE, false>, dim3((unsigned)grid), dim3(PM_THREADS), pm_smem, s, pm));
    }
    CK(h, stage(ST_ROWCOUNT));
    if (n > 0) {
      if (half_emit) {
        CK(h, launch_emit(true, true, direct, em, s));  // HALF rows need the ids to count
      } else {
        CK(h, launch_chain(rowcount_kernel, dim3((unsigned)((n + 127) / 128)), dim3(128), 0, s, em));
      }
    }
    CK(h, stage(ST_SCAN_COUNTS));
    {
      const int tiles = (int)((n_owned + SCAN_TILE - 1) / SCAN_TILE);
      CK(h, launch_chain(scan_kernel<int64_t>, dim3(tiles > 0 ? tiles : 1), dim3(SCAN_THREADS), 0, s,
                         (const int32_t*)h->counts, (int64_t)n_owned, h->offsets, h->offsets32, h->scan_state_counts,
                         h->status_dev, &h->status_dev->max_partners, (long long)h->cap_entries));
    }
    CK(h, stage(ST_EMIT));
    if (!direct && gids == nullptr) {  // (the staged emission writes the list only; with a global-id map it still
                                        //  reads ids the neighbours may overwrite once finalize_kernel has run)
      const int rc_fin = fork_finalize();
      if (rc_fin) return rc_fin;
    }
    if (n > 0) CK(h, launch_emit(half_emit, false, direct, em, s));
    if (finalized_early) CK(h, cudaStreamWaitEvent(s, h->ev_join, 0));
  }
  if (h->sort_rows) CK(h, stage(ST_SORT_ROWS));
  if (h->sort_rows && n_owned > 0) {
    CK(h, launch_chain(sort_rows_kernel, dim3((unsigned)((n_owned + SORT_WARPS - 1) / SORT_WARPS)),
                       dim3(SORT_WARPS * 32), 0, s, (const int64_t*)h->offsets, (int32_t)n_owned, h->partners,
                       (long long)h->cap_entries));
  }
  if (h->mode == NLB200_FULL_ELL_TRANSPOSED) CK(h, stage(ST_ELL));
  if (h->mode == NLB200_FULL_ELL_TRANSPOSED && n_owned > 0) {
    CK(h, launch_chain(ell_kernel, dim3((unsigned)((n_owned + 255) / 256)), dim3(256), 0, s,
                       (const int64_t*)h->offsets, (const int32_t*)h->partners, (int32_t)n_owned, h->ell_rows, h->ell,
                       h->ell_prev, (long long)h->cap_entries, h->status_dev));
  }
  CK(h, stage(ST_STATUS));
  if (!finalized_early) {
    const int rc_fin = launch_finalize(s, true);
    if (rc_fin) return rc_fin;
  }
  h->state_clean = true;
  if (h->profile) CK(h, cudaEventRecord(h->ev[h->n_stages], s));
  return NLB200_OK;
}

int enqueue_dispatch(nlb200_context* h, const void* q, int64_t n_total, int64_t n_owned, const int32_t* gids,
                     cudaStream_t s) {
  if (h->dtype == NLB200_F64) {
    return h->stride == 4 ? enqueue_build<double, 4>(h, (const double*)q, n_total, n_owned, gids, s)
                          : enqueue_build<double, 3>(h, (const double*)q, n_total, n_owned, gids, s);
  }
  return h->stride == 4 ? enqueue_build<float, 4>(h, (const float*)q, n_total, n_owned, gids, s)
                        : enqueue_build<float, 3>(h, (const float*)q, n_total, n_owned, gids, s);
}

int64_t estimate_entries(const nlb200_context* h, int64_t n) {
  // expected entries ~ n * rho * (4/3) pi SL^3, +30 % and a floor for sparse boxes
  const double vol = h->L[0] * h->L[1] * h->L[2];
  const double rho = (double)n / vol;
  double per = rho * 4.18879020478639 * h->sl * h->sl * h->sl;
  if (h->mode == NLB200_HALF_CSR) per *= 0.5;
  double e = (double)n * per * 1.3 + 16.0 * (double)n + 1024.0;
  return (int64_t)e;
}

// New buffer first, old one released only on success: a failed growth (the masks cost 108 * ceil(max_in_cell / 32)
// bytes per particle, so one dense cell of a clustered input can ask for more than the device has) leaves the handle
// exactly as it was — same buffers, same capacities — instead of initialized with a null pointer.
int alloc_mask(nlb200_context* h, int64_t max_in_cell) {
  int64_t wi = (max_in_cell + 31) / 32;
  if (wi < 1) wi = 1;
  const int64_t ncap = (int64_t)align_up((size_t)(h->max_n > 0 ? h->max_n : 1), 32);
  uint32_t* fresh = nullptr;
  CK(h, cudaMalloc(&fresh, sizeof(uint32_t) * (size_t)(27 * wi * ncap)));
  if (h->mask) cudaFree(h->mask);
  h->mask = fresh;
  h->mask_wi = (int32_t)wi;
  h->mask_wr = 0;
  h->mask_ncap = ncap;
  return NLB200_OK;
}

// Before an output buffer is replaced: wait for the build that may still write it, and forget that build — its
// status word must not be reported again by a later nlb200_synchronize over the new, empty buffers.
cudaError_t settle_before_realloc(nlb200_context* h) {
  if (h->build_pending) {
    const cudaError_t e = cudaStreamSynchronize(h->last_stream);
    if (e != cudaSuccess) return e;
  }
  h->build_pending = false;
  h->have_result = false;
  if (h->status_host) h->status_host->flags = 0;
  drop_graph(h);
  return cudaSuccess;
}

// Run masks: [9][wr][ncap] words, wr = words per (row, run).  A run is up to 3 cells: 3 * max_in_cell bounds it.
int alloc_runmask(nlb200_context* h, int64_t max_in_run) {
  int64_t wr = (max_in_run + 31) / 32;
  if (wr < 1) wr = 1;
  const int64_t ncap = (int64_t)align_up((size_t)(h->max_n > 0 ? h->max_n : 1), 32);
  if (h->mask && h->mask_wr == wr && h->mask_ncap == ncap) return NLB200_OK;
  uint32_t* fresh = nullptr;
  CK(h, cudaMalloc(&fresh, sizeof(uint32_t) * (size_t)(9 * wr * ncap)));
  if (h->mask) cudaFree(h->mask);
  h->mask = fresh;
  h->mask_wr = (int32_t)wr;
  h->mask_wi = 0;
  h->mask_ncap = ncap;
  return NLB200_OK;
}

int64_t estimate_max_in_run(const nlb200_context* h, int64_t n) {
  // mean occupancy of three cells + 6 sigma (Poisson) + slack; at least 192 (6 words): a slab rank without a cell
  // window bins on the global grid, so n / cells underestimates its density
  const double avg3 = 3.0 * (double)n / (double)h->n_cells;
  const int64_t est = (int64_t)(avg3 + 6.0 * std::sqrt(avg3) + 8.0);
  return est < 192 ? 192 : est;
}

int alloc_rmask(nlb200_context* h, int64_t words) {
  uint32_t* fresh = nullptr;
  CK(h, cudaMalloc(&fresh, sizeof(uint32_t) * (size_t)(words > 0 ? words : 1)));
  if (h->rmask) cudaFree(h->rmask);
  h->rmask = fresh;
  h->rmask_cap = words;
  return NLB200_OK;
}

// Buffers of the handle's search path for cells of up to `mic` particles.
// mic_is_bound: mic was given by the caller or seen in a build (else it is initialize's density estimate).
int alloc_search_buffers(nlb200_context* h, int64_t mic, bool mic_is_bound = false) {
  h->mic_alloc = mic;
  h->mic_alloc_bound = mic_is_bound;
  const int64_t n = h->max_n > 0 ? h->max_n : 1;
  const int64_t M = h->n_cells;
  if (uses_rowmask(h)) {
    if (!h->cellrec) CK(h, cudaMalloc(&h->cellrec, sizeof(CellRec) * (size_t)M));
    // one bit per test: rows x ceil(27 cells x mic / 32) words bounds a uniform density; clustered inputs report
    // NLB200_ERR_CELL_CAPACITY with the exact need and nlb200_reserve_cell_capacity grows the buffer to it
    int64_t per_row = (27 * std::min<int64_t>(mic, 96) + 31) / 32 + 1;
    const int64_t words = std::max<int64_t>(n * per_row + 4096, h->rmask_need + h->rmask_need / 16 + 4096);
    if (words > h->rmask_cap) {
      const int rc = alloc_rmask(h, words);
      if (rc) return rc;
    }
    // window round of the CTA-per-cell search (variant 5): ~1.2 x the mean window, in chunks of 256 candidates
    const double avg = (double)n / (double)M;
    int64_t wc = (int64_t)(27.0 * (avg > 1.0 ? avg : 1.0) * 1.2);
    wc = (wc + 255) / 256 * 256;
    if (wc < 512) wc = 512;
    if (wc > 6144) wc = 6144;
    h->win_cap = (int32_t)wc;
    // band of the FP32 pre-filter for absolute FP32 records: the evaluation error in the cell frame (384 u ms^2,
    // DESIGN.md §6) plus what the rounding of both particles' coordinates to FP32 can move (SL^2 - r^2)/2 inside
    // r < 2 SL: 2 SL * 2 sqrt(3) * ulp(max |coordinate|)/2 < 8 SL * eps_abs
    double lmax = 0, msmax = 0;
    for (int d = 0; d < 3; d++) {
      lmax = std::max(lmax, h->L[d]);
      msmax = std::max(msmax, h->L[d] / (double)h->gmesh[d]);
    }
    int ex = 0;
    std::frexp(lmax + 2.0 * msmax, &ex);  // value < 2^ex: ulp = 2^(ex - 24)
    const double eps_abs = h->dtype == NLB200_F64 ? std::ldexp(1.0, ex - 25) : 0.0;
    h->band_v3 = (float)(384.0 * std::ldexp(1.0, -24) * msmax * msmax + 8.0 * h->sl * eps_abs);
    if (h->mask) {  // the handle came from the pair-mask path
      cudaFree(h->mask);
      h->mask = nullptr;
      h->mask_wi = 0;
      h->mask_wr = 0;
    }
    return NLB200_OK;
  }
  if (uses_runmask(h)) {
    return alloc_runmask(h, mic_is_bound ? 3 * mic : estimate_max_in_run(h, n));
  }
  if (!uses_v1(h)) return alloc_mask(h, mic);
  return NLB200_OK;
}

int64_t estimate_max_in_cell(const nlb200_context* h, int64_t n) {
  // mean occupancy + 6 sigma of a Poisson cell count + slack; lattices stay well below (13-63 at
  // mean 35.3).  Floor of 80 (3 words): a rank of a slab decomposition bins on the GLOBAL grid, so n / cells
  // underestimates its density by the number of ranks.  Clustered inputs report NLB200_ERR_CELL_CAPACITY and the
  // caller (or nlb200_build_host) grows it.
  const double avg = (double)n / (double)h->n_cells;
  const int64_t est = (int64_t)(avg + 6.0 * std::sqrt(avg) + 8.0);
  return est < 80 ? 80 : est;
}

int alloc_partners(nlb200_context* h, int64_t entries) {
  int32_t* fresh = nullptr;
  CK(h, cudaMalloc(&fresh, sizeof(int32_t) * (size_t)(entries > 0 ? entries : 1)));
  if (h->partners) cudaFree(h->partners);
  h->partners = fresh;
  h->cap_entries = entries;
  return NLB200_OK;
}

}  // namespace

extern "C" {

int nlb200_version(void) { return NLB200_VERSION; }

const char* nlb200_status_string(int status) {
  switch (status) {
    case NLB200_OK: return "ok";
    case NLB200_ERR_INVALID: return "invalid argument";
    case NLB200_ERR_CUDA: return "CUDA error";
    case NLB200_ERR_CAPACITY: return "partner-list capacity exceeded";
    case NLB200_ERR_OUT_OF_BOX: return "particle outside the box";
    case NLB200_ERR_ELL_ROWS: return "row longer than the ELL row capacity";
    case NLB200_ERR_STATE: return "invalid call order";
    case NLB200_ERR_CELL_CAPACITY: return "a cell holds more particles than the pair-mask words cover";
  }
  return "unknown";
}

int nlb200_create(double search_length, double lx, double ly, double lz, int dtype, int mode, nlb200_handle* out) {
  if (!out) return NLB200_ERR_INVALID;
  *out = nullptr;
  if (!(search_length > 0) || !(lx > 0) || !(ly > 0) || !(lz > 0)) return NLB200_ERR_INVALID;
  if (dtype != NLB200_F32 && dtype != NLB200_F64) return NLB200_ERR_INVALID;
  if (mode < NLB200_HALF_CSR || mode > NLB200_FULL_ELL_TRANSPOSED) return NLB200_ERR_INVALID;
  nlb200_context* h = new (std::nothrow) nlb200_context();
  if (!h) return NLB200_ERR_INVALID;
  h->sl = search_length;
  h->L[0] = lx;
  h->L[1] = ly;
  h->L[2] = lz;
  h->dtype = dtype;
  h->mode = mode;
  if (const char* e = std::getenv("NLB200_PDL")) h->pdl = std::atoi(e) != 0;
  const bool ok = (dtype == NLB200_F64) ? make_grid<double>(search_length, h->L, &h->gp64)
                                        : make_grid<float>(search_length, h->L, &h->gp32);
  if (!ok) {
    // reference: mesh_size < 3 makes the wrapped stencil visit a cell twice (duplicate pairs), SURVEY.md §2b
    delete h;
    return NLB200_ERR_INVALID;
  }
  const int32_t* m = (dtype == NLB200_F64) ? h->gp64.mesh : h->gp32.mesh;
  for (int d = 0; d < 3; d++) h->mesh[d] = h->gmesh[d] = m[d];
  h->n_cells = (int64_t)m[0] * m[1] * m[2];
  *out = h;
  return NLB200_OK;
}

int nlb200_set_option(nlb200_handle h, int option, int64_t value) {
  if (!h) return NLB200_ERR_INVALID;
  if (h->initialized) return fail(h, NLB200_ERR_STATE, "options must be set before nlb200_initialize");
  switch (option) {
    case NLB200_OPT_POSITION_STRIDE:
      if (value != 3 && value != 4) return fail(h, NLB200_ERR_INVALID, "position stride must be 3 or 4");
      h->stride = (int)value;
      return NLB200_OK;
    case NLB200_OPT_SORT_ROWS: h->sort_rows = value != 0; return NLB200_OK;
    case NLB200_OPT_ELL_ROWS:
      if (value < 1 || value > 65536) return fail(h, NLB200_ERR_INVALID, "ell rows out of range");
      h->ell_rows = (int)value;
      return NLB200_OK;
    case NLB200_OPT_EXACT_ONLY: h->exact_only = value != 0; return NLB200_OK;
    case NLB200_OPT_USE_GRAPH: h->use_graph = value != 0; return NLB200_OK;
    case NLB200_OPT_KERNEL_VARIANT: h->variant = (int)value; return NLB200_OK;
    case NLB200_OPT_PROFILE: h->profile = value != 0; return NLB200_OK;
    case NLB200_OPT_PDL: h->pdl = value != 0; return NLB200_OK;
    case NLB200_OPT_MAX_IN_CELL:
      if (value < 0 || value > (1 << 20)) return fail(h, NLB200_ERR_INVALID, "max particles per cell out of range");
      h->max_in_cell_opt = value;
      return NLB200_OK;
  }
  return fail(h, NLB200_ERR_INVALID, "unknown option %d", option);
}

int nlb200_set_cell_window(nlb200_handle h, int axis, int32_t first_cell, int32_t n_cells) {
  if (!h) return NLB200_ERR_INVALID;
  if (h->initialized) return fail(h, NLB200_ERR_STATE, "the cell window must be set before nlb200_initialize");
  if (axis < 0 || axis > 2) return fail(h, NLB200_ERR_INVALID, "axis out of range");
  const int32_t gm = h->gmesh[axis];
  if (first_cell < 0 || n_cells < 1 || first_cell + (int64_t)n_cells > gm)
    return fail(h, NLB200_ERR_INVALID, "cell window [%d, %d) outside the grid of %d cells", first_cell,
                first_cell + n_cells, gm);
  // a 3-cell axis is special-cased as fully connected (the reference's wrapped stencil, axis_range): a WINDOW of
  // exactly 3 cells of a longer axis must not be
  if (n_cells == 3 && gm != 3) return fail(h, NLB200_ERR_INVALID, "a cell window needs 1, 2 or at least 4 cells");
  auto apply = [&](auto& gp) {
    gp.mesh[axis] = n_cells;
    gp.coff[axis] = first_cell;
    gp.n_cells = gp.mesh[0] * gp.mesh[1] * gp.mesh[2];
  };
  if (h->dtype == NLB200_F64) apply(h->gp64); else apply(h->gp32);
  h->mesh[axis] = n_cells;
  h->n_cells = (int64_t)h->mesh[0] * h->mesh[1] * h->mesh[2];
  return NLB200_OK;
}

int nlb200_initialize(nlb200_handle h, int64_t max_particles, int64_t max_entries) {
  if (!h) return NLB200_ERR_INVALID;
  if (max_particles < 0 || max_particles > 2147483647ll - 64)
    return fail(h, NLB200_ERR_INVALID, "max_particles out of range (int32 ids)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(h, NLB200_ERR_CUDA, "no CUDA device: libnlist_b200 has no CPU fallback");
  CK(h, cudaGetDevice(&h->device));
  CK(h, set_search_attrs());
  CK(h, set_rowmask_attrs());
  CK(h, set_rowmask4_attrs());
  CK(h, set_runmask_attrs());
  free_buffers(h);
  const int64_t n = max_particles > 0 ? max_particles : 1;
  const int64_t M = h->n_cells;
  h->max_n = max_particles;
  h->tiles_cells = (M + SCAN_TILE - 1) / SCAN_TILE + 1;
  h->tiles_n = (n + SCAN_TILE - 1) / SCAN_TILE + 1;
  // one zeroed region per build: histogram, both scans' look-back state, status block
  size_t off = 0;
  const size_t o_count = off;
  off = align_up(off + sizeof(int32_t) * (size_t)(M + 1), 256);
  const size_t o_sc = off;
  off = align_up(off + sizeof(unsigned long long) * (size_t)(h->tiles_cells + 1), 256);
  const size_t o_sn = off;
  off = align_up(off + sizeof(unsigned long long) * (size_t)(h->tiles_n + 1), 256);
  const size_t o_st = off;
  off = align_up(off + sizeof(DeviceStatus), 256);
  const size_t o_q = off;
  off = align_up(off + sizeof(unsigned long long), 256);
  const size_t o_t = off;
  off = align_up(off + sizeof(unsigned int), 256);
  h->zero_bytes = off;
  h->status_off = o_st;
  CK(h, cudaMalloc(&h->zero_region, off));
  CK(h, cudaMemset(h->zero_region, 0, off));
  h->state_clean = true;
  h->cell_count = reinterpret_cast<int32_t*>(h->zero_region + o_count);
  h->scan_state_cells = reinterpret_cast<unsigned long long*>(h->zero_region + o_sc);
  h->scan_state_counts = reinterpret_cast<unsigned long long*>(h->zero_region + o_sn);
  h->status_dev = reinterpret_cast<DeviceStatus*>(h->zero_region + o_st);
  h->queue = reinterpret_cast<unsigned long long*>(h->zero_region + o_q);
  h->ticket = reinterpret_cast<unsigned int*>(h->zero_region + o_t);
  CK(h, cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->device));
  {
    int l2 = 0;
    CK(h, cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, h->device));
    h->l2_bytes = l2;
  }
  CK(h, cudaMalloc(&h->cell_start, sizeof(int32_t) * (size_t)(M + 1)));
  CK(h, cudaMalloc(&h->cell_rank, sizeof(int2) * (size_t)n));
  CK(h, cudaMalloc(&h->perm, sizeof(int32_t) * (size_t)n));
  CK(h, cudaMalloc(&h->sorted_ids, sizeof(int32_t) * (size_t)n));
  CK(h, cudaMalloc(&h->slot_cell, sizeof(int32_t) * (size_t)n));
  CK(h, cudaMalloc(&h->slot_gid, sizeof(int32_t) * (size_t)n));
  CK(h, cudaMalloc(&h->rec, sizeof(float4) * (size_t)n));
  CK(h, cudaMalloc(&h->counts, sizeof(int32_t) * (size_t)(n + 8)));
  CK(h, cudaMalloc(&h->offsets, sizeof(int64_t) * (size_t)(n + 1)));
  CK(h, cudaMalloc(&h->offsets32, sizeof(int32_t) * (size_t)(n + 1)));
  CK(h, cudaMemset(h->offsets, 0, sizeof(int64_t) * (size_t)(n + 1)));
  CK(h, cudaMemset(h->offsets32, 0, sizeof(int32_t) * (size_t)(n + 1)));
  const int64_t entries = max_entries > 0 ? max_entries : estimate_entries(h, n);
  int rc = alloc_partners(h, entries);
  if (rc) return rc;
  {
    const int64_t mic = h->max_in_cell_opt > 0 ? h->max_in_cell_opt : estimate_max_in_cell(h, n);
    h->path = pick_path(h, mic);
    rc = alloc_search_buffers(h, mic, h->max_in_cell_opt > 0);
    if (rc) return rc;
  }
  if (h->mode == NLB200_FULL_ELL_TRANSPOSED) {
    // neighlist_gpu.hpp:102,271-274: MAX_PARTNERS * N ints, filled with -1 once
    const int64_t tot = (int64_t)h->ell_rows * n;
    CK(h, cudaMalloc(&h->ell, sizeof(int32_t) * (size_t)tot));
    CK(h, cudaMalloc(&h->ell_prev, sizeof(int32_t) * (size_t)n));
    h->ell_last_n = -1;  // first build fills it with -1
  }
  CK(h, cudaMallocHost(&h->status_host, sizeof(DeviceStatus)));
  memset(h->status_host, 0, sizeof(DeviceStatus));
  CK(h, cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  CK(h, cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
  CK(h, cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  CK(h, cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  if (h->profile)
    for (int k = 0; k <= nlb200_context::MAX_STAGES; k++) CK(h, cudaEventCreate(&h->ev[k]));
  CK(h, cudaDeviceSynchronize());
  h->initialized = true;
  h->have_result = false;
  h->build_pending = false;
  h->have_build = false;
  return NLB200_OK;
}

int nlb200_reserve(nlb200_handle h, int64_t max_entries) {
  if (!h || !h->initialized) return h ? fail(h, NLB200_ERR_STATE, "reserve before initialize") : NLB200_ERR_INVALID;
  if (max_entries <= h->cap_entries) return NLB200_OK;
  CK(h, settle_before_realloc(h));
  return alloc_partners(h, max_entries);
}

int nlb200_reserve_cell_capacity(nlb200_handle h, int64_t max_in_cell) {
  if (!h || !h->initialized) return h ? fail(h, NLB200_ERR_STATE, "reserve before initialize") : NLB200_ERR_INVALID;
  if (uses_v1(h)) return NLB200_OK;
  if (uses_rowmask(h)) {
    // the row-mask buffer holds one bit per test: grow it to what the failed build asked for (+ 1/16)
    const int64_t need = h->rmask_need + h->rmask_need / 16 + 4096;
    if (need <= h->rmask_cap) return NLB200_OK;
    CK(h, settle_before_realloc(h));
    return alloc_rmask(h, need);
  }
  if (uses_runmask(h) ? 3 * max_in_cell <= (int64_t)h->mask_wr * 32 : max_in_cell <= (int64_t)h->mask_wi * 32)
    return NLB200_OK;
  CK(h, settle_before_realloc(h));
  if (pick_path(h, max_in_cell) == PATH_ROWMASK) {
    // crowded cells: the per-particle planes of the pair masks would grow with the most crowded cell; the row masks
    // grow with the number of tests.  The first build on the new path reports how many words it needs.
    h->path = PATH_ROWMASK;
    h->rmask_need = 0;
    return alloc_search_buffers(h, max_in_cell, true);
  }
  if (uses_runmask(h)) return alloc_runmask(h, 3 * max_in_cell);
  return alloc_mask(h, max_in_cell);
}

int nlb200_destroy(nlb200_handle h) {
  if (!h) return NLB200_OK;
  if (h->build_pending && h->last_stream) cudaStreamSynchronize(h->last_stream);
  free_buffers(h);
  delete h;
  return NLB200_OK;
}

int nlb200_build_subset(nlb200_handle h, const void* q_dev, int64_t n_total, int64_t n_owned,
                        const int32_t* global_ids_dev, void* stream) {
  if (!h) return NLB200_ERR_INVALID;
  if (!h->initialized) return fail(h, NLB200_ERR_STATE, "build before initialize");
  if (n_total < 0 || n_owned < 0 || n_owned > n_total || n_total > h->max_n)
    return fail(h, NLB200_ERR_INVALID, "particle count %lld (owned %lld) outside [0, %lld]", (long long)n_total,
                (long long)n_owned, (long long)h->max_n);
  if (n_total > 0 && q_dev == nullptr) return fail(h, NLB200_ERR_INVALID, "null position pointer");
  (void)cudaGetLastError();  // a non-sticky error left behind by another library must not be reported as ours
  const size_t align = (h->stride == 4) ? 16 : (h->dtype == NLB200_F64 ? 8 : 4);
  if ((reinterpret_cast<uintptr_t>(q_dev) % align) != 0)
    return fail(h, NLB200_ERR_INVALID, "position pointer must be %zu-byte aligned", align);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int rc = NLB200_OK;
  if (h->mode == NLB200_FULL_ELL_TRANSPOSED && n_owned != h->ell_last_n) {
    // the view's row stride is the particle count (list[k*N + i], kernel_impl.cuh:30): a new N re-lays the matrix
    // out, so start again from the reference's initial state (all -1, neighlist_gpu.hpp:271-274)
    fill_i32_kernel<<<1024, 256, 0, s>>>(h->ell, (int64_t)h->ell_rows * (h->max_n > 0 ? h->max_n : 1), -1);
    CK(h, cudaGetLastError());
    CK(h, cudaMemsetAsync(h->ell_prev, 0, sizeof(int32_t) * (size_t)(h->max_n > 0 ? h->max_n : 1), s));
    h->ell_last_n = n_owned;
  }
  bool can_graph = h->use_graph && !h->profile && s != nullptr && s != cudaStreamLegacy && s != cudaStreamPerThread;
  if (can_graph) {
    // the caller is capturing this stream into a graph of its own (e.g. halo exchange + build as one graph): enqueue
    // the plain kernel chain, which becomes part of that graph
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) can_graph = false;
  }
  if (can_graph) {
    const bool hit = h->graph_exec && h->g_q == q_dev && h->g_n == n_total && h->g_owned == n_owned &&
                     h->g_gids == global_ids_dev;
    if (!hit) {
      drop_graph(h);
      cudaGraph_t graph = nullptr;
      cudaError_t e = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
      if (e == cudaSuccess) {
        rc = enqueue_dispatch(h, q_dev, n_total, n_owned, global_ids_dev, s);
        e = cudaStreamEndCapture(s, &graph);
        if (rc == NLB200_OK && e == cudaSuccess && graph) {
          e = cudaGraphInstantiate(&h->graph_exec, graph, 0);
          if (e == cudaSuccess) {
            h->g_q = q_dev;
            h->g_n = n_total;
            h->g_owned = n_owned;
            h->g_gids = global_ids_dev;
          } else {
            h->graph_exec = nullptr;
          }
        }
        if (graph) cudaGraphDestroy(graph);
        if (rc != NLB200_OK) {
          drop_graph(h);
          return rc;
        }
      }
      (void)cudaGetLastError();
    }
    if (h->graph_exec) {
      CK(h, cudaGraphLaunch(h->graph_exec, s));
    } else {
      rc = enqueue_dispatch(h, q_dev, n_total, n_owned, global_ids_dev, s);
    }
  } else {
    rc = enqueue_dispatch(h, q_dev, n_total, n_owned, global_ids_dev, s);
  }
  if (rc != NLB200_OK) {
    drop_graph(h);  // a failed enqueue leaves the per-build state dirty: the next build clears it itself
    return rc;
  }
  h->last_stream = s;
  h->build_pending = true;
  h->have_result = false;
  h->last_n = n_total;
  h->last_owned = n_owned;
  h->have_build = true;
  return NLB200_OK;
}

int nlb200_mark_enqueued(nlb200_handle h, void* stream) {
  if (!h) return NLB200_ERR_INVALID;
  if (!h->initialized || !h->have_build) return fail(h, NLB200_ERR_STATE, "no build to mark");
  h->last_stream = reinterpret_cast<cudaStream_t>(stream);
  h->build_pending = true;
  h->have_result = false;
  return NLB200_OK;
}

int nlb200_build(nlb200_handle h, const void* q_dev, int64_t n, void* stream) {
  return nlb200_build_subset(h, q_dev, n, n, nullptr, stream);
}

int nlb200_synchronize(nlb200_handle h) {
  if (!h) return NLB200_ERR_INVALID;
  if (!h->initialized) return fail(h, NLB200_ERR_STATE, "synchronize before initialize");
  if (!h->build_pending && !h->have_result) return fail(h, NLB200_ERR_STATE, "no build to synchronize");
  if (h->build_pending) {
    CK(h, cudaStreamSynchronize(h->last_stream));
    h->build_pending = false;
  }
  const DeviceStatus st = *h->status_host;
  h->stats.n = h->last_owned;
  h->stats.number_of_pairs = (int64_t)st.total_entries;
  h->stats.candidates_tested = (int64_t)st.candidates;
  h->stats.band_tests = (int64_t)st.band_tests;
  h->stats.required_entries = (int64_t)st.total_entries;
  h->stats.capacity_entries = h->cap_entries;
  for (int d = 0; d < 3; d++) h->stats.mesh[d] = h->gmesh[d];
  h->stats.max_partners = st.max_partners;
  h->stats.max_in_cell = st.max_in_cell;
  h->rmask_need = (int64_t)st.mask_words;
  h->have_result = true;
  if (st.flags & FLAG_OUT_OF_BOX)
    return fail(h, NLB200_ERR_OUT_OF_BOX, "a particle lies more than one cell outside [0,L] or is NaN");
  if (st.flags & FLAG_MASK_WORDS)
    return fail(h, NLB200_ERR_CELL_CAPACITY, "the row masks need %lld words, the buffer holds %lld (most crowded cell: %d)",
                (long long)st.mask_words, (long long)h->rmask_cap, st.max_in_cell);
  if (st.flags & FLAG_CELL_WORDS)
    return uses_runmask(h)
               ? fail(h, NLB200_ERR_CELL_CAPACITY, "a run of three cells holds more than the %d particles the run-mask words cover (most crowded cell: %d)",
                      h->mask_wr * 32, st.max_in_cell)
               : fail(h, NLB200_ERR_CELL_CAPACITY, "a cell holds %d particles, the pair-mask words cover %d", st.max_in_cell,
                      h->mask_wi * 32);
  if (st.flags & FLAG_CAPACITY)
    return fail(h, NLB200_ERR_CAPACITY, "partner list needs %lld entries, capacity is %lld",
                (long long)st.total_entries, (long long)h->cap_entries);
  if (st.flags & FLAG_ELL_ROWS)
    return fail(h, NLB200_ERR_ELL_ROWS, "a row has %d partners, ELL row capacity is %d", st.max_partners,
                h->ell_rows);
  return NLB200_OK;
}

int nlb200_build_host(nlb200_handle h, const void* q_host, int64_t n, int32_t* number_of_partners_host,
                      int64_t* offsets_host, int32_t* partners_host, int64_t partners_capacity,
                      int64_t* number_of_pairs) {
  if (!h) return NLB200_ERR_INVALID;
  if (!h->initialized) return fail(h, NLB200_ERR_STATE, "build before initialize");
  if (n < 0 || n > h->max_n) return fail(h, NLB200_ERR_INVALID, "particle count out of range");
  const size_t esz = h->dtype == NLB200_F64 ? 8 : 4;
  const size_t bytes = (size_t)n * h->stride * esz;
  if (!h->q_stage) CK(h, cudaMalloc(&h->q_stage, (size_t)(h->max_n > 0 ? h->max_n : 1) * h->stride * esz));
  cudaStream_t s = h->own_stream;
  if (bytes) CK(h, cudaMemcpyAsync(h->q_stage, q_host, bytes, cudaMemcpyHostToDevice, s));
  int rc = nlb200_build(h, h->q_stage, n, s);
  if (rc) return rc;
  rc = nlb200_synchronize(h);
  for (int attempt = 0; attempt < 3 && rc == NLB200_ERR_CELL_CAPACITY; attempt++) {
    // crowded cells: grow the masks (the first growth may move the handle to the row-mask path, whose first build
    // then reports how many words it needs)
    rc = nlb200_reserve_cell_capacity(h, h->stats.max_in_cell);
    if (rc) return rc;
    rc = nlb200_build(h, h->q_stage, n, s);
    if (rc) return rc;
    rc = nlb200_synchronize(h);
  }
  if (rc == NLB200_ERR_CAPACITY) {
    // grow and retry once: the reference's answer to a full buffer is undefined behaviour
    const int64_t need = h->stats.required_entries;
    rc = nlb200_reserve(h, need + need / 16 + 1024);
    if (rc) return rc;
    rc = nlb200_build(h, h->q_stage, n, s);
    if (rc) return rc;
    rc = nlb200_synchronize(h);
  }
  if (rc) return rc;
  const int64_t total = h->stats.number_of_pairs;
  if (number_of_pairs) *number_of_pairs = total;
  if (number_of_partners_host && n)
    CK(h, cudaMemcpyAsync(number_of_partners_host, h->counts, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost,
                          s));
  if (offsets_host)
    CK(h, cudaMemcpyAsync(offsets_host, h->offsets, sizeof(int64_t) * (size_t)(n + 1), cudaMemcpyDeviceToHost, s));
  if (partners_host) {
    if (total > partners_capacity)
      return fail(h, NLB200_ERR_CAPACITY, "host partner buffer holds %lld entries, %lld needed",
                  (long long)partners_capacity, (long long)total);
    if (total)
      CK(h, cudaMemcpyAsync(partners_host, h->partners, sizeof(int32_t) * (size_t)total, cudaMemcpyDeviceToHost, s));
  }
  CK(h, cudaStreamSynchronize(s));
  return NLB200_OK;
}

int nlb200_fetch_partners_host(nlb200_handle h, int32_t* partners_host, int64_t capacity) {
  if (!h) return NLB200_ERR_INVALID;
  if (!h->have_result) return fail(h, NLB200_ERR_STATE, "no synchronized build to fetch");
  const int64_t total = h->stats.number_of_pairs;
  if (total > capacity || (total > 0 && !partners_host))
    return fail(h, NLB200_ERR_CAPACITY, "host partner buffer holds %lld entries, %lld needed", (long long)capacity,
                (long long)total);
  if (total) CK(h, cudaMemcpy(partners_host, h->partners, sizeof(int32_t) * (size_t)total, cudaMemcpyDeviceToHost));
  return NLB200_OK;
}

const int32_t* nlb200_number_of_partners(nlb200_handle h) { return h ? h->counts : nullptr; }
const int64_t* nlb200_offsets(nlb200_handle h) { return h ? h->offsets : nullptr; }
const int32_t* nlb200_offsets32(nlb200_handle h) {
  if (!h) return nullptr;
  if (h->have_result && h->stats.number_of_pairs > 2147483647ll) return nullptr;
  return h->offsets32;
}
const int32_t* nlb200_partners(nlb200_handle h) { return h ? h->partners : nullptr; }
const int32_t* nlb200_ell_transposed(nlb200_handle h) { return h ? h->ell : nullptr; }
const int32_t* nlb200_cell_start(nlb200_handle h) { return h ? h->cell_start : nullptr; }
const int32_t* nlb200_sorted_ids(nlb200_handle h) { return h ? h->sorted_ids : nullptr; }

int64_t nlb200_number_of_pairs(nlb200_handle h) {
  if (!h || !h->have_result) return -1;
  return h->stats.number_of_pairs;
}

int nlb200_get_stats(nlb200_handle h, nlb200_stats* out) {
  if (!h || !out) return NLB200_ERR_INVALID;
  if (!h->have_result) {
    nlb200_stats s{};
    for (int d = 0; d < 3; d++) s.mesh[d] = h->gmesh[d];
    s.capacity_entries = h->cap_entries;
    *out = s;
    return NLB200_OK;
  }
  *out = h->stats;
  return NLB200_OK;
}

int nlb200_get_stage_times(nlb200_handle h, float* ms_out, int32_t* stage_ids_out, int capacity) {
  if (!h || !h->profile || h->build_pending) return -1;
  int k = 0;
  for (; k < h->n_stages && k < capacity; k++) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev[k], h->ev[k + 1]) != cudaSuccess) return -1;
    ms_out[k] = ms;
    stage_ids_out[k] = h->stage_id[k];
  }
  return k;
}

const char* nlb200_stage_name(int stage_id) {
  return (stage_id >= 0 && stage_id < ST_NUM) ? kStageNames[stage_id] : "unknown";
}

int64_t nlb200_required_entries(nlb200_handle h) { return (h && h->have_result) ? h->stats.required_entries : -1; }

const char* nlb200_last_error(nlb200_handle h) { return h ? h->err.c_str() : "null handle"; }

// ---- callers either side of the build ------------------------------------------------------------

int nlb200_track_reference(nlb200_handle h, const void* q_dev, int64_t n, void* stream) {
  if (!h) return NLB200_ERR_INVALID;
  (void)cudaGetLastError();
  if (!h->initialized) return fail(h, NLB200_ERR_STATE, "track before initialize");
  if (n < 0 || n > h->max_n || (n > 0 && !q_dev)) return fail(h, NLB200_ERR_INVALID, "particle count out of range");
  const size_t esz = h->dtype == NLB200_F64 ? 8 : 4;
  if (!h->q_ref) {
    CK(h, cudaMalloc(&h->q_ref, (size_t)(h->max_n > 0 ? h->max_n : 1) * h->stride * esz));
    CK(h, cudaMalloc(&h->disp_dev, sizeof(unsigned long long)));
    CK(h, cudaMallocHost(&h->disp_host, sizeof(unsigned long long)));
  }
  if (n) CK(h, cudaMemcpyAsync(h->q_ref, q_dev, (size_t)n * h->stride * esz, cudaMemcpyDeviceToDevice,
                               reinterpret_cast<cudaStream_t>(stream)));
  h->q_ref_n = n;
  return NLB200_OK;
}

int nlb200_max_displacement(nlb200_handle h, const void* q_dev, int64_t n, void* stream, double* max_disp_host) {
  if (!h || !max_disp_host) return NLB200_ERR_INVALID;
  (void)cudaGetLastError();
  if (!h->q_ref) return fail(h, NLB200_ERR_STATE, "nlb200_max_displacement before nlb200_track_reference");
  if (n != h->q_ref_n) return fail(h, NLB200_ERR_INVALID, "particle count differs from the tracked reference");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  CK(h, cudaMemsetAsync(h->disp_dev, 0, sizeof(unsigned long long), s));
  if (n > 0) {
    const unsigned g = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 8);
    if (h->dtype == NLB200_F64)
      max_disp2_kernel<double><<<g, 256, 0, s>>>((const double*)q_dev, (const double*)h->q_ref, n, h->stride,
                                                h->disp_dev);
    else
      max_disp2_kernel<float><<<g, 256, 0, s>>>((const float*)q_dev, (const float*)h->q_ref, n, h->stride,
                                               h->disp_dev);
    CK(h, cudaGetLastError());
  }
  CK(h, cudaMemcpyAsync(h->disp_host, h->disp_dev, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  CK(h, cudaStreamSynchronize(s));
  double d2;
  memcpy(&d2, h->disp_host, sizeof(double));
  *max_disp_host = std::sqrt(d2);
  return NLB200_OK;
}

int nlb200_lj_forces(nlb200_handle h, const void* q_dev, double rc, double epsilon, double sigma, double* forces_dev,
                     double* energy_dev, void* stream) {
  if (!h) return NLB200_ERR_INVALID;
  (void)cudaGetLastError();
  if (!h->initialized || (!h->build_pending && !h->have_result))
    return fail(h, NLB200_ERR_STATE, "lj_forces needs a build");
  if (h->mode == NLB200_HALF_CSR) return fail(h, NLB200_ERR_INVALID, "lj_forces reads FULL rows");
  if (!(rc > 0) || rc > h->sl || !q_dev || !forces_dev)
    return fail(h, NLB200_ERR_INVALID, "lj_forces: 0 < rc <= search length, non-null buffers");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int32_t n = (int32_t)h->last_owned;
  if (n == 0) return NLB200_OK;
  const unsigned g = (unsigned)(((int64_t)n * 32 + 127) / 128);
  if (h->dtype == NLB200_F64)
    lj_forces_kernel<double><<<g, 128, 0, s>>>((const double*)q_dev, h->stride, n, h->offsets, h->partners, rc * rc,
                                               epsilon, sigma * sigma, forces_dev, energy_dev);
  else
    lj_forces_kernel<float><<<g, 128, 0, s>>>((const float*)q_dev, h->stride, n, h->offsets, h->partners, rc * rc,
                                              epsilon, sigma * sigma, forces_dev, energy_dev);
  CK(h, cudaGetLastError());
  return NLB200_OK;
}

int nlb200_gather_sorted(nlb200_handle h, const void* src_dev, int elem_bytes, int width, void* dst_dev, void* stream) {
  if (!h) return NLB200_ERR_INVALID;
  (void)cudaGetLastError();
  if (!h->initialized || (!h->build_pending && !h->have_result))
    return fail(h, NLB200_ERR_STATE, "gather_sorted needs a build (the cell order comes from it)");
  if ((elem_bytes != 4 && elem_bytes != 8) || width < 1 || width > 16 || !src_dev || !dst_dev)
    return fail(h, NLB200_ERR_INVALID, "gather_sorted: 4- or 8-byte elements, 1..16 per particle");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t total = h->last_n * width;
  if (total == 0) return NLB200_OK;
  const unsigned g = (unsigned)((total + 255) / 256);
  const int32_t* present = h->cell_start + h->n_cells;
  if (elem_bytes == 8)
    gather_sorted_kernel<double><<<g, 256, 0, s>>>((const double*)src_dev, h->sorted_ids, present, width,
                                                  (double*)dst_dev);
  else
    gather_sorted_kernel<float><<<g, 256, 0, s>>>((const float*)src_dev, h->sorted_ids, present, width,
                                                 (float*)dst_dev);
  CK(h, cudaGetLastError());
  return NLB200_OK;
}

// ---- adjacent utilities ----------------------------------------------------------------------------------------

int nlb200_select_slab(const void* q_dev, int64_t n, int dtype, int stride, int axis, double lo, double hi,
                       int32_t* out_idx_dev, int64_t capacity, int64_t* out_count_dev, void* workspace_dev,
                       int64_t workspace_bytes, void* stream) {
  if (n < 0 || axis < 0 || axis > 2 || (stride != 3 && stride != 4)) return NLB200_ERR_INVALID;
  // workspace: flags int32[n+8] | positions int64[n+1] | scan state u64[tiles+2]
  const int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE + 1;
  const size_t o_flags = 0;
  const size_t o_pos = align_up(sizeof(int32_t) * (size_t)(n + 8), 256);
  const size_t o_state = align_up(o_pos + sizeof(int64_t) * (size_t)(n + 1), 256);
  const size_t need = o_state + sizeof(unsigned long long) * (size_t)(tiles + 2);
  if (workspace_dev == nullptr || (size_t)workspace_bytes < need) return NLB200_ERR_CAPACITY;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace_dev);
  int32_t* flags = reinterpret_cast<int32_t*>(ws + o_flags);
  int64_t* pos = reinterpret_cast<int64_t*>(ws + o_pos);
  unsigned long long* state = reinterpret_cast<unsigned long long*>(ws + o_state);
  if (cudaMemsetAsync(state, 0, sizeof(unsigned long long) * (size_t)(tiles + 2), s) != cudaSuccess)
    return NLB200_ERR_CUDA;
  const unsigned g = (unsigned)((n + 255) / 256);
  if (n > 0) {
    if (dtype == NLB200_F64)
      slab_flag_kernel<double><<<g, 256, 0, s>>>((const double*)q_dev, n, stride, axis, lo, hi, flags);
    else
      slab_flag_kernel<float><<<g, 256, 0, s>>>((const float*)q_dev, n, stride, axis, lo, hi, flags);
  }
  scan_kernel<int64_t><<<(unsigned)(tiles - 1 > 0 ? tiles - 1 : 1), SCAN_THREADS, 0, s>>>(flags, n, pos, nullptr, state,
                                                                                         nullptr, nullptr, 0);
  slab_compact_kernel<<<g > 0 ? g : 1, 256, 0, s>>>(flags, pos, n, out_idx_dev, capacity, out_count_dev);
  return cudaGetLastError() == cudaSuccess ? NLB200_OK : NLB200_ERR_CUDA;
}

int nlb200_pack_slab(const void* q_dev, const int32_t* gids_dev, int32_t gid_base, int64_t n, int dtype, int stride,
                     int axis, double lo, double hi, void* out_q_dev, int32_t* out_gid_dev, int64_t capacity,
                     int64_t* out_count_dev, void* workspace_dev, int64_t workspace_bytes, void* stream) {
  if (n < 0 || capacity < 0 || axis < 0 || axis > 2 || (stride != 3 && stride != 4)) return NLB200_ERR_INVALID;
  const int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE + 1;
  const size_t o_pos = align_up(sizeof(int32_t) * (size_t)(n + 8), 256);
  const size_t o_state = align_up(o_pos + sizeof(int64_t) * (size_t)(n + 1), 256);
  const size_t need = o_state + sizeof(unsigned long long) * (size_t)(tiles + 2);
  if (workspace_dev == nullptr || (size_t)workspace_bytes < need) return NLB200_ERR_CAPACITY;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace_dev);
  int32_t* flags = reinterpret_cast<int32_t*>(ws);
  int64_t* pos = reinterpret_cast<int64_t*>(ws + o_pos);
  unsigned long long* state = reinterpret_cast<unsigned long long*>(ws + o_state);
  const size_t esz = dtype == NLB200_F64 ? 8 : 4;
  // all-ones bytes are a NaN in both precisions: every slot starts as an absent ghost
  if (capacity > 0 && cudaMemsetAsync(out_q_dev, 0xFF, (size_t)capacity * stride * esz, s) != cudaSuccess)
    return NLB200_ERR_CUDA;
  if (cudaMemsetAsync(state, 0, sizeof(unsigned long long) * (size_t)(tiles + 2), s) != cudaSuccess)
    return NLB200_ERR_CUDA;
  const unsigned g = (unsigned)((n + 255) / 256);
  if (n > 0) {
    if (dtype == NLB200_F64)
      slab_flag_kernel<double><<<g, 256, 0, s>>>((const double*)q_dev, n, stride, axis, lo, hi, flags);
    else
      slab_flag_kernel<float><<<g, 256, 0, s>>>((const float*)q_dev, n, stride, axis, lo, hi, flags);
  }
  scan_kernel<int64_t><<<(unsigned)(tiles - 1 > 0 ? tiles - 1 : 1), SCAN_THREADS, 0, s>>>(flags, n, pos, nullptr, state,
                                                                                         nullptr, nullptr, 0);
  if (dtype == NLB200_F64)
    slab_pack_kernel<double><<<g > 0 ? g : 1, 256, 0, s>>>((const double*)q_dev, gids_dev, gid_base, flags, pos, n,
                                                          stride, (double*)out_q_dev, out_gid_dev, capacity,
                                                          out_count_dev);
  else
    slab_pack_kernel<float><<<g > 0 ? g : 1, 256, 0, s>>>((const float*)q_dev, gids_dev, gid_base, flags, pos, n, stride,
                                                         (float*)out_q_dev, out_gid_dev, capacity, out_count_dev);
  return cudaGetLastError() == cudaSuccess ? NLB200_OK : NLB200_ERR_CUDA;
}

int nlb200_pack_slab2(const void* q_dev, const int32_t* gids_dev, int64_t n, int dtype, int stride, int axis,
                      double cut_lo, double cut_hi, void* out_q_lo_dev, int32_t* out_gid_lo_dev, void* out_q_hi_dev,
                      int32_t* out_gid_hi_dev, int64_t capacity, int64_t* out_counts_dev, void* workspace_dev,
                      int64_t workspace_bytes, void* stream) {
  if (n < 0 || capacity < 0 || axis < 0 || axis > 2 || (stride != 3 && stride != 4)) return NLB200_ERR_INVALID;
  (void)cudaGetLastError();  // a non-sticky error left behind by another library must not be reported as ours
  const int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE + 1;
  // workspace: [flags_lo | pos_lo] [flags_hi | pos_hi] [scan state lo | scan state hi]
  const size_t o_pos = align_up(sizeof(int32_t) * (size_t)(n + 8), 256);
  const size_t blk = align_up(o_pos + sizeof(int64_t) * (size_t)(n + 1), 256);
  const size_t st_sz = align_up(sizeof(unsigned long long) * (size_t)(tiles + 2), 256);
  if (workspace_dev == nullptr || (size_t)workspace_bytes < 2 * blk + 2 * st_sz) return NLB200_ERR_CAPACITY;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace_dev);
  int32_t* flags[2] = {reinterpret_cast<int32_t*>(ws), reinterpret_cast<int32_t*>(ws + blk)};
  int64_t* pos[2] = {reinterpret_cast<int64_t*>(ws + o_pos), reinterpret_cast<int64_t*>(ws + blk + o_pos)};
  unsigned long long* state[2] = {reinterpret_cast<unsigned long long*>(ws + 2 * blk),
                                  reinterpret_cast<unsigned long long*>(ws + 2 * blk + st_sz)};
  // one memset clears both look-back states; the NaN padding of the unused slots is written by the packing kernel
  if (cudaMemsetAsync(state[0], 0, 2 * st_sz, s) != cudaSuccess) return NLB200_ERR_CUDA;
  const unsigned g = (unsigned)((n + 255) / 256);
  if (n > 0) {
    if (dtype == NLB200_F64)
      slab_flag2_kernel<double><<<g, 256, 0, s>>>((const double*)q_dev, n, stride, axis, cut_lo, cut_hi, flags[0],
                                                 flags[1]);
    else
      slab_flag2_kernel<float><<<g, 256, 0, s>>>((const float*)q_dev, n, stride, axis, cut_lo, cut_hi, flags[0],
                                                flags[1]);
  }
  for (int f = 0; f < 2; f++)
    scan_kernel<int64_t><<<(unsigned)(tiles - 1 > 0 ? tiles - 1 : 1), SCAN_THREADS, 0, s>>>(
        flags[f], n, pos[f], nullptr, state[f], nullptr, nullptr, 0);
  const int64_t span = n > capacity ? n : capacity;
  const unsigned gp = (unsigned)((span + 255) / 256 > 0 ? (span + 255) / 256 : 1);
  if (dtype == NLB200_F64)
    slab_pack2_kernel<double><<<gp, 256, 0, s>>>(
        (const double*)q_dev, gids_dev, flags[0], pos[0], flags[1], pos[1], n, stride, (double*)out_q_lo_dev,
        out_gid_lo_dev, (double*)out_q_hi_dev, out_gid_hi_dev, capacity, out_counts_dev);
  else
    slab_pack2_kernel<float><<<gp, 256, 0, s>>>(
        (const float*)q_dev, gids_dev, flags[0], pos[0], flags[1], pos[1], n, stride, (float*)out_q_lo_dev,
        out_gid_lo_dev, (float*)out_q_hi_dev, out_gid_hi_dev, capacity, out_counts_dev);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    std::fprintf(stderr, "nlb200_pack_slab2: %s (n=%lld capacity=%lld)\n", cudaGetErrorString(e), (long long)n,
                 (long long)capacity);
    return NLB200_ERR_CUDA;
  }
  return NLB200_OK;
}

int nlb200_pack_faces(const void* q_dev, const int32_t* gids_dev, int64_t n, int dtype, int stride, int axis,
                      double cut_lo, double cut_hi, void* out_q_lo_dev, int32_t* out_gid_lo_dev, void* out_q_hi_dev,
                      int32_t* out_gid_hi_dev, int64_t capacity, int64_t* out_counts_dev, void* state_dev,
                      void* stream) {
  if (n < 0 || capacity < 0 || axis < 0 || axis > 2 || (stride != 3 && stride != 4) || !state_dev || !out_counts_dev)
    return NLB200_ERR_INVALID;
  (void)cudaGetLastError();  // a non-sticky error left behind by another library must not be reported as ours
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const unsigned g = (unsigned)((n + 255) / 256 > 0 ? (n + 255) / 256 : 1);
  unsigned long long* st = reinterpret_cast<unsigned long long*>(state_dev);
  if (dtype == NLB200_F64)
    pack_faces_kernel<double><<<g, 256, 0, s>>>((const double*)q_dev, gids_dev, n, stride, axis, cut_lo, cut_hi,
                                               (double*)out_q_lo_dev, out_gid_lo_dev, (double*)out_q_hi_dev,
                                               out_gid_hi_dev, capacity, st, out_counts_dev);
  else
    pack_faces_kernel<float><<<g, 256, 0, s>>>((const float*)q_dev, gids_dev, n, stride, axis, cut_lo, cut_hi,
                                              (float*)out_q_lo_dev, out_gid_lo_dev, (float*)out_q_hi_dev,
                                              out_gid_hi_dev, capacity, st, out_counts_dev);
  return cudaGetLastError() == cudaSuccess ? NLB200_OK : NLB200_ERR_CUDA;
}

// ---- halo exchange by peer stores (CUDA IPC) ---------------------------------------------------------------------
int nlb200_p2p_alloc(int64_t bytes, void** dev_ptr, void* ipc_handle_64) {
  if (bytes <= 0 || !dev_ptr || !ipc_handle_64) return NLB200_ERR_INVALID;
  (void)cudaGetLastError();
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  void* p = nullptr;
  if (cudaMalloc(&p, (size_t)bytes) != cudaSuccess) return NLB200_ERR_CUDA;
  if (cudaMemset(p, 0, (size_t)bytes) != cudaSuccess ||
      cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(ipc_handle_64), p) != cudaSuccess) {
    cudaFree(p);
    (void)cudaGetLastError();
    return NLB200_ERR_CUDA;
  }
  *dev_ptr = p;
  return NLB200_OK;
}

int nlb200_p2p_open(const void* ipc_handle_64, void** peer_ptr) {
  if (!ipc_handle_64 || !peer_ptr) return NLB200_ERR_INVALID;
  (void)cudaGetLastError();
  cudaIpcMemHandle_t hdl;
  memcpy(&hdl, ipc_handle_64, sizeof(hdl));
  if (cudaIpcOpenMemHandle(peer_ptr, hdl, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
    (void)cudaGetLastError();
    return NLB200_ERR_CUDA;
  }
  return NLB200_OK;
}

int nlb200_p2p_close(void* peer_ptr) {
  if (!peer_ptr) return NLB200_OK;
  return cudaIpcCloseMemHandle(peer_ptr) == cudaSuccess ? NLB200_OK : NLB200_ERR_CUDA;
}

int nlb200_p2p_free(void* dev_ptr) {
  if (!dev_ptr) return NLB200_OK;
  return cudaFree(dev_ptr) == cudaSuccess ? NLB200_OK : NLB200_ERR_CUDA;
}

int nlb200_pack_faces_p2p(const void* q_dev, const int32_t* gids_dev, int64_t n, int dtype, int stride, int axis,
                          double cut_lo, double cut_hi, void* peer_q_lo, int32_t* peer_gid_lo, void* peer_q_hi,
                          int32_t* peer_gid_hi, int64_t capacity, int64_t* out_counts_dev, void* state_dev,
                          void* ctrl_dev, void* peer_ready_lo, void* peer_ready_hi, void* stream) {
  if (n < 0 || capacity < 0 || axis < 0 || axis > 2 || (stride != 3 && stride != 4) || !state_dev ||
      !out_counts_dev || !ctrl_dev)
    return NLB200_ERR_INVALID;
  (void)cudaGetLastError();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const unsigned g = (unsigned)((n + 255) / 256 > 0 ? (n + 255) / 256 : 1);
  unsigned long long* st = reinterpret_cast<unsigned long long*>(state_dev);
  HaloCtrl* ctrl = reinterpret_cast<HaloCtrl*>(ctrl_dev);
  unsigned long long* rl = reinterpret_cast<unsigned long long*>(peer_ready_lo);
  unsigned long long* rh = reinterpret_cast<unsigned long long*>(peer_ready_hi);
  HaloPackArgs hp{};
  hp.axis = axis;
  hp.cut_lo = cut_lo;
  hp.cut_hi = cut_hi;
  hp.out_q_lo = peer_q_lo;
  hp.out_gid_lo = peer_gid_lo;
  hp.out_q_hi = peer_q_hi;
  hp.out_gid_hi = peer_gid_hi;
  hp.capacity = capacity;
  hp.state = st;
  hp.out_counts = reinterpret_cast<long long*>(out_counts_dev);
  hp.ctrl = ctrl;
  hp.peer_ready_lo = rl;
  hp.peer_ready_hi = rh;
  if (dtype == NLB200_F64)
    pack_faces_p2p_kernel<double><<<g, 256, 0, s>>>((const double*)q_dev, gids_dev, n, stride, hp);
  else
    pack_faces_p2p_kernel<float><<<g, 256, 0, s>>>((const float*)q_dev, gids_dev, n, stride, hp);
  return cudaGetLastError() == cudaSuccess ? NLB200_OK : NLB200_ERR_CUDA;
}

int nlb200_set_halo_pack(nlb200_handle h, int axis, double cut_lo, double cut_hi, void* peer_q_lo, int32_t* peer_gid_lo,
                         void* peer_q_hi, int32_t* peer_gid_hi, int64_t capacity, int64_t* out_counts_dev,
                         void* state_dev, void* peer_ready_lo, void* peer_ready_hi, int32_t* send_idx_lo_dev,
                         int32_t* send_idx_hi_dev) {
  if (!h) return NLB200_ERR_INVALID;
  if (h->build_pending && h->last_stream) CK(h, cudaStreamSynchronize(h->last_stream));
  drop_graph(h);  // the arguments are captured with the binning kernels
  if (!state_dev) {  // undo
    h->halo_pack_on = false;
    h->halo_pack = HaloPackArgs{};
    return NLB200_OK;
  }
  if (!h->halo_ctrl) return fail(h, NLB200_ERR_STATE, "nlb200_set_halo_pack needs nlb200_set_halo_sync first");
  if (axis < 0 || axis > 2 || capacity < 0 || !out_counts_dev)
    return fail(h, NLB200_ERR_INVALID, "nlb200_set_halo_pack: bad argument");
  HaloPackArgs hp{};
  hp.axis = axis;
  hp.cut_lo = cut_lo;
  hp.cut_hi = cut_hi;
  hp.out_q_lo = peer_q_lo;
  hp.out_gid_lo = peer_gid_lo;
  hp.out_q_hi = peer_q_hi;
  hp.out_gid_hi = peer_gid_hi;
  hp.capacity = capacity;
  hp.state = reinterpret_cast<unsigned long long*>(state_dev);
  hp.out_counts = reinterpret_cast<long long*>(out_counts_dev);
  hp.ctrl = reinterpret_cast<HaloCtrl*>(h->halo_ctrl);
  hp.peer_ready_lo = reinterpret_cast<unsigned long long*>(peer_ready_lo);
  hp.peer_ready_hi = reinterpret_cast<unsigned long long*>(peer_ready_hi);
  hp.send_idx_lo = send_idx_lo_dev;
  hp.send_idx_hi = send_idx_hi_dev;
  h->halo_pack = hp;
  h->halo_pack_on = true;
  return NLB200_OK;
}

int nlb200_halo_refresh(nlb200_handle h, const void* q_dev, void* stream) {
  if (!h || !q_dev) return NLB200_ERR_INVALID;
  if (!h->halo_pack_on || !h->halo_pack.send_idx_lo || !h->halo_pack.send_idx_hi)
    return fail(h, NLB200_ERR_STATE, "nlb200_halo_refresh needs nlb200_set_halo_pack with send-index buffers");
  if (!h->have_build) return fail(h, NLB200_ERR_STATE, "nlb200_halo_refresh before the first build");
  (void)cudaGetLastError();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t cap = h->halo_pack.capacity > 0 ? h->halo_pack.capacity : 1;
  const unsigned g = (unsigned)((cap + 255) / 256);
  if (h->dtype == NLB200_F64)
    halo_refresh_kernel<double><<<g, 256, 0, s>>>((const double*)q_dev, h->stride, h->halo_pack);
  else
    halo_refresh_kernel<float><<<g, 256, 0, s>>>((const float*)q_dev, h->stride, h->halo_pack);
  return cudaGetLastError() == cudaSuccess ? NLB200_OK : NLB200_ERR_CUDA;
}

int nlb200_set_halo_sync(nlb200_handle h, void* ctrl_dev, void* peer_free_lo, void* peer_free_hi) {
  if (!h) return NLB200_ERR_INVALID;
  if (h->build_pending && h->last_stream) CK(h, cudaStreamSynchronize(h->last_stream));
  h->halo_ctrl = ctrl_dev;
  h->halo_free_lo = ctrl_dev ? peer_free_lo : nullptr;
  h->halo_free_hi = ctrl_dev ? peer_free_hi : nullptr;
  if (!ctrl_dev) {
    h->halo_pack_on = false;
    h->halo_pack = HaloPackArgs{};
  }
  drop_graph(h);  // the pointers are arguments of the captured finalize_kernel
  return NLB200_OK;
}

int nlb200_halo_wait(void* ctrl_dev, int faces, void* stream) {
  if (!ctrl_dev) return NLB200_ERR_INVALID;
  (void)cudaGetLastError();
  halo_wait_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<HaloCtrl*>(ctrl_dev), faces);
  return cudaGetLastError() == cudaSuccess ? NLB200_OK : NLB200_ERR_CUDA;
}

int nlb200_halo_done(void* ctrl_dev, void* peer_free_lo, void* peer_free_hi, void* stream) {
  if (!ctrl_dev) return NLB200_ERR_INVALID;
  (void)cudaGetLastError();
  halo_done_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<HaloCtrl*>(ctrl_dev), reinterpret_cast<unsigned long long*>(peer_free_lo),
      reinterpret_cast<unsigned long long*>(peer_free_hi));
  return cudaGetLastError() == cudaSuccess ? NLB200_OK : NLB200_ERR_CUDA;
}

int64_t nlb200_select_slab_workspace(int64_t n) {
  const int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE + 1;
  const size_t o_pos = align_up(sizeof(int32_t) * (size_t)(n + 8), 256);
  const size_t o_state = align_up(o_pos + sizeof(int64_t) * (size_t)(n + 1), 256);
  return (int64_t)(o_state + sizeof(unsigned long long) * (size_t)(tiles + 2));
}

int nlb200_shift_axis(void* q_dev, int64_t count, int dtype, int stride, int axis, double delta, void* stream) {
  if (count < 0 || (stride != 3 && stride != 4) || axis < 0 || axis > 2 || (dtype != NLB200_F64 && dtype != NLB200_F32))
    return NLB200_ERR_INVALID;
  if (count == 0) return NLB200_OK;
  if (!q_dev) return NLB200_ERR_INVALID;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const unsigned g = (unsigned)((count + 255) / 256);
  if (dtype == NLB200_F64)
    shift_axis_kernel<double><<<g, 256, 0, s>>>((double*)q_dev, count, stride, axis, delta);
  else
    shift_axis_kernel<float><<<g, 256, 0, s>>>((float*)q_dev, count, stride, axis, (float)delta);
  return cudaGetLastError() == cudaSuccess ? NLB200_OK : NLB200_ERR_CUDA;
}

int nlb200_gather_records(const void* src_dev, const int32_t* idx_dev, int64_t count, int dtype, int stride,
                          void* dst_dev, void* stream) {
  if (count < 0 || (stride != 3 && stride != 4)) return NLB200_ERR_INVALID;
  if (count == 0) return NLB200_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const unsigned g = (unsigned)((count * stride + 255) / 256);
  if (dtype == NLB200_F64)
    gather_records_kernel<double><<<g, 256, 0, s>>>((const double*)src_dev, idx_dev, count, stride, (double*)dst_dev);
  else
    gather_records_kernel<float><<<g, 256, 0, s>>>((const float*)src_dev, idx_dev, count, stride, (float*)dst_dev);
  return cudaGetLastError() == cudaSuccess ? NLB200_OK : NLB200_ERR_CUDA;
}

}  // extern "C"
