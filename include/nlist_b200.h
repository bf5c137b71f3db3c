/*
 * nlist_b200.h — C ABI of the B200-native Verlet neighbor-list builder (libnlist_b200.so).
 *
 * This is the drop-in boundary for the list-build path of kohnakagawa/md_neighbor_list.  The reference has no FFI:
 * its boundary is the header-only C++ class instantiated by the drivers.  Every entry point below
 * names the reference interface it replaces (file:line relative to the reference root); include/nlist_b200_shim.hpp
 * re-creates the reference's class/method names on top of this ABI so that a make_list.cu / make_list.cpp shaped
 * driver compiles against either implementation.
 *
 * Conventions
 *   - plain C types only; device pointers are raw `void*` / typed pointers into CUDA device memory;
 *     `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - every call returns an nlb200_status (0 = ok).  nlb200_last_error() gives the message of the last failure on
 *     that handle.  Capacity overflow is DETECTED and reported (the reference's fixed capacities overflow silently:
 *     neighlist_cpu.hpp:37,76-78; neighlist_gpu.hpp:70,74,102,111).
 *   - the handle owns all outputs; getters return borrowed device pointers valid until the next build / reserve /
 *     destroy on that handle (reference ownership: neighlist_gpu.hpp:94-123, accessors 468-487).
 *   - per-handle state, no globals: any number of handles may live in one process (the reference allows one,
 *     neighlist_gpu.hpp:15-18,303).
 *   - no CPU fallback exists: every entry point that computes needs a CUDA device and fails with
 *     NLB200_ERR_CUDA otherwise.
 *
 * Semantics preserved from the reference (SURVEY.md §8b "Semantics to preserve")
 *   - open boundary: cell indices are clamped, distances are plain Euclidean (neighlist_cpu.hpp:219-223,
 *     kernel_impl.cuh:25-29) — no minimum image.
 *   - accept a pair iff !(r2 > SL2), r2 = fma(dz,dz, fma(dy,dy, dx*dx)) evaluated in the input precision
 *     (the contraction nvcc/g++ apply to `drx*drx + dry*dry + drz*drz`, kernel_impl.cuh:28, neighlist_cpu.hpp:222),
 *     SL2 = SL*SL rounded once in the input precision (neighlist_gpu.hpp:254, neighlist_cpu.hpp:394).
 *   - partner ids index the caller's array; the caller's array is never permuted.
 *   - HALF: row i holds the partners j > i (neighlist_cpu.hpp:225-236).  FULL: every j != i (kernel_impl.cuh:29).
 */
#ifndef NLIST_B200_H_
#define NLIST_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NLB200_VERSION 100

typedef struct nlb200_context* nlb200_handle;

typedef enum {
  NLB200_OK = 0,
  NLB200_ERR_INVALID = 1,     /* bad argument / precondition (e.g. fewer than 3 cells on an axis) */
  NLB200_ERR_CUDA = 2,        /* CUDA runtime failure or no device */
  NLB200_ERR_CAPACITY = 3,    /* partner-list capacity too small; nlb200_required_entries() tells how much */
  NLB200_ERR_OUT_OF_BOX = 4,  /* a particle lies more than one cell outside [0,L] (or is NaN) */
  NLB200_ERR_ELL_ROWS = 5,    /* a row is longer than the ELL row capacity (reference: silent overflow) */
  NLB200_ERR_STATE = 6,       /* call order violated (e.g. build before initialize) */
  NLB200_ERR_CELL_CAPACITY = 7 /* a cell holds more particles than the pair-mask words cover (reference: NMAX_IN_MESH
                                  = 70 unchecked, neighlist_gpu.hpp:74,111); nlb200_reserve_cell_capacity() grows it */
} nlb200_status;

/* Position element type — the reference's `Dtype` toggle, make_list.cu:6-12. */
typedef enum { NLB200_F32 = 0, NLB200_F64 = 1 } nlb200_dtype;

/* Output structure.
 *   HALF_CSR            : NeighList / NeighListAVX2 / NeighListAVX512 (neighlist_cpu.hpp:361-377) — each pair once,
 *                         keyed by the smaller index.
 *   FULL_CSR            : the rows of NeighListGPU (kernel_impl.cuh:3-35) stored in CSR.
 *   FULL_ELL_TRANSPOSED : FULL_CSR plus the reference GPU layout list[k*N + i], padded with -1
 *                         (kernel_impl.cuh:30, neighlist_gpu.hpp:271-274, consumed at make_list.cu:178-182). */
typedef enum { NLB200_HALF_CSR = 0, NLB200_FULL_CSR = 1, NLB200_FULL_ELL_TRANSPOSED = 2 } nlb200_mode;

typedef enum {
  /* elements per position record: 4 = {x,y,z,w} (double4/float4, make_list.cu:8,11; AVX Vec make_list.cpp:28),
   * 3 = {x,y,z} (scalar CPU Vec, make_list.cpp:30).  Default 4. */
  NLB200_OPT_POSITION_STRIDE = 1,
  /* 0 (default): rows in deterministic stencil order (cell-major, ids ascending inside a cell) — the order the
   * reference kernels discover partners in.  1: rows sorted ascending by partner id (what the reference tests
   * do before comparing, make_list.cpp:120-128,211; make_list.cu:183-184). */
  NLB200_OPT_SORT_ROWS = 2,
  /* row capacity of the ELL-transposed view — the reference's MAX_PARTNERS (neighlist_gpu.hpp:70-71). Default 200. */
  NLB200_OPT_ELL_ROWS = 3,
  /* 1: skip the FP32 pre-filter and run the exact input-precision test on every candidate (validation only). */
  NLB200_OPT_EXACT_ONLY = 4,
  /* 1 (default): replay the build as a CUDA graph when (q, n, stream) repeat. 0: plain launches. */
  NLB200_OPT_USE_GRAPH = 5,
  /* search / emission pair, for tuning and ablation.  0 (default): run masks for FULL and HALF lists (emission with
   * the partner ids gathered from global memory below 2^20 particles, through a shared-memory window from there on),
   * row masks once a cell may hold more than 256 particles.  1: one CTA per cell, every test evaluated twice
   * (count, fill; also what NLB200_OPT_EXACT_ONLY runs).  2: pair masks.  4: pair masks, HALF rows filtered by id in
   * the emission.  5, 6: row masks.  7: pair masks with 64-bit mask indices.  8: run masks.  9 / 10: run masks with
   * the gathering / the window emission forced.  100 + p (pair masks), 200 + p (run masks): p work units per cell.
   * Every variant produces the same rows in the same order. */
  NLB200_OPT_KERNEL_VARIANT = 6,
  /* 1: record a CUDA event between the stages of every build on the build's stream (disables graph replay);
   * read the per-stage device times with nlb200_get_stage_times.  The reference's counterpart is
   * `make cuda_profile=yes` + nvprof (Makefile:21,29-31). */
  NLB200_OPT_PROFILE = 7,
  /* most particles one cell may hold — the reference's NMAX_IN_MESH (neighlist_gpu.hpp:74).  0 (default): estimated
   * from the mean occupancy at initialize (mean + 6 sigma).  Exceeding it is detected (NLB200_ERR_CELL_CAPACITY). */
  NLB200_OPT_MAX_IN_CELL = 8,
  /* 0 (default): the kernels of a build run in plain stream order (inside the replayed CUDA graph).  1: they are
   * chained by programmatic dependent launch — a kernel's CTAs are scheduled while its predecessor drains and wait
   * on the device for its results (no effect on the results; measured slower on the B200: 183 vs 172 us per build,
   * DESIGN.md §5, so it is off unless asked for).  The environment variable NLB200_PDL=0/1 sets the default at
   * nlb200_create. */
  NLB200_OPT_PDL = 9
} nlb200_option;

typedef struct {
  int64_t n;                   /* particles of the last build */
  int64_t number_of_pairs;     /* emitted list entries (HALF: pairs, FULL: 2 x pairs) */
  int64_t candidates_tested;   /* distance tests evaluated by one search pass */
  int64_t band_tests;          /* candidates that fell in the FP32 pre-filter's uncertainty band and were re-tested
                                  exactly in the input precision */
  int64_t required_entries;    /* entries the last build needed (valid also after NLB200_ERR_CAPACITY) */
  int64_t capacity_entries;    /* current partner-list capacity */
  int32_t mesh[3];             /* cells per axis (neighlist_cpu.hpp:384-387) */
  int32_t max_partners;        /* longest row */
  int32_t max_in_cell;         /* most populated cell (reference: nmax_in_mesh_, neighlist_gpu.hpp:169) */
  int32_t reserved;
} nlb200_stats;

/* ---- lifecycle ------------------------------------------------------------------------------------------------ */

/* Replaces the constructors NeighListGPU(search_length, Lx, Ly, Lz) (neighlist_gpu.hpp:236-255) and
 * NeighList*(search_length, Lx, Ly, Lz) (neighlist_cpu.hpp:380-395).  search_length already includes the margin
 * (make_list.cpp:23).  mesh_size = int(L / search_length) must be >= 3 on every axis. */
int nlb200_create(double search_length, double lx, double ly, double lz, int dtype, int mode, nlb200_handle* out);

/* Compile-time -D switches / edited constants of the reference (Makefile:64-119, neighlist_gpu.hpp:70-75) become
 * run-time options.  Must be called before nlb200_initialize. */
int nlb200_set_option(nlb200_handle h, int option, int64_t value);

/* Multi-GPU (no reference counterpart): restricts the handle's cell grid along `axis` to the cells
 * [first_cell, first_cell + n_cells) of the global grid (mesh = int(L / search_length) cells, nlb200_get_stats().mesh).
 * Particles are still assigned to cells on the GLOBAL grid — the rows of a slab rank stay bit-identical to the rows
 * a single-GPU build gives those particles — but the rank bins, sorts and searches only the cells of its slab plus
 * the ghost layer instead of a grid that is mostly empty.  Every particle given to the handle must fall inside the
 * window (else NLB200_ERR_OUT_OF_BOX).  n_cells: 1, 2 or >= 4 (or the whole axis).  Before nlb200_initialize. */
int nlb200_set_cell_window(nlb200_handle h, int axis, int32_t first_cell, int32_t n_cells);

/* Replaces Initialize(particle_number) (neighlist_gpu.hpp:268-287, neighlist_cpu.hpp:408-415): allocates for up to
 * max_particles.  max_entries is the partner-list capacity; 0 = estimate from the density max_particles/(Lx*Ly*Lz)
 * (the reference hand-edits MAX_PARTNERS per density, neighlist_gpu.hpp:70-71). */
int nlb200_initialize(nlb200_handle h, int64_t max_particles, int64_t max_entries);

/* Grow (never shrink) the partner-list capacity; invalidates borrowed pointers. */
int nlb200_reserve(nlb200_handle h, int64_t max_entries);

/* Grow (never shrink) the per-cell particle capacity (NLB200_OPT_MAX_IN_CELL) after NLB200_ERR_CELL_CAPACITY;
 * nlb200_get_stats().max_in_cell tells how much the last build needed. */
int nlb200_reserve_cell_capacity(nlb200_handle h, int64_t max_in_cell);

/* Replaces the destructors (neighlist_gpu.hpp:256-258, neighlist_cpu.hpp:396-398). */
int nlb200_destroy(nlb200_handle h);

/* ---- the hot path --------------------------------------------------------------------------------------------- */

/* Replaces NeighListGPU::MakeNeighList(q, particle_number, sync=false, ...) (neighlist_gpu.hpp:289-466) and
 * NeighList*::MakeNeighList(q, particle_number) (neighlist_cpu.hpp:417-435).
 * q_dev: device pointer to n position records in the handle's dtype and stride.  Asynchronous on `stream`; no host
 * synchronisation happens inside.  Device-side failures (capacity, out-of-box) surface at nlb200_synchronize. */
int nlb200_build(nlb200_handle h, const void* q_dev, int64_t n, void* stream);

/* Multi-GPU form (no reference counterpart): q_dev holds n_total = owned + ghost records, rows are
 * emitted only for the first n_owned; if global_ids_dev != NULL partner ids (and the HALF-mode j > i rule) use
 * global_ids_dev[local index] instead of the local index.  Ghost records (index >= n_owned) whose x is NaN are absent:
 * padding of fixed-capacity halo buffers (nlb200_pack_slab). */
int nlb200_build_subset(nlb200_handle h, const void* q_dev, int64_t n_total, int64_t n_owned,
                        const int32_t* global_ids_dev, void* stream);

/* For callers that capture nlb200_build / nlb200_build_subset into a CUDA graph of their own (on a capturing stream the
 * library enqueues its plain kernel chain instead of replaying its own graph) and replay it: tells the handle that the captured build was enqueued again on `stream`, so that
 * nlb200_synchronize waits for it and fetches its status. */
int nlb200_mark_enqueued(nlb200_handle h, void* stream);

/* Replaces the `sync` argument / cudaDeviceSynchronize of neighlist_gpu.hpp:465 and make_list.cu:128: waits for the
 * last build on its stream, fetches the device status word and statistics.  Returns the build's status. */
int nlb200_synchronize(nlb200_handle h);

/* Convenience for host callers (the shape of the CPU classes, make_list.cpp:143-163): H2D copy of q_host, build,
 * synchronize, grow-and-retry on NLB200_ERR_CAPACITY, D2H of the outputs the caller asks for (NULL = skip).
 * offsets_host receives n+1 int64 values; partners_host must hold partners_capacity entries. */
int nlb200_build_host(nlb200_handle h, const void* q_host, int64_t n, int32_t* number_of_partners_host,
                      int64_t* offsets_host, int32_t* partners_host, int64_t partners_capacity,
                      int64_t* number_of_pairs);

/* D2H copy of the partner list of the last synchronized build into partners_host (capacity entries). */
int nlb200_fetch_partners_host(nlb200_handle h, int32_t* partners_host, int64_t capacity);

/* ---- accessors (borrowed device pointers) --------------------------------------------------------------------- */

/* number_of_partners() — neighlist_gpu.hpp:476-482, neighlist_cpu.hpp:455-461.  int32[n]. */
const int32_t* nlb200_number_of_partners(nlb200_handle h);
/* key_pointer() — neighlist_cpu.hpp:447-453; 64-bit because 16M particles x ~150 partners exceeds INT32_MAX
 * (SURVEY.md §7 "Index width").  int64[n+1]. */
const int64_t* nlb200_offsets(nlb200_handle h);
/* 32-bit view of the offsets for reference-shaped callers; NULL (and NLB200_ERR_INVALID from
 * nlb200_synchronize-time check) when the total exceeds INT32_MAX.  int32[n+1]. */
const int32_t* nlb200_offsets32(nlb200_handle h);
/* sorted_list() — neighlist_cpu.hpp:439-445; rows of the FULL list for the GPU modes.  int32[number_of_pairs]. */
const int32_t* nlb200_partners(nlb200_handle h);
/* neigh_list() — neighlist_gpu.hpp:468-474: list[k*n + i], k < ell_rows, -1 padded.  Only in
 * NLB200_FULL_ELL_TRANSPOSED mode, else NULL. */
const int32_t* nlb200_ell_transposed(nlb200_handle h);
/* number_of_pairs() — neighlist_gpu.hpp:484-487 (thrust::reduce), neighlist_cpu.hpp:437.  Needs a prior
 * nlb200_synchronize; returns -1 otherwise. */
int64_t nlb200_number_of_pairs(nlb200_handle h);

/* Cell-binning results, exposed for parity tests and for downstream kernels that want cell-sorted access
 * (reference: mesh_index_ / ptcl_id_in_mesh_, neighlist_cpu.hpp:146-165,
 * neighlist_gpu.hpp:153-199).  cell_start: int32[M+1]; sorted_ids: int32[n], ids ascending inside a cell. */
const int32_t* nlb200_cell_start(nlb200_handle h);
const int32_t* nlb200_sorted_ids(nlb200_handle h);

int nlb200_get_stats(nlb200_handle h, nlb200_stats* out);
/* Per-stage device times of the last (synchronized) build when NLB200_OPT_PROFILE is on: fills ms_out/stage_ids_out
 * (up to `capacity` stages, in execution order) and returns the stage count, or -1. */
int nlb200_get_stage_times(nlb200_handle h, float* ms_out, int32_t* stage_ids_out, int capacity);
const char* nlb200_stage_name(int stage_id);
int64_t nlb200_required_entries(nlb200_handle h);
const char* nlb200_last_error(nlb200_handle h);
const char* nlb200_status_string(int status);
int nlb200_version(void);

/* ---- callers either side of the build ------------------------------------------------------------ */

/* f2, Verlet-list lifetime.  The search length includes a margin (make_list.cpp:23: 3.0 + 0.3) so that a list stays
 * valid while no particle has moved more than margin/2 since it was built; the reference's driver fakes this with 100
 * identical rebuilds (make_list.cpp:153-155).  nlb200_track_reference remembers the positions a list was built from
 * (device copy, stream-ordered); nlb200_max_displacement returns max_i |q_now[i] - q_ref[i]| (synchronises `stream`):
 * rebuild when it exceeds margin/2. */
int nlb200_track_reference(nlb200_handle h, const void* q_dev, int64_t n, void* stream);
int nlb200_max_displacement(nlb200_handle h, const void* q_dev, int64_t n, void* stream, double* max_disp_host);

/* f1, the physical reorder the reference stubbed out (SortPtclData, neighlist_cpu.hpp:176-180; CopyGather + SORT_FREQ,
 * neighlist_gpu.hpp:72,144-151): dst[slot] = src[sorted_ids[slot]] for a per-particle array of `width` elements of
 * elem_bytes (4 or 8) each, in the cell order of the last build — lets a downstream kernel read coalesced. */
int nlb200_gather_sorted(nlb200_handle h, const void* src_dev, int elem_bytes, int width, void* dst_dev, void* stream);

/* f4, a consumer of the list: Lennard-Jones forces f[n][3] (and, if energy_dev != NULL, per-particle energies whose
 * sum is the total) over the FULL rows of the last build, cutoff rc <= search length, plain Euclidean distance like the
 * list.  The reference allocates the momenta `p` and never uses them (make_list.cpp:135-140); this closes the loop:
 * list quality end to end, and the build-versus-use cost. */
int nlb200_lj_forces(nlb200_handle h, const void* q_dev, double rc, double epsilon, double sigma, double* forces_dev,
                     double* energy_dev, void* stream);

/* ---- adjacent utilities (device side of the drivers) ---------------------------------------------------------- */

/* Ghost selection for slab decomposition: writes the indices i < n with lo <= q[i][axis] < hi into
 * out_idx_dev (ascending, deterministic) and the count into *out_count_dev.  capacity = size of out_idx_dev. */
int nlb200_select_slab(const void* q_dev, int64_t n, int dtype, int stride, int axis, double lo, double hi,
                       int32_t* out_idx_dev, int64_t capacity, int64_t* out_count_dev, void* workspace_dev,
                       int64_t workspace_bytes, void* stream);

/* Fixed-capacity halo packing (multi-GPU build without host synchronisation): the records of q with
 * lo <= q[i][axis] < hi go to out_q_dev[0..count) (ascending i) with their global ids (gids_dev[i], or gid_base + i when
 * gids_dev is NULL) in out_gid_dev; the remaining slots up to `capacity` are filled with NaN records.  A ghost record
 * whose x is NaN is ABSENT for nlb200_build_subset: it is binned nowhere and appears in no row, so both sides of an
 * exchange can always move `capacity` records.  *out_count_dev receives the true count (> capacity = overflow).
 * Workspace as for nlb200_select_slab. */
int nlb200_pack_slab(const void* q_dev, const int32_t* gids_dev, int32_t gid_base, int64_t n, int dtype, int stride,
                     int axis, double lo, double hi, void* out_q_dev, int32_t* out_gid_dev, int64_t capacity,
                     int64_t* out_count_dev, void* workspace_dev, int64_t workspace_bytes, void* stream);

/* Both faces of a slab in one pass: records with q[i][axis] < cut_lo go to the *_lo buffers, records with
 * q[i][axis] >= cut_hi to the *_hi buffers (either pair may be NULL: end slab), each as in nlb200_pack_slab;
 * out_counts_dev[0..1] receive the two true counts.  Workspace: 2 x nlb200_select_slab_workspace(n). */
int nlb200_pack_slab2(const void* q_dev, const int32_t* gids_dev, int64_t n, int dtype, int stride, int axis,
                      double cut_lo, double cut_hi, void* out_q_lo_dev, int32_t* out_gid_lo_dev, void* out_q_hi_dev,
                      int32_t* out_gid_hi_dev, int64_t capacity, int64_t* out_counts_dev, void* workspace_dev,
                      int64_t workspace_bytes, void* stream);

/* The same selection in ONE kernel launch (what a slab rank runs before every build): positions from atomics, so the
 * ghosts of a face arrive in no particular order — nlb200_build_subset sorts a cell's particles by global id, its
 * rows do not depend on it.  state_dev: 32 bytes of device memory the caller zeroes ONCE (the kernel leaves them
 * zeroed); out_counts_dev[0..1] receive the two true counts (> capacity = overflow); unused slots hold NaN records. */
int nlb200_pack_faces(const void* q_dev, const int32_t* gids_dev, int64_t n, int dtype, int stride, int axis,
                      double cut_lo, double cut_hi, void* out_q_lo_dev, int32_t* out_gid_lo_dev, void* out_q_hi_dev,
                      int32_t* out_gid_hi_dev, int64_t capacity, int64_t* out_counts_dev, void* state_dev,
                      void* stream);

/* ---- halo exchange by peer stores: the packing kernel IS the transfer (no NCCL call on the step) ------------------
 * A slab rank keeps its assembly buffer [owned | ghosts from below | ghosts from above] (and the matching global ids,
 * and a 64-byte control block) in memory it allocates with nlb200_p2p_alloc; the 64-byte IPC handles are exchanged
 * once through any host channel (the Python driver uses torch.distributed, a C++ host MPI or a file) and opened with
 * nlb200_p2p_open.  Per step, on one stream:
 *     nlb200_pack_faces_p2p   writes this rank's face particles INTO THE NEIGHBOURS' ghost regions over NVLink, pads
 *                             the unused slots with NaN records, then raises the neighbours' `ready` flags;
 *     nlb200_halo_wait        the ghosts of both faces have arrived (nlb200_pack_faces_p2p also waits for them before it
 *                             ends: the call is only needed by a host that packs and builds on different streams);
 *     nlb200_build_subset     the build;
 *     nlb200_halo_done        tells the neighbours that their ghosts may be overwritten, advances the step counter.
 * Flags carry a device-side step number (control block: u64 step, ready[2], free_from[2], error, pad[2]), so the four
 * calls replay as a CUDA graph; every device-side wait is bounded (~1 s) and sets `error` instead of hanging.
 * peer_ready_lo / peer_free_lo point at ready[1] / free_from[1] of the LOWER neighbour's control block (this rank is
 * its upper face), peer_ready_hi / peer_free_hi at ready[0] / free_from[0] of the upper neighbour's; NULL = no such
 * neighbour (end slab).  state_dev: 64 bytes of device memory zeroed once by the caller. */
int nlb200_p2p_alloc(int64_t bytes, void** dev_ptr, void* ipc_handle_64);
int nlb200_p2p_open(const void* ipc_handle_64, void** peer_ptr);
int nlb200_p2p_close(void* peer_ptr);
int nlb200_p2p_free(void* dev_ptr);
int nlb200_pack_faces_p2p(const void* q_dev, const int32_t* gids_dev, int64_t n, int dtype, int stride, int axis,
                          double cut_lo, double cut_hi, void* peer_q_lo, int32_t* peer_gid_lo, void* peer_q_hi,
                          int32_t* peer_gid_hi, int64_t capacity, int64_t* out_counts_dev, void* state_dev,
                          void* ctrl_dev, void* peer_ready_lo, void* peer_ready_hi, void* stream);
int nlb200_halo_wait(void* ctrl_dev, int faces, void* stream);
/* Folds nlb200_halo_done into the last kernel of every build of `h` (one launch less per step); NULL ctrl_dev undoes
 * it.  nlb200_pack_faces_p2p already waits for this rank's own ghosts before it ends, so a step is two calls:
 * nlb200_pack_faces_p2p, nlb200_build_subset. */
int nlb200_set_halo_sync(nlb200_handle h, void* ctrl_dev, void* peer_free_lo, void* peer_free_hi);
int nlb200_halo_done(void* ctrl_dev, void* peer_free_lo, void* peer_free_hi, void* stream);
/* Folds nlb200_pack_faces_p2p into every nlb200_build_subset of `h` that is given global ids and ghost slots (after
 * nlb200_set_halo_sync; same arguments as nlb200_pack_faces_p2p, the records [0, n_owned) of the build are the ones
 * packed): the binning kernel that reads an owned record also sends it if it lies beyond a cut, the neighbours' flags
 * are raised by its last CTA, and the ghosts are binned by a second launch whose CTAs wait for this rank's own flags.
 * A step is then ONE call, nlb200_build_subset — no separate pass over the positions, no packing launch.
 * state_dev == NULL undoes it (so does nlb200_set_halo_sync(h, NULL, ...)). */
int nlb200_set_halo_pack(nlb200_handle h, int axis, double cut_lo, double cut_hi, void* peer_q_lo, int32_t* peer_gid_lo,
                         void* peer_q_hi, int32_t* peer_gid_hi, int64_t capacity, int64_t* out_counts_dev,
                         void* state_dev, void* peer_ready_lo, void* peer_ready_hi, int32_t* send_idx_lo_dev,
                         int32_t* send_idx_hi_dev);
/* Incremental halo refresh (the drivers' 100 identical rebuilds, make_list.cpp:153-155, stand for
 * the MD steps a list survives): between two builds the list stays valid while nlb200_max_displacement <= margin / 2,
 * but its consumer needs the CURRENT positions of the ghosts.  With send_idx_lo_dev / send_idx_hi_dev given to
 * nlb200_set_halo_pack (capacity int32 each, or NULL: no recording) every build records which owned record went into
 * which ghost slot; nlb200_halo_refresh re-sends exactly those records of q_dev into the same slots of the neighbours
 * (no selection, no compaction; global ids and the list stay as they are), raises the neighbours' flags and waits for
 * this rank's own ghosts.  It is one step of the flag protocol: every rank calls it, consumes its ghosts, then ends the
 * step with nlb200_halo_done(ctrl, peer_free_lo, peer_free_hi, stream). */
int nlb200_halo_refresh(nlb200_handle h, const void* q_dev, void* stream);

/* Bytes of workspace nlb200_select_slab / nlb200_pack_slab need for n particles. */
int64_t nlb200_select_slab_workspace(int64_t n);

/* f3 helper (periodic images): q[i][axis] += delta for i in [0, count) — an image is a copy of a
 * particle shifted by one box length; absent (NaN) records stay absent.  The reference wraps cell indices only
 * (neighlist_cpu.hpp:61-66) and measures plain distances (:219-223); minimum-image lists are built from the
 * open-boundary build over image ghosts (nlb200_pack_slab2 + this + nlb200_build_subset). */
int nlb200_shift_axis(void* q_dev, int64_t count, int dtype, int stride, int axis, double delta, void* stream);

/* Gathers position records: dst[k] = src[idx[k]] (stride elements each). */
int nlb200_gather_records(const void* src_dev, const int32_t* idx_dev, int64_t count, int dtype, int stride,
                          void* dst_dev, void* stream);

/* ---- workload generators (host; mirror the reference drivers, not part of the list build) ---------------------- */

/* make_list.cpp:51-77 / make_list.cu:42-66 `init`: jittered FCC lattice, `std::mt19937 mt(seed)` (the reference uses
 * 2), U[0,0.1) jitter drawn x,y,z.  sx/sy/sz <= 0 -> int(L/s) as in the reference.  Writes `stride` doubles per
 * particle (w = 0).  q == NULL returns the particle count. */
int64_t nlb200_workload_fcc(double density, double L, int sx, int sy, int sz, uint32_t seed, double* q, int stride,
                            int64_t capacity);
/* x,y,z ~ U[0,L) from std::mt19937_64(seed). */
int64_t nlb200_workload_uniform(int64_t n, double L, uint64_t seed, double* q, int stride);
/* 50 % uniform background + 50 % in `blobs` isotropic Gaussian blobs (sigma = L/40, centres
 * uniform, reflected into [0,L)), std::mt19937_64(seed). */
int64_t nlb200_workload_clustered(int64_t n, double L, int blobs, uint64_t seed, double* q, int stride);

#ifdef __cplusplus
}
#endif
#endif /* NLIST_B200_H_ */
