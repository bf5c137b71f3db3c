// nlist_b200_shim.hpp — the reference's class interface on top of the C ABI (include/nlist_b200.h).
//
// The reference has no FFI; its boundary is the header-only template class the drivers instantiate
// (SURVEY.md §8b):
//     NeighListGPU<Vectype, Dtype> nl(SEARCH_LENGTH, L, L, L); nl.Initialize(N);
//     nl.MakeNeighList(q, N, false, tblock, smem_hei); nl.number_of_pairs(); nl.neigh_list(); nl.number_of_partners();
//                                                                                         (make_list.cu:122-142)
//     NeighList<Vec> nl(SL, L, L, L); nl.Initialize(N); nl.MakeNeighList(q, N);
//     nl.number_of_pairs(); nl.sorted_list(); nl.key_pointer(); nl.number_of_partners();  (make_list.cpp:143-163)
// A driver shaped like make_list.cu / make_list.cpp compiles against these classes unchanged apart from the include
// and the namespace (drivers/make_list_b200.cpp is one).  Everything here is host C++; CUDA kernels are reached only
// through libnlist_b200.so.
//
// Differences that are deliberate (nlist_b200.h "Conventions"): per-instance state (the reference allows one live
// NeighListGPU per process, neighlist_gpu.hpp:15-18,303); every capacity is checked and a failure prints the
// library's message and exits with status 1, the reference's own way of failing (make_list.cpp:73-76).
#ifndef NLIST_B200_SHIM_HPP_
#define NLIST_B200_SHIM_HPP_

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <type_traits>
#include <vector>

#include "nlist_b200.h"

namespace nlb200 {

inline void die(nlb200_handle h, int status, const char* what) {
  std::fprintf(stderr, "%s: %s (%s)\n", what, h ? nlb200_last_error(h) : "", nlb200_status_string(status));
  std::exit(1);
}
inline void cuda_or_die(cudaError_t e, const char* what) {
  if (e != cudaSuccess) {
    std::fprintf(stderr, "%s: %s\n", what, cudaGetErrorString(e));
    std::exit(1);
  }
}

// Paired host/device buffer with the member names the reference drivers use on cuda_ptr<T> (cuda_ptr.cuh:11-112):
// allocate, host2dev, dev2host, set_val, operator[] on the host copy, implicit conversion to the device pointer.
// It can also wrap a device buffer owned by the library (borrow), which is how the accessors below hand out results.
template <typename T>
class cuda_ptr {
 public:
  cuda_ptr() = default;
  cuda_ptr(const cuda_ptr&) = delete;
  cuda_ptr& operator=(const cuda_ptr&) = delete;
  ~cuda_ptr() { release(); }

  void allocate(std::size_t n) {
    release();
    n_ = n;
    owns_dev_ = true;
    cuda_or_die(cudaMalloc(reinterpret_cast<void**>(&dev_), sizeof(T) * (n ? n : 1)), "cudaMalloc");
    cuda_or_die(cudaMallocHost(reinterpret_cast<void**>(&host_), sizeof(T) * (n ? n : 1)), "cudaMallocHost");
  }
  // view of a library-owned device buffer; the host mirror is ours
  void borrow(const T* dev, std::size_t n) {
    if (!owns_dev_ && host_ && n <= n_) {
      dev_ = const_cast<T*>(dev);
      n_used_ = n;
      return;
    }
    release();
    n_ = n_used_ = n;
    dev_ = const_cast<T*>(dev);
    owns_dev_ = false;
    cuda_or_die(cudaMallocHost(reinterpret_cast<void**>(&host_), sizeof(T) * (n ? n : 1)), "cudaMallocHost");
  }
  void host2dev() { cuda_or_die(cudaMemcpy(dev_, host_, sizeof(T) * size(), cudaMemcpyHostToDevice), "host2dev"); }
  void dev2host() { cuda_or_die(cudaMemcpy(host_, dev_, sizeof(T) * size(), cudaMemcpyDeviceToHost), "dev2host"); }
  void set_val(const T v) {
    for (std::size_t i = 0; i < size(); i++) host_[i] = v;
    host2dev();
  }
  std::size_t size() const { return owns_dev_ ? n_ : n_used_; }
  T& operator[](std::size_t i) { return host_[i]; }
  const T& operator[](std::size_t i) const { return host_[i]; }
  operator T*() { return dev_; }
  T* dev() { return dev_; }
  T* host() { return host_; }

 private:
  void release() {
    if (owns_dev_ && dev_) cudaFree(dev_);
    if (host_) cudaFreeHost(host_);
    dev_ = host_ = nullptr;
    n_ = n_used_ = 0;
  }
  T* dev_ = nullptr;
  T* host_ = nullptr;
  std::size_t n_ = 0, n_used_ = 0;
  bool owns_dev_ = true;
};

template <typename Dtype>
constexpr int dtype_code() {
  static_assert(std::is_same<Dtype, double>::value || std::is_same<Dtype, float>::value, "Dtype: float or double");
  return std::is_same<Dtype, double>::value ? NLB200_F64 : NLB200_F32;
}

// neighlist_gpu.hpp:43-488.  Vec = double4 / float4 (make_list.cu:6-12): four Dtype per record.
template <typename Vec, typename Dtype>
class NeighListGPU {
 public:
  static constexpr int MAX_PARTNERS = 200;  // row capacity of the list[k*N + i] view (neighlist_gpu.hpp:70)

  NeighListGPU(const Dtype search_length, const Dtype Lx, const Dtype Ly, const Dtype Lz) {
    static_assert(sizeof(Vec) == 4 * sizeof(Dtype), "Vec must be {x, y, z, w} of Dtype");
    const int st = nlb200_create(search_length, Lx, Ly, Lz, dtype_code<Dtype>(), NLB200_FULL_ELL_TRANSPOSED, &h_);
    if (st) die(nullptr, st, "NeighListGPU: box must hold at least 3 cells of the search length per axis");
    check(nlb200_set_option(h_, NLB200_OPT_ELL_ROWS, MAX_PARTNERS), "set_option");
    cuda_or_die(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking), "cudaStreamCreate");
  }
  ~NeighListGPU() {
    nlb200_destroy(h_);
    if (stream_) cudaStreamDestroy(stream_);
  }
  NeighListGPU(const NeighListGPU&) = delete;  // neighlist_gpu.hpp:260-266
  NeighListGPU& operator=(const NeighListGPU&) = delete;

  void Initialize(const int32_t particle_number) { check(nlb200_initialize(h_, particle_number, 0), "Initialize"); }

  // tblock_size / smem_hei selected the reference kernels' launch shape (make_list.cu:103-110); the library picks
  // its own, so they are accepted and ignored.
  void MakeNeighList(cuda_ptr<Vec>& q, const int32_t particle_number, const bool sync = true,
                     const int32_t /*tblock_size*/ = 128, const int32_t /*smem_hei*/ = 7) {
    n_ = particle_number;
    q_last_ = q.dev();
    check(nlb200_build(h_, q_last_, particle_number, stream_), "MakeNeighList");
    if (sync) synchronize();
  }
  // the `sync` of the reference is a cudaDeviceSynchronize (neighlist_gpu.hpp:465).  A partner list or cell capacity
  // that turns out too small (the reference overflows silently) is grown here and the build repeated.
  void synchronize() {
    for (int attempt = 0; attempt < 4; attempt++) {
      const int st = nlb200_synchronize(h_);
      if (st == NLB200_OK) return;
      if (st == NLB200_ERR_CAPACITY) {
        check(nlb200_reserve(h_, nlb200_required_entries(h_)), "reserve");
      } else if (st == NLB200_ERR_CELL_CAPACITY) {
        nlb200_stats s;
        nlb200_get_stats(h_, &s);
        check(nlb200_reserve_cell_capacity(h_, s.max_in_cell), "reserve_cell_capacity");
      } else {
        die(h_, st, "MakeNeighList");
      }
      check(nlb200_build(h_, q_last_, n_, stream_), "MakeNeighList");
    }
    die(h_, NLB200_ERR_CAPACITY, "MakeNeighList");
  }

  int32_t number_of_pairs() {  // neighlist_gpu.hpp:484-487 (sum of the partner counts)
    synchronize();
    const int64_t p = nlb200_number_of_pairs(h_);
    if (p > std::numeric_limits<int32_t>::max()) die(h_, NLB200_ERR_INVALID, "number_of_pairs exceeds int32");
    return static_cast<int32_t>(p);
  }
  int64_t number_of_pairs64() {
    synchronize();
    return nlb200_number_of_pairs(h_);
  }
  cuda_ptr<int32_t>& neigh_list() {  // list[k*N + i], -1 padded (kernel_impl.cuh:30)
    list_.borrow(nlb200_ell_transposed(h_), static_cast<std::size_t>(MAX_PARTNERS) * n_);
    return list_;
  }
  cuda_ptr<int32_t>& number_of_partners() {
    np_.borrow(nlb200_number_of_partners(h_), n_);
    return np_;
  }
  nlb200_handle handle() { return h_; }

  // ---- the callers either side of the build: what a driver does with the list -------------------
  // f2, Verlet-list lifetime: SEARCH_LENGTH includes a margin (make_list.cpp:23) so that a list survives until some
  // particle has moved margin / 2; the reference's drivers rebuild 100 times instead (make_list.cpp:153-155).
  void TrackReference(cuda_ptr<Vec>& q) { check(nlb200_track_reference(h_, q.dev(), n_, stream_), "TrackReference"); }
  double MaxDisplacement(cuda_ptr<Vec>& q) {
    double d = 0.0;
    check(nlb200_max_displacement(h_, q.dev(), n_, stream_, &d), "MaxDisplacement");
    return d;
  }
  bool NeedsRebuild(cuda_ptr<Vec>& q, const Dtype margin) { return MaxDisplacement(q) > 0.5 * (double)margin; }
  // f4, a consumer of the FULL rows: Lennard-Jones forces f[3N] (and per-particle energies) with cutoff rc <= search
  // length — the momenta `p` the reference allocates and never uses (make_list.cpp:135-140) get their forces here.
  void LJForces(cuda_ptr<Vec>& q, const double rc, const double epsilon, const double sigma, cuda_ptr<double>& forces,
                cuda_ptr<double>* energy = nullptr) {
    synchronize();
    check(nlb200_lj_forces(h_, q.dev(), rc, epsilon, sigma, forces.dev(), energy ? energy->dev() : nullptr, stream_),
          "LJForces");
    cuda_or_die(cudaStreamSynchronize(stream_), "LJForces");
  }
  // f1, the physical reorder the reference stubbed out (SortPtclData, neighlist_cpu.hpp:176-180; CopyGather,
  // neighlist_gpu.hpp:144-151): dst[slot] = src[cell order of the last build], `width` elements of T per particle.
  template <typename T>
  void GatherSorted(cuda_ptr<T>& src, cuda_ptr<T>& dst, const int width) {
    static_assert(sizeof(T) == 4 || sizeof(T) == 8, "4- or 8-byte elements");
    synchronize();
    check(nlb200_gather_sorted(h_, src.dev(), (int)sizeof(T), width, dst.dev(), stream_), "GatherSorted");
    cuda_or_die(cudaStreamSynchronize(stream_), "GatherSorted");
  }

 private:
  void check(int st, const char* what) {
    if (st) die(h_, st, what);
  }
  nlb200_handle h_ = nullptr;
  cudaStream_t stream_ = nullptr;
  int32_t n_ = 0;
  const void* q_last_ = nullptr;
  cuda_ptr<int32_t> list_, np_;
};

// Periodic boundaries (minimum image) with the NeighListGPU interface — SURVEY.md §8f f3.  The reference wraps cell
// indices only (neighlist_cpu.hpp:61-66) and measures plain distances (:219-223); an MD caller needs the minimum image.
// Every particle within the search length of a face gets an image beyond the opposite face, axis by axis (edge and
// corner images are images of images); the images ride behind the particles as ghost records whose global id is the
// imaged particle's id, and one open-boundary build over the box extended by the search length gives rows whose
// partners are the minimum-image neighbours:
//     q_all = [ particles | x images (lo, hi) | y images | z images ] + SL,   n_owned = N
// Image buffers have a fixed capacity (absent slots are NaN records), so a build needs no host synchronisation.
// Preconditions: positions in [0, L), L > 2 * search_length per axis (strictly: L >= 2 SL (1 + 1e-9), checked).  Output: CSR (offsets + partners), FULL or HALF.
template <typename Vec, typename Dtype>
class NeighListPeriodicGPU {
 public:
  NeighListPeriodicGPU(const Dtype search_length, const Dtype Lx, const Dtype Ly, const Dtype Lz,
                       const bool half = false, const double slack = 1.5)
      : sl_(search_length), slack_(slack) {
    static_assert(sizeof(Vec) == 4 * sizeof(Dtype), "Vec must be {x, y, z, w} of Dtype");
    L_[0] = Lx; L_[1] = Ly; L_[2] = Lz;
    for (int a = 0; a < 3; a++)
      // strictly longer than two search lengths (same margin as the Python classes, periodic.py): at L == 2 SL a pair
      // at r == SL and its image at L - r == SL would both pass `!(r2 > SL2)` and row i would list j twice
      if (!(L_[a] >= 2.0 * sl_ * (1.0 + 1e-9)))
        die(nullptr, NLB200_ERR_INVALID, "NeighListPeriodicGPU: every box edge must exceed 2 search lengths");
    const int st = nlb200_create(sl_, Lx + 2 * sl_, Ly + 2 * sl_, Lz + 2 * sl_, dtype_code<Dtype>(),
                                 half ? NLB200_HALF_CSR : NLB200_FULL_CSR, &h_);
    if (st) die(nullptr, st, "NeighListPeriodicGPU");
    half_ = half;
    cuda_or_die(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking), "cudaStreamCreate");
  }
  ~NeighListPeriodicGPU() {
    nlb200_destroy(h_);
    if (stream_) cudaStreamDestroy(stream_);
    if (ws_) cudaFree(ws_);
    if (cnt_) cudaFree(cnt_);
  }
  NeighListPeriodicGPU(const NeighListPeriodicGPU&) = delete;
  NeighListPeriodicGPU& operator=(const NeighListPeriodicGPU&) = delete;

  void Initialize(const int32_t particle_number) {
    n_ = particle_number;
    int64_t cur = n_;
    for (int a = 0; a < 3; a++) {
      cap_[a] = ((int64_t)(cur * sl_ / L_[a] * slack_) + 256 + 31) / 32 * 32;
      cur += 2 * cap_[a];
    }
    n_total_ = cur;
    const double dens = n_ / (L_[0] * L_[1] * L_[2]);
    const int64_t entries =
        (int64_t)(n_ * dens * 4.18879 * sl_ * sl_ * sl_ * (half_ ? 0.5 : 1.0) * 1.3) + 16 * (int64_t)n_ + 1024;
    check(nlb200_initialize(h_, n_total_, entries), "Initialize");
    q_all_.allocate((std::size_t)n_total_);
    gid_.allocate((std::size_t)n_total_);
    for (int64_t i = 0; i < n_total_; i++) gid_[i] = i < n_ ? (int32_t)i : 0;
    gid_.host2dev();
    ws_bytes_ = 2 * nlb200_select_slab_workspace(n_total_) + 512;
    cuda_or_die(cudaMalloc(&ws_, (std::size_t)ws_bytes_), "cudaMalloc");
    cuda_or_die(cudaMalloc(reinterpret_cast<void**>(&cnt_), 6 * sizeof(int64_t)), "cudaMalloc");
  }

  void MakeNeighList(cuda_ptr<Vec>& q, const int32_t particle_number, const bool sync = true) {
    if (particle_number != n_) die(h_, NLB200_ERR_INVALID, "MakeNeighList: particle number differs from Initialize");
    enqueue(q.dev());
    q_last_ = q.dev();
    if (sync) synchronize();
  }
  void synchronize() {
    for (int attempt = 0; attempt < 4; attempt++) {
      const int st = nlb200_synchronize(h_);
      if (st == NLB200_OK) break;
      if (st == NLB200_ERR_CAPACITY) {
        check(nlb200_reserve(h_, nlb200_required_entries(h_)), "reserve");
      } else if (st == NLB200_ERR_CELL_CAPACITY) {
        nlb200_stats s;
        nlb200_get_stats(h_, &s);
        check(nlb200_reserve_cell_capacity(h_, s.max_in_cell), "reserve_cell_capacity");
      } else {
        die(h_, st, "MakeNeighList");
      }
      if (attempt == 3) die(h_, NLB200_ERR_CAPACITY, "MakeNeighList");
      enqueue(q_last_);
    }
    int64_t c[6];
    cuda_or_die(cudaMemcpy(c, cnt_, sizeof(c), cudaMemcpyDeviceToHost), "image counts");
    for (int a = 0; a < 3; a++)
      if (c[2 * a] > cap_[a] || c[2 * a + 1] > cap_[a])
        die(h_, NLB200_ERR_CAPACITY, "NeighListPeriodicGPU: more periodic images than the capacity, raise `slack`");
  }
  int64_t number_of_pairs64() {
    synchronize();
    return nlb200_number_of_pairs(h_);
  }
  cuda_ptr<int32_t>& number_of_partners() {
    np_.borrow(nlb200_number_of_partners(h_), n_);
    return np_;
  }
  cuda_ptr<int64_t>& offsets() {
    off_.borrow(nlb200_offsets(h_), (std::size_t)n_ + 1);
    return off_;
  }
  cuda_ptr<int32_t>& partners() {
    list_.borrow(nlb200_partners(h_), (std::size_t)nlb200_number_of_pairs(h_));
    return list_;
  }

 private:
  void check(int st, const char* what) {
    if (st) die(h_, st, what);
  }
  void enqueue(const void* q_dev) {
    const int dt = dtype_code<Dtype>();
    Dtype* qa = reinterpret_cast<Dtype*>(q_all_.dev());
    cuda_or_die(cudaMemcpyAsync(qa, q_dev, sizeof(Vec) * (std::size_t)n_, cudaMemcpyDeviceToDevice, stream_), "copy");
    int64_t cur = n_;
    for (int a = 0; a < 3; a++) {
      Dtype* lo = qa + 4 * cur;
      Dtype* hi = qa + 4 * (cur + cap_[a]);
      check(nlb200_pack_slab2(qa, gid_.dev(), cur, dt, 4, a, sl_, L_[a] - sl_, lo, gid_.dev() + cur, hi,
                              gid_.dev() + cur + cap_[a], cap_[a], cnt_ + 2 * a, ws_, ws_bytes_, stream_),
            "nlb200_pack_slab2");
      check(nlb200_shift_axis(lo, cap_[a], dt, 4, a, L_[a], stream_), "nlb200_shift_axis");   // near the lower face: + L
      check(nlb200_shift_axis(hi, cap_[a], dt, 4, a, -L_[a], stream_), "nlb200_shift_axis");
      cur += 2 * cap_[a];
    }
    for (int a = 0; a < 3; a++) check(nlb200_shift_axis(qa, n_total_, dt, 4, a, sl_, stream_), "nlb200_shift_axis");
    check(nlb200_build_subset(h_, qa, n_total_, n_, gid_.dev(), stream_), "MakeNeighList");
  }
  nlb200_handle h_ = nullptr;
  cudaStream_t stream_ = nullptr;
  double sl_, slack_, L_[3];
  bool half_ = false;
  int32_t n_ = 0;
  int64_t n_total_ = 0, cap_[3] = {0, 0, 0}, ws_bytes_ = 0;
  void* ws_ = nullptr;
  int64_t* cnt_ = nullptr;
  const void* q_last_ = nullptr;
  cuda_ptr<Vec> q_all_;
  cuda_ptr<int32_t> gid_, np_, list_;
  cuda_ptr<int64_t> off_;
};

// neighlist_cpu.hpp:15-464 (and the AVX2 / AVX-512 classes, which share the interface): host positions in, half list
// in CSR out, key = smaller index.  Vec may be {x,y,z} (scalar build, make_list.cpp:30) or {x,y,z,w} (make_list.cpp:28).
template <typename Vec>
class NeighList {
 public:
  NeighList(const double search_length, const double Lx, const double Ly, const double Lz) {
    static_assert(sizeof(Vec) == 3 * sizeof(double) || sizeof(Vec) == 4 * sizeof(double), "Vec of 3 or 4 doubles");
    const int st = nlb200_create(search_length, Lx, Ly, Lz, NLB200_F64, NLB200_HALF_CSR, &h_);
    if (st) die(nullptr, st, "NeighList: box must hold at least 3 cells of the search length per axis");
    const int st2 = nlb200_set_option(h_, NLB200_OPT_POSITION_STRIDE, sizeof(Vec) / sizeof(double));
    if (st2) die(h_, st2, "set_option");
  }
  ~NeighList() { nlb200_destroy(h_); }
  NeighList(const NeighList&) = delete;  // neighlist_cpu.hpp:400-406
  NeighList& operator=(const NeighList&) = delete;

  void Initialize(const int32_t particle_number) {
    const int st = nlb200_initialize(h_, particle_number, 0);
    if (st) die(h_, st, "Initialize");
    np_.assign(particle_number > 0 ? particle_number : 1, 0);
    off64_.assign(static_cast<std::size_t>(particle_number) + 1, 0);
    kp_.assign(static_cast<std::size_t>(particle_number) + 1, 0);
  }

  void MakeNeighList(const Vec* q, const int32_t particle_number) {
    int64_t total = 0;
    // one build; counts and offsets come back with it, the list is fetched once its size is known
    int st = nlb200_build_host(h_, q, particle_number, np_.data(), off64_.data(), nullptr, 0, &total);
    if (st) die(h_, st, "MakeNeighList");
    if (static_cast<int64_t>(list_.size()) < total) list_.resize(static_cast<std::size_t>(total + total / 8 + 1024));
    st = nlb200_fetch_partners_host(h_, list_.data(), static_cast<int64_t>(list_.size()));
    if (st) die(h_, st, "MakeNeighList");
    if (total > std::numeric_limits<int32_t>::max()) die(h_, NLB200_ERR_INVALID, "list exceeds int32 key_pointer");
    pairs_ = static_cast<int32_t>(total);
    for (int32_t i = 0; i <= particle_number; i++) kp_[i] = static_cast<int32_t>(off64_[i]);
  }

  int32_t number_of_pairs() const { return pairs_; }
  int32_t* sorted_list() { return list_.data(); }
  const int32_t* sorted_list() const { return list_.data(); }
  int32_t* key_pointer() { return kp_.data(); }
  const int32_t* key_pointer() const { return kp_.data(); }
  int32_t* number_of_partners() { return np_.data(); }
  const int32_t* number_of_partners() const { return np_.data(); }
  nlb200_handle handle() { return h_; }

 private:
  nlb200_handle h_ = nullptr;
  int32_t pairs_ = 0;
  std::vector<int32_t> np_, kp_, list_;
  std::vector<int64_t> off64_;
};

}  // namespace nlb200

#endif  // NLIST_B200_SHIM_HPP_
