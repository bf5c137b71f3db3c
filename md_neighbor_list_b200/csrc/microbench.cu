// microbench.cu — issue-rate probes that decide the shape of the pair-search inner loop on B200 (sm_100a):
// scalar FFMA vs packed FFMA2 (fma.rn.f32x2), DFMA, FSETP, shared-memory broadcast LDS.128.  Not part of the
// library; results are recorded in profiles/ and DESIGN.md.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int ITERS = 4096;

__global__ void k_ffma(float* out, float a, float b) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = fmaf(x[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma2(float* out, float a, float b) {
  unsigned long long x[8], av, bv;
  asm volatile("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
  asm volatile("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
#pragma unroll
  for (int i = 0; i < 8; i++) {
    float v = threadIdx.x * 1e-3f + i;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(x[i]) : "f"(v));
  }
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(av), "l"(bv));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    float lo, hi;
    asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i]));
    s += lo + hi;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dfma(float* out, double a, double b) {
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
}

// the pre-filter inner loop as in search_kernel: LDS.128 broadcast + 3 FFMA + FSETP + predicated IADD
__global__ void k_loop_scalar(float* out, int nj) {
  extern __shared__ float4 sj[];
  for (int k = threadIdx.x; k < nj; k += blockDim.x) sj[k] = make_float4(k * 1e-3f, k * 2e-3f, k * 3e-3f, -k * 1e-3f);
  __syncthreads();
  const float xi = threadIdx.x * 1e-2f, yi = xi + 1, zi = xi + 2, alo = 0.5f;
  int cnt = 0;
  for (int rep = 0; rep < 16; rep++) {
#pragma unroll 8
    for (int k = 0; k < nj; k++) {
      const float4 j = sj[k];
      const float t = fmaf(xi, j.x, fmaf(yi, j.y, fmaf(zi, j.z, j.w)));
      cnt += (t >= alo) ? 1 : 0;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)cnt;
}

// same, two j per iteration with packed FFMA2; smem layout {x0,x1,y0,y1},{z0,z1,w0,w1}
__global__ void k_loop_packed(float* out, int nj) {
  extern __shared__ float4 sj[];
  for (int k = threadIdx.x; k < nj; k += blockDim.x) sj[k] = make_float4(k * 1e-3f, k * 2e-3f, k * 3e-3f, -k * 1e-3f);
  __syncthreads();
  const float xi = threadIdx.x * 1e-2f, yi = xi + 1, zi = xi + 2, alo = 0.5f;
  unsigned long long xi2, yi2, zi2;
  asm volatile("mov.b64 %0, {%1, %1};" : "=l"(xi2) : "f"(xi));
  asm volatile("mov.b64 %0, {%1, %1};" : "=l"(yi2) : "f"(yi));
  asm volatile("mov.b64 %0, {%1, %1};" : "=l"(zi2) : "f"(zi));
  int cnt = 0;
  const ulonglong2* s2 = reinterpret_cast<const ulonglong2*>(sj);
  for (int rep = 0; rep < 16; rep++) {
#pragma unroll 4
    for (int k = 0; k < nj; k += 2) {
      const ulonglong2 a = s2[k];      // {x0,x1},{y0,y1}
      const ulonglong2 b = s2[k + 1];  // {z0,z1},{w0,w1}
      unsigned long long t;
      asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(t) : "l"(zi2), "l"(b.x), "l"(b.y));
      asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(t) : "l"(yi2), "l"(a.y));
      asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(t) : "l"(xi2), "l"(a.x));
      float t0, t1;
      asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(t0), "=f"(t1) : "l"(t));
      cnt += (t0 >= alo) ? 1 : 0;
      cnt += (t1 >= alo) ? 1 : 0;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)cnt;
}

template <typename F>
float time_ms(F f, int reps = 5) {
  cudaEvent_t a, b;
  CHECK(cudaEventCreate(&a));
  CHECK(cudaEventCreate(&b));
  f();
  CHECK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CHECK(cudaEventRecord(a));
    f();
    CHECK(cudaEventRecord(b));
    CHECK(cudaEventSynchronize(b));
    float ms;
    CHECK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p;
  CHECK(cudaGetDeviceProperties(&p, 0));
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("device %s, %d SMs, max clock %.0f MHz\n", p.name, p.multiProcessorCount, clk / 1000.0);
  const int blocks = p.multiProcessorCount * 8, threads = 256;
  float* out;
  CHECK(cudaMalloc(&out, sizeof(float) * blocks * threads));
  const double lanes = (double)blocks * threads * ITERS * 8;
  float ms;
  ms = time_ms([&] { k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
  printf("FFMA   : %.3f ms  %.2f T lane-FMA/s  (%.1f lane-FMA/clk/SM @1.9GHz)\n", ms, lanes / ms * 1e-9,
         lanes / (ms * 1e-3) / p.multiProcessorCount / 1.9e9);
  ms = time_ms([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
  printf("FFMA2  : %.3f ms  %.2f T lane-FMA/s  (%.1f lane-FMA/clk/SM @1.9GHz, 2 per instr)\n", ms, 2 * lanes / ms * 1e-9,
         2 * lanes / (ms * 1e-3) / p.multiProcessorCount / 1.9e9);
  ms = time_ms([&] { k_dfma<<<blocks, threads>>>(out, 1.0001, 0.5); });
  printf("DFMA   : %.3f ms  %.2f T lane-FMA/s  (%.1f lane-FMA/clk/SM @1.9GHz)\n", ms, lanes / ms * 1e-9,
         lanes / (ms * 1e-3) / p.multiProcessorCount / 1.9e9);
  const int nj = 1024;
  for (int th : {64, 128, 256}) {
    const int bl = p.multiProcessorCount * (2048 / th);
    const double tests = (double)bl * th * nj * 16;
    ms = time_ms([&] { k_loop_scalar<<<bl, th, nj * sizeof(float4)>>>(out, nj); });
    printf("loop scalar (block %3d): %.3f ms  %.1f G tests/s  %.2f issue-clk/warp-iter @1.9GHz\n", th, ms,
           tests / ms * 1e-6, (ms * 1e-3 * 1.9e9) * p.multiProcessorCount * 4 / (tests / 32));
    ms = time_ms([&] { k_loop_packed<<<bl, th, nj * sizeof(float4)>>>(out, nj); });
    printf("loop packed (block %3d): %.3f ms  %.1f G tests/s  %.2f issue-clk/warp-iter @1.9GHz\n", th, ms,
           tests / ms * 1e-6, (ms * 1e-3 * 1.9e9) * p.multiProcessorCount * 4 / (tests / 32));
  }
  return 0;
}
