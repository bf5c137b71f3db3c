// ORACLE — TEST INFRASTRUCTURE ONLY.
// Stand-in for the CUDA-samples header the reference includes but does not vendor (cuda_ptr.cuh:9): only
// checkCudaErrors is used by the reference's GPU path.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define checkCudaErrors(call)                                                                              \
  do {                                                                                                     \
    cudaError_t e_ = (call);                                                                               \
    if (e_ != cudaSuccess) {                                                                               \
      std::fprintf(stderr, "CUDA error %s at %s:%d: %s\n", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      std::exit(EXIT_FAILURE);                                                                             \
    }                                                                                                      \
  } while (0)
