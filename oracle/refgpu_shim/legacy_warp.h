// ORACLE — TEST INFRASTRUCTURE ONLY.
// Force-included when the reference's GPU driver (make_list.cu, kernel_impl.cuh, device_util.cuh) is compiled for
// sm_100: the pre-Volta warp intrinsics it uses (no *_sync) are rejected by ptxas for sm_70+.  They are mapped onto
// the *_sync forms over the currently active lanes — what the sm_35 hardware did implicitly.  Every toolkit header
// that itself mentions these names is included first, so that the macros only touch the reference's code.  The
// headers the reference forgets to include (<random>, <numeric>, thrust/gather.h ...: make_list.cu:32,152,
// neighlist_gpu.hpp:150) come along.  No reference source is edited.
#pragma once
#include <cmath>
#include <numeric>
#include <random>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cublas_v2.h>
#include <thrust/copy.h>
#include <thrust/device_new.h>
#include <thrust/device_ptr.h>
#include <thrust/fill.h>
#include <thrust/gather.h>
#include <thrust/reduce.h>
#include <thrust/scan.h>
#include <thrust/sequence.h>
#include <thrust/sort.h>
#define __ballot(p) __ballot_sync(__activemask(), (p))
#define __any(p) __any_sync(__activemask(), (p))
#define __all(p) __all_sync(__activemask(), (p))
#define __shfl(...) __shfl_sync(__activemask(), __VA_ARGS__)
#define __shfl_xor(...) __shfl_xor_sync(__activemask(), __VA_ARGS__)
