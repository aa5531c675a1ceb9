// Shared helpers for the poet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/poet_b200.h"

#define POET_NUM_SMS 148   // B200: 2 dies x 74 SMs; grids are sized in multiples of this

#define POET_REQUIRE(cond, code) do { if (!(cond)) return (code); } while (0)

static inline int poet_launch_status() {
  cudaError_t e = cudaPeekAtLastError();
  return e == cudaSuccess ? POET_OK : (int)e;
}

static inline bool poet_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static inline int poet_ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
