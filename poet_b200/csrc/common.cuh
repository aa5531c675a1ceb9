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

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------------
// A step is several hundred small dependent kernels; with plain stream order each one pays launch latency + CTA
// scheduling + its prologue AFTER its predecessor has drained.  Every kernel of this library therefore (1) signals
// at entry that its dependents may be scheduled and (2) waits for its prerequisites (completion + memory flush)
// before its first global-memory access; launches carry the programmatic-stream-serialization attribute, which
// stream capture turns into programmatic graph edges.  POET_PDL=0 launches with plain stream order (the two
// instructions are then no-ops).
__device__ __forceinline__ void poet_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void poet_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void poet_pdl_entry() { poet_pdl_launch_dependents(); poet_pdl_wait(); }

#include <cstdlib>
#include <utility>
static inline bool poet_pdl_enabled() {
  static const bool on = []() { const char* e = getenv("POET_PDL"); return e ? atoi(e) != 0 : true; }();
  return on;
}

template <typename... KArgs, typename... Args>
static inline void poet_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = poet_pdl_enabled() ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  (void)cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);   // status is read back by poet_launch_status()
}
