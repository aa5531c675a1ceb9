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

// ---- counter-based dropout ---------------------------------------------------------------------------------
// Train-mode nn.Dropout of the reference layers (models/deformable_transformer.py:178-286, default p = 0.1,
// main.py:94) without a mask tensor: element `idx` of dropout site `site` is kept iff hash(seed, site, idx) >= p*2^32,
// and the kept value is scaled by 1/(1-p).  `seed` is read from DEVICE memory when the kernel runs, so a step replayed
// from a CUDA graph draws a fresh mask every replay, and the backward kernels regenerate exactly the forward's mask
// from the same (seed, site, idx).  PyTorch's Philox stream cannot be reproduced (trap 2): parity with
// the reference is defined in eval() / dropout 0; train-mode checks are statistical.
struct PoetDropout {
  const unsigned long long* seed;   // device pointer, nullptr = dropout off
  uint32_t site;                    // distinct per dropout site of the model
  uint32_t threshold;               // p * 2^32
  float scale;                      // 1 / (1 - p)
  uint32_t threshold16;             // pair scheme (two elements per hash): p quantised to 1/65536 ...
  float scale16;                    // ... and the matching 1 / (1 - threshold16 / 65536)
};

static inline PoetDropout poet_make_dropout(const void* seed, uint32_t site, float p) {
  PoetDropout d;
  d.seed = (p > 0.f) ? reinterpret_cast<const unsigned long long*>(seed) : nullptr;
  d.site = site;
  const double t = (double)p * 4294967296.0;
  d.threshold = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
  d.scale = p < 1.f ? 1.f / (1.f - p) : 0.f;
  d.threshold16 = d.threshold >> 16;
  d.scale16 = 1.f / (1.f - (float)d.threshold16 * (1.f / 65536.f));
  return d;
}

struct PoetDropKey { uint32_t k0, k1; };
__device__ __forceinline__ PoetDropKey poet_drop_key(const PoetDropout& d) {
  const unsigned long long s = __ldg(d.seed);
  PoetDropKey k;
  k.k0 = (uint32_t)s ^ (d.site * 0x9E3779B9u);
  k.k1 = (uint32_t)(s >> 32) + d.site * 0x85EBCA6Bu;
  return k;
}
// murmur3-style mix of the 64-bit element index under the 64-bit key (two multiply rounds + finaliser)
__device__ __forceinline__ uint32_t poet_drop_hash(PoetDropKey k, uint64_t idx) {
  uint32_t h = (uint32_t)idx * 0xCC9E2D51u;
  h = (h << 15) | (h >> 17);
  h = (h * 0x1B873593u) ^ k.k0;
  h = ((h << 13) | (h >> 19)) * 5u + 0xE6546B64u;
  h ^= k.k1 + (uint32_t)(idx >> 32) * 0x9E3779B1u;
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
__device__ __forceinline__ float poet_drop_mult(PoetDropKey k, uint64_t idx, uint32_t threshold, float scale) {
  return poet_drop_hash(k, idx) >= threshold ? scale : 0.f;
}
// Pair scheme for the GEMM epilogues (the FFN hidden activation: one hash per two elements): elements 2j and 2j+1
// of the flattened [rows, N] matrix use the low / high 16 bits of hash(j).  Bit 0 / bit 1 of the result = keep.
__device__ __forceinline__ uint32_t poet_drop_keep2(PoetDropKey k, uint64_t pair_idx, uint32_t threshold16) {
  const uint32_t h = poet_drop_hash(k, pair_idx);
  return ((h & 0xffffu) >= threshold16 ? 1u : 0u) | ((h >> 16) >= threshold16 ? 2u : 0u);
}

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
