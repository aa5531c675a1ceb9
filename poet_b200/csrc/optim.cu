// Optimizer step on the flat gradient arena: global-norm clipping + AdamW, all parameters in one launch.
// Replaces `torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)` + `torch.optim.AdamW.step()` as called
// at reference engine.py:77-81 with the three lr groups of main.py:253-277.
//
//   poet_sumsq            sum of squares of the (all-reduced) gradient arena -> one double on the device
//   poet_adamw_clip_multi per parameter tensor (pointer table, like poet_split_bf16_multi): clip coefficient from
//                         that double (no host round trip), decoupled weight decay, Adam moments, parameter update,
//                         and - for matrices the GEMMs consume - the bf16 hi/lo planes of the NEW weights, so the
//                         next forward needs no separate split pass.
#include "common.cuh"
#include <cuda_bf16.h>

namespace {

__global__ void __launch_bounds__(256) sumsq_kernel(const float4* __restrict__ x, int64_t n4, double* __restrict__ out) {
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x + i);
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  acc = warp_sum(acc);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += (double)part[i];
    atomicAdd(out, s);
  }
}

struct AdamEntry {
  float4* param;          // fp32 parameter tensor (16-byte aligned; a tail of numel % 4 elements is updated scalar-wise)
  int64_t arena_off4;     // offset (in float4) of its gradient / moments in the flat arenas
  uint2* hi;              // bf16 planes of the updated weights (0: none)
  uint2* lo;
  int64_t n;              // numel
  int64_t first_chunk;    // chunks of 1024 float4: ceil(ceil(n / 4) / 1024) per tensor
  int32_t group;          // lr group index
  int32_t pad;
};

struct AdamArgs {
  const AdamEntry* table; int n_tensors;
  const float4* grad; float4* m; float4* v;
  const double* sumsq;
  const float* tensor_sumsq;   // per-tensor squared gradient norms (poet_grad_sumsq_multi) or nullptr
  double* sumsq_out;           // total written back by block 0 when the per-tensor norms are the source
  float* touched;              // per tensor: accumulated squared norms (> 0 <=> the tensor owns optimizer state) or nullptr
  float max_norm, lr[4], beta1, beta2, eps, weight_decay, bc1, bc2_sqrt;
};

// Per-tensor squared L2 norms of the gradient arena, one block per chunk of 1024 float4 (same chunking as the update).
// Their sum is the global norm of clip_grad_norm_; a tensor whose norm is exactly zero received no gradient in this
// step (torch leaves its .grad at None and AdamW skips it: no weight decay, no moment decay).
__global__ void __launch_bounds__(256) grad_sumsq_multi_kernel(const AdamEntry* __restrict__ table, int n_tensors,
                                                               const float4* __restrict__ grad, float* __restrict__ tsumsq) {
  const int64_t chunk = blockIdx.x;
  int lo_i = 0, hi_i = n_tensors - 1;
  while (lo_i < hi_i) {
    const int mid = (lo_i + hi_i + 1) >> 1;
    if (table[mid].first_chunk <= chunk) lo_i = mid; else hi_i = mid - 1;
  }
  const AdamEntry e = table[lo_i];
  const int64_t base = (chunk - e.first_chunk) * 1024;
  const int64_t n4 = e.n >> 2;
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int64_t i = base + threadIdx.x + j * 256;
    if (i < n4) {
      const float4 g = __ldg(grad + e.arena_off4 + i);
      acc += g.x * g.x + g.y * g.y + g.z * g.z + g.w * g.w;
    }
  }
  if ((e.n & 3) && base + 1024 > n4 && threadIdx.x == 0) {
    const float* gs = reinterpret_cast<const float*>(grad + e.arena_off4);
    for (int64_t t = n4 * 4; t < e.n; ++t) acc += gs[t] * gs[t];
  }
  acc = warp_sum(acc);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += part[i];
    if (s != 0.f) atomicAdd(tsumsq + lo_i, s);
  }
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, const AdamArgs& a, float coef, float decay,
                                            float step_size) {
  const float gk = g * coef;
  p *= decay;                                                          // decoupled weight decay (AdamW)
  m = a.beta1 * m + (1.f - a.beta1) * gk;                              // == lerp(m, g, 1 - beta1)
  v = a.beta2 * v + (1.f - a.beta2) * gk * gk;
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
  p -= step_size * (m / denom);
}

__global__ void __launch_bounds__(256) adamw_clip_kernel(const AdamArgs a) {
  const int64_t chunk = blockIdx.x;
  int lo_i = 0, hi_i = a.n_tensors - 1;
  while (lo_i < hi_i) {
    const int mid = (lo_i + hi_i + 1) >> 1;
    if (a.table[mid].first_chunk <= chunk) lo_i = mid; else hi_i = mid - 1;
  }
  const AdamEntry e = a.table[lo_i];
  double total = 0.0;
  if (a.tensor_sumsq != nullptr) {
    // global norm = sum of the per-tensor norms (every block folds the ~150 floats itself: no second reduction pass)
    __shared__ double red[8];
    double part = 0.0;
    for (int t = threadIdx.x; t < a.n_tensors; t += 256) part += (double)a.tensor_sumsq[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) total += red[i];
    if (chunk == 0 && threadIdx.x == 0 && a.sumsq_out != nullptr) *a.sumsq_out = total;
    const float mine = a.tensor_sumsq[lo_i];
    if (chunk == e.first_chunk && threadIdx.x == 0 && a.touched != nullptr) a.touched[lo_i] += mine;
    if (mine == 0.f) return;                                           // no gradient this step: torch.optim.AdamW skips it
  } else if (a.max_norm > 0.f) {
    total = *a.sumsq;
  }
  // torch.nn.utils.clip_grad_norm_: coef = min(1, max_norm / (||g|| + 1e-6)); max_norm <= 0 disables clipping
  float coef = 1.f;
  if (a.max_norm > 0.f) coef = fminf(1.f, a.max_norm / ((float)sqrt(total) + 1e-6f));
  const float lr = a.lr[e.group];
  const float decay = 1.f - lr * a.weight_decay, step_size = lr / a.bc1;
  const int64_t base = (chunk - e.first_chunk) * 1024;
  const int64_t n4 = e.n >> 2;
  if ((e.n & 3) && base + 1024 > n4 && threadIdx.x == 0) {             // scalar tail (biases like [66]), last chunk only
    float* ps = reinterpret_cast<float*>(e.param);
    const float* gs = reinterpret_cast<const float*>(a.grad + e.arena_off4);
    float* ms = reinterpret_cast<float*>(a.m + e.arena_off4);
    float* vs = reinterpret_cast<float*>(a.v + e.arena_off4);
    for (int64_t t = n4 * 4; t < e.n; ++t) adam_update(ps[t], gs[t], ms[t], vs[t], a, coef, decay, step_size);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int64_t i = base + threadIdx.x + j * 256;
    if (i >= n4) continue;
    float4 p = e.param[i];
    float4 g = a.grad[e.arena_off4 + i], m = a.m[e.arena_off4 + i], v = a.v[e.arena_off4 + i];
    adam_update(p.x, g.x, m.x, v.x, a, coef, decay, step_size);
    adam_update(p.y, g.y, m.y, v.y, a, coef, decay, step_size);
    adam_update(p.z, g.z, m.z, v.z, a, coef, decay, step_size);
    adam_update(p.w, g.w, m.w, v.w, a, coef, decay, step_size);
    e.param[i] = p; a.m[e.arena_off4 + i] = m; a.v[e.arena_off4 + i] = v;
    if (e.hi) {
      const float h0 = __bfloat162float(__float2bfloat16_rn(p.x)), h1 = __bfloat162float(__float2bfloat16_rn(p.y));
      const float h2 = __bfloat162float(__float2bfloat16_rn(p.z)), h3 = __bfloat162float(__float2bfloat16_rn(p.w));
      e.hi[i] = make_uint2(pack2(h0, h1), pack2(h2, h3));
      if (e.lo) e.lo[i] = make_uint2(pack2(p.x - h0, p.y - h1), pack2(p.z - h2, p.w - h3));
    }
  }
}

}  // namespace

extern "C" int poet_sumsq(const float* x, int64_t n, double* out, poet_stream_t stream) {
  POET_REQUIRE(x && out, POET_ERR_NULL_POINTER);
  POET_REQUIRE(n > 0 && n % 4 == 0, POET_ERR_BAD_SHAPE);
  POET_REQUIRE(poet_aligned16(x), POET_ERR_BAD_ALIGNMENT);
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(double), s);
  if (e != cudaSuccess) return (int)e;
  int grid = poet_ceil_div(n / 4, 256 * 8);
  if (grid > POET_NUM_SMS * 8) grid = POET_NUM_SMS * 8;
  sumsq_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const float4*>(x), n / 4, out);
  return poet_launch_status();
}

extern "C" int poet_grad_sumsq_multi(const void* table, int n_tensors, int64_t total_chunks, const float* grad,
                                     float* tensor_sumsq, poet_stream_t stream) {
  POET_REQUIRE(table && grad && tensor_sumsq, POET_ERR_NULL_POINTER);
  POET_REQUIRE(n_tensors > 0 && total_chunks > 0 && total_chunks < ((int64_t)1 << 31), POET_ERR_BAD_SHAPE);
  POET_REQUIRE(poet_aligned16(grad), POET_ERR_BAD_ALIGNMENT);
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(tensor_sumsq, 0, sizeof(float) * (size_t)n_tensors, s);
  if (e != cudaSuccess) return (int)e;
  grad_sumsq_multi_kernel<<<(unsigned)total_chunks, 256, 0, s>>>(reinterpret_cast<const AdamEntry*>(table), n_tensors,
                                                                reinterpret_cast<const float4*>(grad), tensor_sumsq);
  return poet_launch_status();
}

extern "C" int poet_adamw_clip_multi(const void* table, int n_tensors, int64_t total_chunks, const float* grad, float* m,
                                     float* v, double* sumsq, const float* tensor_sumsq, float* touched, float max_norm,
                                     const float* lr_host, int n_groups, float beta1, float beta2, float eps,
                                     float weight_decay, int64_t step, poet_stream_t stream) {
  POET_REQUIRE(table && grad && m && v && lr_host, POET_ERR_NULL_POINTER);
  POET_REQUIRE(max_norm <= 0.f || sumsq != nullptr || tensor_sumsq != nullptr, POET_ERR_NULL_POINTER);
  POET_REQUIRE(n_tensors > 0 && total_chunks > 0 && total_chunks < ((int64_t)1 << 31) && n_groups >= 1 && n_groups <= 4 &&
               step >= 1, POET_ERR_BAD_SHAPE);
  POET_REQUIRE(poet_aligned16(grad) && poet_aligned16(m) && poet_aligned16(v), POET_ERR_BAD_ALIGNMENT);
  static_assert(sizeof(AdamEntry) == 56, "table layout is part of the ABI (7 x 8 bytes)");
  AdamArgs a;
  a.table = reinterpret_cast<const AdamEntry*>(table); a.n_tensors = n_tensors;
  a.grad = reinterpret_cast<const float4*>(grad); a.m = reinterpret_cast<float4*>(m); a.v = reinterpret_cast<float4*>(v);
  a.sumsq = sumsq; a.sumsq_out = sumsq; a.tensor_sumsq = tensor_sumsq; a.touched = touched; a.max_norm = max_norm;
  for (int i = 0; i < 4; ++i) a.lr[i] = lr_host[i < n_groups ? i : 0];
  a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
  a.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  a.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  adamw_clip_kernel<<<(unsigned)total_chunks, 256, 0, (cudaStream_t)stream>>>(a);
  return poet_launch_status();
}
