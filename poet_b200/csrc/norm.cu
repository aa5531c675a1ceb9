// Residual + LayerNorm (forward/backward), column sums (bias gradients), elementwise add.
// Replaces the `src = norm(src + dropout(src2))` pairs of the reference
// (models/deformable_transformer.py:196-197, 202-203, 270-271, 279-280, 286-287).  Train-mode dropout of the
// residual branch is counter-based (common.cuh PoetDropout): applied to r on the fly in the forward, and the
// backward writes the branch gradient dr = dz * mask / (1-p) next to dz; p = 0 / eval is the parity path.
// All of these are HBM-bound streaming kernels: one warp per row, 128-bit accesses, grid sized in
// multiples of the SM count with a grid-stride loop.
#include "common.cuh"

namespace {

template <int NV>   // C = NV * 128
__global__ void __launch_bounds__(256) add_ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ r,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         const float* __restrict__ pos, float* __restrict__ y,
                                                         float* __restrict__ y2, float* __restrict__ xhat,
                                                         float* __restrict__ rstd_out, int R, float eps, const PoetDropout drop) {
  poet_pdl_entry();
  constexpr int C = NV * 128;
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const bool dropping = drop.seed != nullptr && r != nullptr;      // y = LN(x + dropout(r))
  PoetDropKey key{0u, 0u};
  if (dropping) key = poet_drop_key(drop);
  float4 g[NV], bt[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) { g[i] = ldg4(gamma + i * 128 + lane * 4); bt[i] = ldg4(beta + i * 128 + lane * 4); }
  for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < R; row += gridDim.x * warps_per_block) {
    const int64_t base = (int64_t)row * C + lane * 4;
    float4 z[NV];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      z[i] = ld4(x + base + i * 128);
      if (r) {
        float4 t = ld4(r + base + i * 128);
        if (dropping) {
          const uint64_t e = (uint64_t)(base + i * 128);
          t.x *= poet_drop_mult(key, e, drop.threshold, drop.scale); t.y *= poet_drop_mult(key, e + 1, drop.threshold, drop.scale);
          t.z *= poet_drop_mult(key, e + 2, drop.threshold, drop.scale); t.w *= poet_drop_mult(key, e + 3, drop.threshold, drop.scale);
        }
        z[i].x += t.x; z[i].y += t.y; z[i].z += t.z; z[i].w += t.w;
      }
      sum += z[i].x + z[i].y + z[i].z + z[i].w;
    }
    const float mean = warp_sum(sum) * (1.f / C);
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      z[i].x -= mean; z[i].y -= mean; z[i].z -= mean; z[i].w -= mean;
      var += z[i].x * z[i].x + z[i].y * z[i].y + z[i].z * z[i].z + z[i].w * z[i].w;
    }
    const float rstd = rsqrtf(warp_sum(var) * (1.f / C) + eps);
    if (rstd_out && lane == 0) rstd_out[row] = rstd;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 h = make_float4(z[i].x * rstd, z[i].y * rstd, z[i].z * rstd, z[i].w * rstd);
      if (xhat) st4(xhat + base + i * 128, h);
      float4 o = make_float4(h.x * g[i].x + bt[i].x, h.y * g[i].y + bt[i].y, h.z * g[i].z + bt[i].z, h.w * g[i].w + bt[i].w);
      st4(y + base + i * 128, o);
      if (y2) {
        float4 pp = ld4(pos + base + i * 128);
        st4(y2 + base + i * 128, make_float4(o.x + pp.x, o.y + pp.y, o.z + pp.z, o.w + pp.w));
      }
    }
  }
}

// dz = rstd * (dxh - mean(dxh) - xhat * mean(dxh * xhat)),  dxh = dy * gamma
// dgamma += sum_rows dy * xhat, dbeta += sum_rows dy  (per-warp register partials -> smem -> atomics)
template <int NV, bool COLSUM, bool MANY>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ dy2,
                                                     const float* __restrict__ dy3, const float* __restrict__ dy4,
                                                     const float* __restrict__ xhat, const float* __restrict__ rstd,
                                                     const float* __restrict__ gamma, float* __restrict__ dz,
                                                     float* __restrict__ dgamma, float* __restrict__ dbeta, int R,
                                                     float* __restrict__ dr, float* __restrict__ dr_colsum,
                                                     const PoetDropout drop) {
  poet_pdl_entry();
  constexpr int C = NV * 128;
  __shared__ float s_dg[C], s_db[C], s_dc[COLSUM ? C : 1];
  const bool dropping = drop.seed != nullptr && dr != nullptr;     // dr = dz * mask / (1 - p): gradient of the dropped branch
  PoetDropKey key{0u, 0u};
  if (dropping) key = poet_drop_key(drop);
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int i = threadIdx.x; i < C; i += blockDim.x) { s_dg[i] = 0.f; s_db[i] = 0.f; if (COLSUM) s_dc[i] = 0.f; }
  __syncthreads();
  // pc: column sums of the residual branch's gradient (dr with dropout, else dz) = the bias gradient of the Linear that
  // produced r (reference: linear2 / output_proj / out_proj feeding norm2 / norm1), saved a pass over the gradient
  float4 g[NV], pg[NV], pb[NV], pc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    g[i] = ldg4(gamma + i * 128 + lane * 4);
    pg[i] = make_float4(0.f, 0.f, 0.f, 0.f); pb[i] = pg[i]; pc[i] = pg[i];
  }
  for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < R; row += gridDim.x * warps_per_block) {
    const int64_t base = (int64_t)row * C + lane * 4;
    float4 d[NV], h[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      d[i] = ld4(dy + base + i * 128);
      if (dy2) { float4 t = ld4(dy2 + base + i * 128); d[i].x += t.x; d[i].y += t.y; d[i].z += t.z; d[i].w += t.w; }
      // further consumers of y (a tensor read by several ops): their gradients are summed here, in registers, instead of
      // by accumulation kernels between the backward kernels of the decoder's dependent chain
      if (MANY && dy3) { float4 t = ld4(dy3 + base + i * 128); d[i].x += t.x; d[i].y += t.y; d[i].z += t.z; d[i].w += t.w; }
      if (MANY && dy4) { float4 t = ld4(dy4 + base + i * 128); d[i].x += t.x; d[i].y += t.y; d[i].z += t.z; d[i].w += t.w; }
      h[i] = ld4(xhat + base + i * 128);
      pg[i].x += d[i].x * h[i].x; pg[i].y += d[i].y * h[i].y; pg[i].z += d[i].z * h[i].z; pg[i].w += d[i].w * h[i].w;
      pb[i].x += d[i].x; pb[i].y += d[i].y; pb[i].z += d[i].z; pb[i].w += d[i].w;
      d[i].x *= g[i].x; d[i].y *= g[i].y; d[i].z *= g[i].z; d[i].w *= g[i].w;
      s1 += d[i].x + d[i].y + d[i].z + d[i].w;
      s2 += d[i].x * h[i].x + d[i].y * h[i].y + d[i].z * h[i].z + d[i].w * h[i].w;
    }
    const float m1 = warp_sum(s1) * (1.f / C), m2 = warp_sum(s2) * (1.f / C);
    const float rs = __ldg(rstd + row);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 o = make_float4(rs * (d[i].x - m1 - h[i].x * m2), rs * (d[i].y - m1 - h[i].y * m2),
                                   rs * (d[i].z - m1 - h[i].z * m2), rs * (d[i].w - m1 - h[i].w * m2));
      st4(dz + base + i * 128, o);
      float4 c = o;
      if (dropping) {
        const uint64_t e = (uint64_t)(base + i * 128);
        c = make_float4(o.x * poet_drop_mult(key, e, drop.threshold, drop.scale),
                        o.y * poet_drop_mult(key, e + 1, drop.threshold, drop.scale),
                        o.z * poet_drop_mult(key, e + 2, drop.threshold, drop.scale),
                        o.w * poet_drop_mult(key, e + 3, drop.threshold, drop.scale));
        st4(dr + base + i * 128, c);
      }
      if (COLSUM) { pc[i].x += c.x; pc[i].y += c.y; pc[i].z += c.z; pc[i].w += c.w; }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = i * 128 + lane * 4;
    atomicAdd(&s_dg[c + 0], pg[i].x); atomicAdd(&s_dg[c + 1], pg[i].y); atomicAdd(&s_dg[c + 2], pg[i].z); atomicAdd(&s_dg[c + 3], pg[i].w);
    atomicAdd(&s_db[c + 0], pb[i].x); atomicAdd(&s_db[c + 1], pb[i].y); atomicAdd(&s_db[c + 2], pb[i].z); atomicAdd(&s_db[c + 3], pb[i].w);
    if (COLSUM) {
      atomicAdd(&s_dc[c + 0], pc[i].x); atomicAdd(&s_dc[c + 1], pc[i].y); atomicAdd(&s_dc[c + 2], pc[i].z); atomicAdd(&s_dc[c + 3], pc[i].w);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(dgamma + i, s_dg[i]); atomicAdd(dbeta + i, s_db[i]);
    if (COLSUM) atomicAdd(dr_colsum + i, s_dc[i]);
  }
}

// out[n] (+)= sum_m X[m,n].  One warp covers 128 columns (float4 per lane), the 8 warps of a block take
// alternate rows of a 64-row chunk (8 independent 128-bit loads in flight per thread); per-block partials
// are combined in smem and leave as one atomic per column.  Tail columns (N % 4 != 0) use the scalar path.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, int64_t ldx, float* __restrict__ out,
                                                     int M, int N, int rows_per_block, int vec,
                                                     const uint8_t* __restrict__ row_mask) {
  poet_pdl_entry();
  __shared__ float part[8][128];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 128 + lane * 4;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (vec && c0 + 3 < N) {
#pragma unroll 4
    for (int r = r0 + w; r < r1; r += 8) {
      if (row_mask != nullptr && row_mask[r] != 0) continue;
      const float4 v = ldg4(X + (int64_t)r * ldx + c0);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  } else {
    for (int r = r0 + w; r < r1; r += 8) {
      if (row_mask != nullptr && row_mask[r] != 0) continue;
      const float* p = X + (int64_t)r * ldx + c0;
      if (c0 + 0 < N) acc.x += __ldg(p + 0);
      if (c0 + 1 < N) acc.y += __ldg(p + 1);
      if (c0 + 2 < N) acc.z += __ldg(p + 2);
      if (c0 + 3 < N) acc.w += __ldg(p + 3);
    }
  }
  part[w][lane * 4 + 0] = acc.x; part[w][lane * 4 + 1] = acc.y; part[w][lane * 4 + 2] = acc.z; part[w][lane * 4 + 3] = acc.w;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c < N) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s += part[i][threadIdx.x];
      atomicAdd(out + c, s);
    }
  }
}

__global__ void __launch_bounds__(256) add_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                                                  float4* __restrict__ out, int64_t n4) {
  poet_pdl_entry();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = a[i];
    if (b) { float4 w = b[i]; v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w; }
    out[i] = v;
  }
}

__global__ void __launch_bounds__(256) mask_rows_kernel(float* __restrict__ x, const uint8_t* __restrict__ mask, int R, int C) {
  poet_pdl_entry();
  const int64_t total = (int64_t)R * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    if (mask[i / C]) x[i] = 0.f;
}

// x <- dropout(x) in place with the pair scheme of the GEMM epilogue (same mask for the same (seed, site, index)):
// the fallback for an FFN hidden activation whose producing GEMM is not on the tensor-core epilogue path.
__global__ void __launch_bounds__(256) dropout_pairs_kernel(float4* __restrict__ x, int64_t n4, const PoetDropout drop) {
  poet_pdl_entry();
  const PoetDropKey key = poet_drop_key(drop);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = x[i];
    const uint32_t k0 = poet_drop_keep2(key, (uint64_t)(2 * i), drop.threshold16);
    const uint32_t k1 = poet_drop_keep2(key, (uint64_t)(2 * i + 1), drop.threshold16);
    v.x = (k0 & 1u) ? v.x * drop.scale16 : 0.f; v.y = (k0 & 2u) ? v.y * drop.scale16 : 0.f;
    v.z = (k1 & 1u) ? v.z * drop.scale16 : 0.f; v.w = (k1 & 2u) ? v.w * drop.scale16 : 0.f;
    x[i] = v;
  }
}

inline int row_grid(int R) {
  int blocks = poet_ceil_div(R, 8);                       // 8 warps (rows) per block
  int cap = POET_NUM_SMS * 8;
  return blocks < cap ? blocks : cap;
}

}  // namespace

extern "C" int poet_add_layernorm_fwd(const float* x, const float* r, const float* gamma, const float* beta,
                                      const float* pos, float* y, float* y2, float* xhat, float* rstd, int R, int C,
                                      float eps, const void* drop_seed, uint32_t drop_site, float drop_p,
                                      poet_stream_t stream) {
  POET_REQUIRE(x && gamma && beta && y, POET_ERR_NULL_POINTER);
  POET_REQUIRE(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || drop_seed != nullptr), POET_ERR_BAD_SHAPE);
  const PoetDropout drop = poet_make_dropout(drop_seed, drop_site, drop_p);
  POET_REQUIRE((y2 == nullptr) == (pos == nullptr), POET_ERR_NULL_POINTER);
  POET_REQUIRE(R > 0 && C % 128 == 0 && C <= 1024, POET_ERR_BAD_SHAPE);
  POET_REQUIRE(poet_aligned16(x) && poet_aligned16(y) && poet_aligned16(gamma) && poet_aligned16(beta) &&
               (!r || poet_aligned16(r)) && (!pos || poet_aligned16(pos)) && (!y2 || poet_aligned16(y2)) &&
               (!xhat || poet_aligned16(xhat)), POET_ERR_BAD_ALIGNMENT);
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = row_grid(R);
  switch (C / 128) {
    case 1: poet_launch(add_ln_fwd_kernel<1>, dim3(grid), dim3(256), 0, s, x, r, gamma, beta, pos, y, y2, xhat, rstd, R, eps, drop); break;
    case 2: poet_launch(add_ln_fwd_kernel<2>, dim3(grid), dim3(256), 0, s, x, r, gamma, beta, pos, y, y2, xhat, rstd, R, eps, drop); break;
    case 4: poet_launch(add_ln_fwd_kernel<4>, dim3(grid), dim3(256), 0, s, x, r, gamma, beta, pos, y, y2, xhat, rstd, R, eps, drop); break;
    case 8: poet_launch(add_ln_fwd_kernel<8>, dim3(grid), dim3(256), 0, s, x, r, gamma, beta, pos, y, y2, xhat, rstd, R, eps, drop); break;
    default: return POET_ERR_UNSUPPORTED;
  }
  return poet_launch_status();
}

extern "C" int poet_layernorm_bwd(const float* dy, const float* dy2, const float* dy3, const float* dy4, const float* xhat, const float* rstd,
                                  const float* gamma, float* dz, float* dgamma, float* dbeta, int R, int C,
                                  float* dr, float* dr_colsum, const void* drop_seed, uint32_t drop_site, float drop_p,
                                  poet_stream_t stream) {
  POET_REQUIRE(dy && xhat && rstd && gamma && dz && dgamma && dbeta, POET_ERR_NULL_POINTER);
  POET_REQUIRE(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || (drop_seed != nullptr && dr != nullptr)), POET_ERR_BAD_SHAPE);
  POET_REQUIRE(!dr || poet_aligned16(dr), POET_ERR_BAD_ALIGNMENT);
  const PoetDropout drop = poet_make_dropout(drop_seed, drop_site, drop_p);
  POET_REQUIRE(R > 0 && C % 128 == 0 && C <= 1024, POET_ERR_BAD_SHAPE);
  POET_REQUIRE(poet_aligned16(dy) && poet_aligned16(xhat) && poet_aligned16(dz) && poet_aligned16(gamma) &&
               (!dy2 || poet_aligned16(dy2)) && (!dy3 || poet_aligned16(dy3)) && (!dy4 || poet_aligned16(dy4)), POET_ERR_BAD_ALIGNMENT);
  cudaStream_t s = (cudaStream_t)stream;
  int grid = row_grid(R);
  if (grid > POET_NUM_SMS * 2) grid = POET_NUM_SMS * 2;   // fewer blocks -> fewer global atomics on dgamma/dbeta
  const bool many = dy3 != nullptr || dy4 != nullptr;       // the two-pointer instantiation keeps the encoder's 26 MB passes at their old speed
#define POET_LN_BWD(NVV)                                                                                                      \
  do {                                                                                                                          \
    if (many) {                                                                                                                 \
      if (dr_colsum) poet_launch(ln_bwd_kernel<NVV, true, true>, dim3(grid), dim3(256), 0, s, dy, dy2, dy3, dy4, xhat, rstd, gamma, dz, dgamma, dbeta, R, dr, dr_colsum, drop); \
      else poet_launch(ln_bwd_kernel<NVV, false, true>, dim3(grid), dim3(256), 0, s, dy, dy2, dy3, dy4, xhat, rstd, gamma, dz, dgamma, dbeta, R, dr, dr_colsum, drop); \
    } else {                                                                                                                    \
      if (dr_colsum) poet_launch(ln_bwd_kernel<NVV, true, false>, dim3(grid), dim3(256), 0, s, dy, dy2, dy3, dy4, xhat, rstd, gamma, dz, dgamma, dbeta, R, dr, dr_colsum, drop); \
      else poet_launch(ln_bwd_kernel<NVV, false, false>, dim3(grid), dim3(256), 0, s, dy, dy2, dy3, dy4, xhat, rstd, gamma, dz, dgamma, dbeta, R, dr, dr_colsum, drop); \
    }                                                                                                                           \
  } while (0)
  switch (C / 128) {
    case 1: POET_LN_BWD(1); break;
    case 2: POET_LN_BWD(2); break;
    case 4: POET_LN_BWD(4); break;
    case 8: POET_LN_BWD(8); break;
    default: return POET_ERR_UNSUPPORTED;
  }
#undef POET_LN_BWD
  return poet_launch_status();
}

extern "C" int poet_colsum(const float* X, int64_t ldx, float* out, int M, int N, int accumulate, poet_stream_t stream) {
  return poet_colsum_masked(X, ldx, nullptr, out, M, N, accumulate, stream);
}

extern "C" int poet_colsum_masked(const float* X, int64_t ldx, const uint8_t* row_mask, float* out, int M, int N,
                                  int accumulate, poet_stream_t stream) {
  POET_REQUIRE(X && out, POET_ERR_NULL_POINTER);
  POET_REQUIRE(M > 0 && N > 0 && ldx >= N, POET_ERR_BAD_SHAPE);
  cudaStream_t s = (cudaStream_t)stream;
  if (!accumulate) {
    cudaError_t e = cudaMemsetAsync(out, 0, (size_t)N * sizeof(float), s);
    if (e != cudaSuccess) return (int)e;
  }
  const int col_tiles = poet_ceil_div(N, 128);
  int rows_per_block = 64;
  while ((int64_t)poet_ceil_div(M, rows_per_block) * col_tiles > 8 * POET_NUM_SMS) rows_per_block *= 2;
  const int row_chunks = poet_ceil_div(M, rows_per_block);
  const int vec = poet_aligned16(X) && (ldx % 4 == 0);
  poet_launch(colsum_kernel, dim3(col_tiles, row_chunks), dim3(256), 0, s, X, ldx, out, M, N, rows_per_block, vec, row_mask);
  return poet_launch_status();
}

extern "C" int poet_add(const float* a, const float* b, float* out, int64_t n, poet_stream_t stream) {
  POET_REQUIRE(a && out, POET_ERR_NULL_POINTER);
  POET_REQUIRE(n > 0 && n % 4 == 0, POET_ERR_BAD_SHAPE);
  POET_REQUIRE(poet_aligned16(a) && poet_aligned16(out) && (!b || poet_aligned16(b)), POET_ERR_BAD_ALIGNMENT);
  int64_t n4 = n / 4;
  int grid = poet_ceil_div(n4, 256);
  if (grid > POET_NUM_SMS * 8) grid = POET_NUM_SMS * 8;
  poet_launch(add_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b),
                                                     reinterpret_cast<float4*>(out), n4);
  return poet_launch_status();
}

extern "C" int poet_dropout(float* x, int64_t n, const void* drop_seed, uint32_t drop_site, float drop_p,
                            poet_stream_t stream) {
  POET_REQUIRE(x && drop_seed, POET_ERR_NULL_POINTER);
  POET_REQUIRE(n > 0 && n % 4 == 0 && drop_p > 0.f && drop_p < 1.f, POET_ERR_BAD_SHAPE);
  POET_REQUIRE(poet_aligned16(x), POET_ERR_BAD_ALIGNMENT);
  const PoetDropout drop = poet_make_dropout(drop_seed, drop_site, drop_p);
  const int64_t n4 = n / 4;
  int grid = poet_ceil_div(n4, 256);
  if (grid > POET_NUM_SMS * 8) grid = POET_NUM_SMS * 8;
  poet_launch(dropout_pairs_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<float4*>(x), n4, drop);
  return poet_launch_status();
}

extern "C" float poet_dropout_scale(float drop_p, int pair_scheme) {
  const PoetDropout d = poet_make_dropout(reinterpret_cast<const void*>(1), 0, drop_p);
  return drop_p > 0.f ? (pair_scheme ? d.scale16 : d.scale) : 1.f;
}

extern "C" int poet_mask_rows(float* x, const uint8_t* mask, int R, int C, poet_stream_t stream) {
  POET_REQUIRE(x && mask, POET_ERR_NULL_POINTER);
  POET_REQUIRE(R > 0 && C > 0, POET_ERR_BAD_SHAPE);
  int grid = poet_ceil_div((int64_t)R * C, 256);
  if (grid > POET_NUM_SMS * 8) grid = POET_NUM_SMS * 8;
  poet_launch(mask_rows_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, x, mask, R, C);
  return poet_launch_status();
}
