// Decoder self-attention core for Q <= 32 object queries.
// Replaces the softmax(q k^T / sqrt(D)) v inside nn.MultiheadAttention as called at reference
// models/deformable_transformer.py:277-278 (q = k = tgt + query_pos, v = tgt, no masks: dummy
// queries attend and are attended, SURVEY.md §7 "exact semantics").  The in/out projections are
// poet_gemm calls.  One warp (= one CTA) per (image, head); lane i owns query row i.  The kernels sit on the decoder's
// dependent chain, so they are written for latency: every lane fetches ITS row of q / k / v (and grad_out) with all
// loads in flight at once, the rows are exchanged through shared memory, and the Q-step loops read broadcast
// shared-memory words -- the first version walked the key rows in global memory one dependent L2 round trip after the
// other (17.7 us forward, 19.5 us backward per decoder layer under ncu, three to four times a query-row GEMM).
#include "common.cuh"

namespace {

constexpr int kWarps = 1;

// D floats of a row with 128-bit loads (head slices start at multiples of D >= 8 floats: 16-byte aligned)
template <int D>
__device__ __forceinline__ void load_vec(const float* __restrict__ p, float (&x)[D]) {
#pragma unroll
  for (int d = 0; d < D; d += 4) {
    const float4 t = ldg4(p + d);
    x[d] = t.x; x[d + 1] = t.y; x[d + 2] = t.z; x[d + 3] = t.w;
  }
}
template <int D>
__device__ __forceinline__ void store_vec(float* __restrict__ p, const float (&x)[D], float s) {
#pragma unroll
  for (int d = 0; d < D; d += 4) st4(p + d, make_float4(x[d] * s, x[d + 1] * s, x[d + 2] * s, x[d + 3] * s));
}

template <int D>
__device__ __forceinline__ void row_to_smem(float* dst, const float (&x)[D]) {
#pragma unroll
  for (int d = 0; d < D; d += 4) *reinterpret_cast<float4*>(dst + d) = make_float4(x[d], x[d + 1], x[d + 2], x[d + 3]);
}
template <int D>
__device__ __forceinline__ void row_from_smem(const float* src, float (&x)[D]) {
#pragma unroll
  for (int d = 0; d < D; d += 4) {
    const float4 t = *reinterpret_cast<const float4*>(src + d);
    x[d] = t.x; x[d + 1] = t.y; x[d + 2] = t.z; x[d + 3] = t.w;
  }
}

template <int D>
__global__ void __launch_bounds__(kWarps * 32) mha_fwd_kernel(const float* __restrict__ q, int64_t ldq,
                                                              const float* __restrict__ k, int64_t ldk,
                                                              const float* __restrict__ v, int64_t ldv,
                                                              float* __restrict__ out, float* __restrict__ probs,
                                                              int B, int Q, int M, float scale, const PoetDropout drop) {
  poet_pdl_entry();
  constexpr int LD = D + 4;                                   // row stride: the 8 lanes of a quarter-warp store to distinct banks
  __shared__ __align__(16) float s_k[32][LD], s_v[32][LD];
  const int warp = blockIdx.x, lane = threadIdx.x & 31;
  if (warp >= B * M) return;
  const int b = warp / M, m = warp % M;
  const bool row = lane < Q;
  // attention-probability dropout of nn.MultiheadAttention(dropout=p) (reference deformable_transformer.py:253,277):
  // out = (P * mask / (1-p)) V; `probs` keeps the un-dropped P, the backward regenerates the mask
  const bool dropping = drop.seed != nullptr;
  PoetDropKey key{0u, 0u};
  if (dropping) key = poet_drop_key(drop);
  const int64_t r = (int64_t)b * Q + (row ? lane : 0);
  float qi[D];
  {
    float kr[D], vr[D];                                       // this lane's rows: 3 D/4 independent 128-bit loads in flight
    load_vec<D>(q + r * ldq + m * D, qi);
    load_vec<D>(k + r * ldk + m * D, kr);
    load_vec<D>(v + r * ldv + m * D, vr);
    row_to_smem<D>(s_k[lane], kr);
    row_to_smem<D>(s_v[lane], vr);
  }
  __syncwarp();
#pragma unroll
  for (int d = 0; d < D; ++d) qi[d] *= scale;
  float sc[32];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    sc[j] = -INFINITY;
    if (j < Q) {
      float kj[D];
      row_from_smem<D>(s_k[j], kj);
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) s = fmaf(qi[d], kj[d], s);
      sc[j] = s;
      mx = fmaxf(mx, s);
    }
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j)
    if (j < Q) { sc[j] = expf(sc[j] - mx); sum += sc[j]; }
  const float inv = 1.f / sum;
  float o[D];
#pragma unroll
  for (int d = 0; d < D; ++d) o[d] = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j)
    if (j < Q) {
      float pj = sc[j] * inv;
      const int64_t pidx = (((int64_t)b * M + m) * Q + lane) * Q + j;
      if (row && probs) probs[pidx] = pj;
      if (dropping) pj *= poet_drop_mult(key, (uint64_t)pidx, drop.threshold, drop.scale);
      float vj[D];
      row_from_smem<D>(s_v[j], vj);
#pragma unroll
      for (int d = 0; d < D; ++d) o[d] = fmaf(pj, vj[d], o[d]);
    }
  if (row) {
    float* op = out + ((int64_t)b * Q + lane) * (M * D) + m * D;
#pragma unroll
    for (int d = 0; d < D; d += 4) st4(op + d, make_float4(o[d], o[d + 1], o[d + 2], o[d + 3]));
  }
}

template <int D>
__global__ void __launch_bounds__(kWarps * 32) mha_bwd_kernel(const float* __restrict__ q, int64_t ldq,
                                                              const float* __restrict__ k, int64_t ldk,
                                                              const float* __restrict__ v, int64_t ldv,
                                                              const float* __restrict__ probs, const float* __restrict__ go,
                                                              float* __restrict__ gq, int64_t ldgq, float* __restrict__ gk,
                                                              int64_t ldgk, float* __restrict__ gv, int64_t ldgv,
                                                              int B, int Q, int M, float scale, const PoetDropout drop) {
  poet_pdl_entry();
  const bool dropping = drop.seed != nullptr;
  PoetDropKey key{0u, 0u};
  if (dropping) key = poet_drop_key(drop);
  constexpr int LD = D + 4;
  __shared__ __align__(16) float s_q[32][LD], s_k[32][LD], s_v[32][LD], s_g[32][LD];
  __shared__ float s_p[32][33];
  __shared__ float s_ds[32][33];
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x;
  if (warp >= B * M) return;
  const int b = warp / M, m = warp % M;
  const bool row = lane < Q;
  const int64_t r = (int64_t)b * Q + (row ? lane : 0);
  float gi[D];
  // this lane's rows of grad_out / q / k / v and its row of the saved probabilities: every load of the kernel is issued here
  float pr[32];
  {
    float t0[D], t1[D], t2[D];
    load_vec<D>(go + r * (M * D) + m * D, gi);
    load_vec<D>(q + r * ldq + m * D, t0);
    load_vec<D>(k + r * ldk + m * D, t1);
    load_vec<D>(v + r * ldv + m * D, t2);
    const float* prow = probs + (((int64_t)b * M + m) * Q + (row ? lane : 0)) * Q;
#pragma unroll
    for (int j = 0; j < 32; ++j) pr[j] = (row && j < Q) ? __ldg(prow + j) : 0.f;
    row_to_smem<D>(s_g[lane], gi);
    row_to_smem<D>(s_q[lane], t0);
    row_to_smem<D>(s_k[lane], t1);
    row_to_smem<D>(s_v[lane], t2);
  }
  __syncwarp();
  // dp_ij = <go_i, v_j>;  ds_ij = p_ij (dp_ij - sum_j p_ij dp_ij)
  // with dropout: out = (P o Mk) V, Mk = mask / (1-p): dP = (dO V^T) o Mk, dV = (P o Mk)^T dO, softmax backward on P
  float dp[32], mk[32];
  float dsum = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    dp[j] = 0.f; mk[j] = 1.f;
    if (j < Q) {
      float vj[D];
      row_from_smem<D>(s_v[j], vj);
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) s = fmaf(gi[d], vj[d], s);
      const int64_t pidx = (((int64_t)b * M + m) * Q + (row ? lane : 0)) * Q + j;
      if (dropping) { mk[j] = poet_drop_mult(key, (uint64_t)pidx, drop.threshold, drop.scale); s *= mk[j]; }
      dp[j] = s;
      dsum = fmaf(pr[j], s, dsum);
    }
  }
  float dq[D];
#pragma unroll
  for (int d = 0; d < D; ++d) dq[d] = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j)
    if (j < Q) {
      const float ds = pr[j] * (dp[j] - dsum);
      s_p[lane][j] = pr[j] * mk[j];
      s_ds[lane][j] = ds;
      float kj[D];
      row_from_smem<D>(s_k[j], kj);
#pragma unroll
      for (int d = 0; d < D; ++d) dq[d] = fmaf(ds, kj[d], dq[d]);
    }
  if (row) store_vec<D>(gq + ((int64_t)b * Q + lane) * ldgq + m * D, dq, scale);
  __syncwarp();
  // lane j: dk_j = scale * sum_i ds_ij q_i ; dv_j = sum_i p_ij go_i
  float dk[D], dv[D];
#pragma unroll
  for (int d = 0; d < D; ++d) { dk[d] = 0.f; dv[d] = 0.f; }
  for (int i = 0; i < Q; ++i) {
    const float ds = row ? s_ds[i][lane] : 0.f, pp = row ? s_p[i][lane] : 0.f;
    float qi[D], gp[D];
    row_from_smem<D>(s_q[i], qi);
    row_from_smem<D>(s_g[i], gp);
#pragma unroll
    for (int d = 0; d < D; ++d) { dk[d] = fmaf(ds, qi[d], dk[d]); dv[d] = fmaf(pp, gp[d], dv[d]); }
  }
  if (row) {
    store_vec<D>(gk + ((int64_t)b * Q + lane) * ldgk + m * D, dk, scale);
    store_vec<D>(gv + ((int64_t)b * Q + lane) * ldgv + m * D, dv, 1.f);
  }
}

}  // namespace

extern "C" int poet_mha_smallq_fwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                                   float* out, float* probs, int B, int Q, int M, int D, float scale, const void* drop_seed,
                                   uint32_t drop_site, float drop_p, poet_stream_t stream) {
  POET_REQUIRE(q && k && v && out, POET_ERR_NULL_POINTER);
  POET_REQUIRE(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || drop_seed != nullptr), POET_ERR_BAD_SHAPE);
  const PoetDropout drop = poet_make_dropout(drop_seed, drop_site, drop_p);
  POET_REQUIRE(B > 0 && M > 0 && Q >= 1 && Q <= 32, POET_ERR_BAD_SHAPE);
  POET_REQUIRE(poet_aligned16(out) && poet_aligned16(q) && poet_aligned16(k) && poet_aligned16(v) &&
               ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0, POET_ERR_BAD_ALIGNMENT);
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = poet_ceil_div(B * M, kWarps);
  switch (D) {
    case 8: poet_launch(mha_fwd_kernel<8>, dim3(grid), dim3(kWarps * 32), 0, s, q, ldq, k, ldk, v, ldv, out, probs, B, Q, M, scale, drop); break;
    case 16: poet_launch(mha_fwd_kernel<16>, dim3(grid), dim3(kWarps * 32), 0, s, q, ldq, k, ldk, v, ldv, out, probs, B, Q, M, scale, drop); break;
    case 32: poet_launch(mha_fwd_kernel<32>, dim3(grid), dim3(kWarps * 32), 0, s, q, ldq, k, ldk, v, ldv, out, probs, B, Q, M, scale, drop); break;
    case 64: poet_launch(mha_fwd_kernel<64>, dim3(grid), dim3(kWarps * 32), 0, s, q, ldq, k, ldk, v, ldv, out, probs, B, Q, M, scale, drop); break;
    default: return POET_ERR_UNSUPPORTED;
  }
  return poet_launch_status();
}

extern "C" int poet_mha_smallq_bwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                                   const float* probs, const float* grad_out, float* gq, int64_t ldgq, float* gk,
                                   int64_t ldgk, float* gv, int64_t ldgv, int B, int Q, int M, int D, float scale,
                                   const void* drop_seed, uint32_t drop_site, float drop_p, poet_stream_t stream) {
  POET_REQUIRE(q && k && v && probs && grad_out && gq && gk && gv, POET_ERR_NULL_POINTER);
  POET_REQUIRE(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || drop_seed != nullptr), POET_ERR_BAD_SHAPE);
  const PoetDropout drop = poet_make_dropout(drop_seed, drop_site, drop_p);
  POET_REQUIRE(B > 0 && M > 0 && Q >= 1 && Q <= 32, POET_ERR_BAD_SHAPE);
  POET_REQUIRE(poet_aligned16(q) && poet_aligned16(k) && poet_aligned16(v) && poet_aligned16(grad_out) && poet_aligned16(gq) &&
               poet_aligned16(gk) && poet_aligned16(gv) && ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldgq % 4 == 0 &&
               ldgk % 4 == 0 && ldgv % 4 == 0, POET_ERR_BAD_ALIGNMENT);
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = poet_ceil_div(B * M, kWarps);
#define POET_MHA_BWD(DD) poet_launch(mha_bwd_kernel<DD>, dim3(grid), dim3(kWarps * 32), 0, s, q, ldq, k, ldk, v, ldv, probs, grad_out, gq, ldgq, gk, ldgk, gv, ldgv, B, Q, M, scale, drop)
  switch (D) {
    case 8: POET_MHA_BWD(8); break;
    case 16: POET_MHA_BWD(16); break;
    case 32: POET_MHA_BWD(32); break;
    case 64: POET_MHA_BWD(64); break;
    default: return POET_ERR_UNSUPPORTED;
  }
#undef POET_MHA_BWD
  return poet_launch_status();
}
