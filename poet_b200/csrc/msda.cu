// Multi-scale deformable attention: gather (forward) and scatter (backward) kernels.
//
// Replaces the third-party op behind `from deformable_attention import MSDeformAttn`
// (reference models/deformable_transformer.py:24; upstream ms_deform_attn_forward/backward).
// Sampling convention (pinned by oracle.msda_core_direct): pixel x = loc_x*W - 0.5,
// y = loc_y*H - 0.5, bilinear with zero padding, a sample contributes iff -1 < x < W, -1 < y < H.
//
// Work decomposition (general path, any S / Lq): one thread per (b, q, m, 4-channel group).
// The G = D/4 lanes that share a (b,q,m) sit next to each other in a warp, so every bilinear corner
// is one contiguous D*4-byte segment per group (64 B at D=16, 128 B at D=32) fetched with
// 128-bit loads, and the output row [B,Lq,M*D] is written as one fully coalesced float4 stream.
// mode 1 fuses the module's softmax over L*P and loc = ref + off/(W_l,H_l), so the [B,Lq,M,L,P,2]
// location and [B,Lq,M,L,P] attention tensors are never materialised (SURVEY.md §8d msda_block).
//
// Slab path (forward, D = 16, many queries per image: the encoder): one CTA per (image, head, query
// chunk).  The head's slice of the flattened multi-level feature map, value[b, :, m, :] = S rows of
// 64 B, is staged ONCE into shared memory by TMA (2-D boxes of 64 rows, one mbarrier) and all gathers
// of the chunk are served from there: the general path is bound by L1 wavefronts (8 distinct 64-byte
// segments per warp load), the slab path by the 128 B/clk shared-memory crossbar.  Eight lanes share a
// (q,m): lanes 0-3 take the left corner column, lanes 4-7 the right one, so a quarter-warp always reads
// 128 contiguous bytes (x0 and x0+1 of one row): bank-conflict free.  The two halves are summed with
// one shuffle at the end.
#include <cstdlib>
#include <cuda.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "tma_host.cuh"

namespace {

constexpr int kMaxLevels = 4;
constexpr int kMaxLP = 32;   // L*P register budget

struct Levels {
  int H[kMaxLevels];
  int W[kMaxLevels];
  int start[kMaxLevels];
  float inv_H[kMaxLevels];   // 1/H_l, 1/W_l: off/(W,H) as a multiply (<= 1 ulp from the reference's divide)
  float inv_W[kMaxLevels];
};

struct MsdaArgs {
  const float* value; const float* a; int64_t lda; const float* w; int64_t ldw; const float* ref;
  float* out;
  const float* grad_out; float* grad_value; float* grad_a; float* grad_w;
  int B, S, Lq, M, D, L, P;
  int red_levels;            // backward: levels [0, red_levels) scatter grad_value with global reductions; the rest is
                             // produced by msda_bwd_dense_kernel (L = all levels: the classic path)
  Levels lv;
};

template <int LP>
__device__ __forceinline__ void load_row(const float* __restrict__ p, float (&dst)[LP]) {
#pragma unroll
  for (int i = 0; i < LP; i += 4) {
    float4 v = ldg4(p + i);
    dst[i] = v.x; dst[i + 1] = v.y; dst[i + 2] = v.z; dst[i + 3] = v.w;
  }
}

template <int LP>
__device__ __forceinline__ void softmax_inplace(float (&x)[LP]) {
  float mx = x[0];
#pragma unroll
  for (int i = 1; i < LP; ++i) mx = fmaxf(mx, x[i]);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < LP; ++i) { x[i] = __expf(x[i] - mx); sum += x[i]; }   // ex2.approx: ~2 ulp, far inside the budget
  const float inv = __fdividef(1.f, sum);
#pragma unroll
  for (int i = 0; i < LP; ++i) x[i] *= inv;
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
template <int L, int P, int G, bool FUSED>
__global__ void __launch_bounds__(256) msda_fwd_kernel(const MsdaArgs p) {
  poet_pdl_entry();
  constexpr int LP = L * P, D = 4 * G;                       // G lanes per (b,q,m)
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)p.B * p.Lq * p.M * G;
  if (t >= total) return;
  const int c4 = (int)(t % G);
  const int64_t bqm = t / G;
  const int m = (int)(bqm % p.M);
  const int64_t bq = bqm / p.M;
  const int b = (int)(bq / p.Lq);

  float aw[LP];
  load_row<LP>(p.w + bq * p.ldw + m * LP, aw);
  if (FUSED) softmax_inplace<LP>(aw);

  const float* arow = p.a + bq * p.lda + m * LP * 2;
  const int vstride = p.M * D;
  const float* vbase = p.value + ((int64_t)b * p.S * p.M + m) * D + c4 * 4;

  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int l = 0; l < L; ++l) {
    const int H = p.lv.H[l], W = p.lv.W[l];
    const float* vl = vbase + (int64_t)p.lv.start[l] * vstride;
    float rx = 0.f, ry = 0.f;
    if (FUSED) { float2 r = __ldg(reinterpret_cast<const float2*>(p.ref + (bq * L + l) * 2)); rx = r.x; ry = r.y; }
    float xy[2 * P];
#pragma unroll
    for (int i = 0; i < 2 * P; i += 4) {
      float4 v = ldg4(arow + l * 2 * P + i);
      xy[i] = v.x; xy[i + 1] = v.y; xy[i + 2] = v.z; xy[i + 3] = v.w;
    }
#pragma unroll
    for (int s = 0; s < P; ++s) {
      float lx = xy[2 * s], ly = xy[2 * s + 1];
      if (FUSED) { lx = rx + lx * p.lv.inv_W[l]; ly = ry + ly * p.lv.inv_H[l]; }
      const float x = lx * (float)W - 0.5f, y = ly * (float)H - 0.5f;
      if (x > -1.f && y > -1.f && x < (float)W && y < (float)H) {
        const float xf = floorf(x), yf = floorf(y);
        const int x0 = (int)xf, y0 = (int)yf;
        const float fx = x - xf, fy = y - yf;
        const float a = aw[l * P + s];
        const float w00 = (1.f - fy) * (1.f - fx) * a, w01 = (1.f - fy) * fx * a;
        const float w10 = fy * (1.f - fx) * a, w11 = fy * fx * a;
        const bool xl = x0 >= 0, xh = x0 + 1 < W, yl = y0 >= 0, yh = y0 + 1 < H;
        const float* p00 = vl + (y0 * W + x0) * vstride;
        float4 v00 = make_float4(0.f, 0.f, 0.f, 0.f), v01 = v00, v10 = v00, v11 = v00;
        if (yl && xl) v00 = ldg4(p00);
        if (yl && xh) v01 = ldg4(p00 + vstride);
        if (yh && xl) v10 = ldg4(p00 + W * vstride);
        if (yh && xh) v11 = ldg4(p00 + (W + 1) * vstride);
        acc.x += w00 * v00.x + w01 * v01.x + w10 * v10.x + w11 * v11.x;
        acc.y += w00 * v00.y + w01 * v01.y + w10 * v10.y + w11 * v11.y;
        acc.z += w00 * v00.z + w01 * v01.z + w10 * v10.z + w11 * v11.z;
        acc.w += w00 * v00.w + w01 * v01.w + w10 * v10.w + w11 * v11.w;
      }
    }
  }
  st4(p.out + t * 4, acc);
}


// ------------------------------------------------------------------------------------------
// forward, value slab staged in shared memory by TMA
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}

constexpr int kSlabBoxRows = 64;                 // TMA box: 64 rows x 16 floats = 4 KB
constexpr int kSlabThreads = 256;
constexpr int kSlabWarps = kSlabThreads / 32;
constexpr int kRecCorner = 16 * 8 + 16;          // 16 points x {offset, weight} + 16 B pad: the 4 corners x 2 queries of a warp load hit distinct banks
constexpr int kRecBytesPerWarp = 2 * 4 * kRecCorner;    // 2 queries x 4 corners x 16 points x {offset, weight}

// Each warp walks pairs of queries of its CTA's chunk in two phases:
//   A  one lane per (query, sampling point): softmax over the 16 logits (shuffles inside the 16-lane half),
//      location -> the four bilinear corners as {byte offset into the slab, weight * attention} records,
//      out-of-range corners as weight 0 on a clamped offset (phase B is branch-free), records -> per-warp smem;
//   B  16 lanes per query = 2 rows x 2 columns x 4 channel groups: per point one 8-byte record (broadcast) and
//      one 128-bit slab load; a quarter-warp (column pair x 4 groups) reads 128 contiguous bytes: conflict-free.
// The per-point arithmetic is done once per (q,m) instead of once per lane, which is what bounds the general
// kernel (issue-bound), and the corner reads come from shared memory instead of 8 L1 wavefronts per warp load.
template <int L, int P, bool FUSED>
__global__ void __launch_bounds__(kSlabThreads, 2)
msda_fwd_slab_kernel(const MsdaArgs p, const __grid_constant__ CUtensorMap tm_v, int nsplit, int q_per_cta, int n_boxes) {
  poet_pdl_entry();
  constexpr int LP = L * P, D = 16;
  static_assert(LP == 16, "one lane per sampling point in a 16-lane half-warp");
  extern __shared__ uint8_t slab_raw[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t slab = (smem_addr(slab_raw) + 127u) & ~127u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rec = slab + (uint32_t)(n_boxes * kSlabBoxRows * D * 4) + (uint32_t)warp * kRecBytesPerWarp;
  const int chunk = blockIdx.x % nsplit, bm = blockIdx.x / nsplit;
  const int m = bm % p.M, b = bm / p.M;
  const uint32_t bar_a = smem_addr(&bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a),
                 "r"((uint32_t)(n_boxes * kSlabBoxRows * D * 4)) : "memory");
    for (int i = 0; i < n_boxes; ++i)      // rows past the end of the tensor are zero-filled; rows of the next image are never read
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(slab + (uint32_t)(i * kSlabBoxRows * D * 4)), "l"(&tm_v), "r"(m * D), "r"(b * p.S + i * kSlabBoxRows),
                     "r"(bar_a) : "memory");
  }
  const int q_beg = chunk * q_per_cta, q_end = min(p.Lq, q_beg + q_per_cta);
  const int n_pairs = (q_end - q_beg + 1) >> 1;
  // phase A roles
  const int qa = lane >> 4, pt = lane & 15, lvl = pt / P;
  const int H = p.lv.H[lvl], W = p.lv.W[lvl], start = p.lv.start[lvl];
  const float inv_W = p.lv.inv_W[lvl], inv_H = p.lv.inv_H[lvl];
  // phase B roles
  const int r = (lane >> 3) & 1, h = (lane >> 2) & 1, c4 = lane & 3;
  const uint32_t rec_rd = rec + (uint32_t)((qa * 4 + r * 2 + h) * kRecCorner);
  const uint32_t rec_wr = rec + (uint32_t)(qa * 4 * kRecCorner + pt * 8);

  // parameters of this lane's (query, point); prefetched one pair ahead
  auto load_params = [&](int pair, float& logit, float2& xy, float2& rf) {
    const int q = min(q_beg + 2 * pair + qa, q_end - 1);
    const int64_t bq = (int64_t)b * p.Lq + q;
    logit = __ldg(p.w + bq * p.ldw + m * LP + pt);
    xy = __ldg(reinterpret_cast<const float2*>(p.a + bq * p.lda + (m * LP + pt) * 2));
    rf = make_float2(0.f, 0.f);
    if (FUSED) rf = __ldg(reinterpret_cast<const float2*>(p.ref + (bq * L + lvl) * 2));
  };
  // parameters are fetched a whole group of G pairs ahead (the loads miss L1 by construction: each line is
  // used once), so their latency is covered by G pairs of work instead of one
  constexpr int G = 4;
  float lg[G], n_lg[G];
  float2 xyv[G], rfv[G], n_xyv[G], n_rfv[G];
#pragma unroll
  for (int j = 0; j < G; ++j) load_params(warp + kSlabWarps * j, lg[j], xyv[j], rfv[j]);

  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "SLAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
      "@p bra.uni SLAB_DONE;\n"
      "bra.uni SLAB_WAIT;\n"
      "SLAB_DONE:\n"
      "}\n" ::"r"(bar_a) : "memory");

  for (int base = warp; base < n_pairs; base += kSlabWarps * G) {
#pragma unroll
    for (int j = 0; j < G; ++j) load_params(base + kSlabWarps * (G + j), n_lg[j], n_xyv[j], n_rfv[j]);
#pragma unroll
    for (int j = 0; j < G; ++j) {
    const int pair = base + kSlabWarps * j;
    if (pair >= n_pairs) break;                              // warp-uniform
    const float logit = lg[j];
    const float2 xy = xyv[j], rf = rfv[j];
    // ---- phase A ----
    float a = logit;
    if (FUSED) {
      float mx = a;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float e = __expf(a - mx);
      float sum = e;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      a = e * __fdividef(1.f, sum);
    }
    float lx = xy.x, ly = xy.y;
    if (FUSED) { lx = rf.x + lx * inv_W; ly = rf.y + ly * inv_H; }
    const float x = lx * (float)W - 0.5f, y = ly * (float)H - 0.5f;
    const bool inside = x > -1.f && y > -1.f && x < (float)W && y < (float)H;
    const float xf = floorf(x), yf = floorf(y);
    const float fx = x - xf, fy = y - yf;
    // clamp through float so that wild locations cannot overflow the int conversion
    const int x0 = (int)fminf(fmaxf(xf, -1.f), (float)W), y0 = (int)fminf(fmaxf(yf, -1.f), (float)H);
    const bool xl = inside && x0 >= 0, xh = inside && x0 + 1 < W, yl = y0 >= 0, yh = y0 + 1 < H;
    const int xc0 = min(max(x0, 0), W - 1), xc1 = min(max(x0 + 1, 0), W - 1);
    const int yc0 = min(max(y0, 0), H - 1), yc1 = min(max(y0 + 1, 0), H - 1);
    const float w00 = (xl && yl) ? (1.f - fy) * (1.f - fx) * a : 0.f, w01 = (xh && yl) ? (1.f - fy) * fx * a : 0.f;
    const float w10 = (xl && yh) ? fy * (1.f - fx) * a : 0.f, w11 = (xh && yh) ? fy * fx * a : 0.f;
    const uint32_t o00 = (uint32_t)(start + yc0 * W + xc0) * (D * 4), o01 = (uint32_t)(start + yc0 * W + xc1) * (D * 4);
    const uint32_t o10 = (uint32_t)(start + yc1 * W + xc0) * (D * 4), o11 = (uint32_t)(start + yc1 * W + xc1) * (D * 4);
    __syncwarp();                                            // phase B of the previous pair has read its records
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(rec_wr), "r"(o00), "r"(__float_as_uint(w00)) : "memory");
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(rec_wr + kRecCorner), "r"(o01), "r"(__float_as_uint(w01)) : "memory");
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(rec_wr + 2 * kRecCorner), "r"(o10), "r"(__float_as_uint(w10)) : "memory");
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(rec_wr + 3 * kRecCorner), "r"(o11), "r"(__float_as_uint(w11)) : "memory");
    __syncwarp();
    // ---- phase B ----
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t vbase = slab + (uint32_t)c4 * 16;
#pragma unroll
    for (int s = 0; s < LP; s += 2) {                        // one 128-bit record load covers two sampling points
      uint32_t off0, off1; float wgt0, wgt1;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(off0), "=f"(wgt0), "=r"(off1), "=f"(wgt1) : "r"(rec_rd + s * 8));
      const float4 v0 = lds4(vbase + off0), v1 = lds4(vbase + off1);
      acc.x += wgt0 * v0.x; acc.y += wgt0 * v0.y; acc.z += wgt0 * v0.z; acc.w += wgt0 * v0.w;
      acc.x += wgt1 * v1.x; acc.y += wgt1 * v1.y; acc.z += wgt1 * v1.z; acc.w += wgt1 * v1.w;
    }
#pragma unroll
    for (int o = 4; o <= 8; o <<= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
      acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
      acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
    }
    const int q = q_beg + 2 * pair + qa;
    if (q < q_end && (lane & 12) == 0) st4(p.out + (((int64_t)b * p.Lq + q) * p.M + m) * D + c4 * 4, acc);
    }
#pragma unroll
    for (int j = 0; j < G; ++j) { lg[j] = n_lg[j]; xyv[j] = n_xyv[j]; rfv[j] = n_rfv[j]; }
  }
}

// number of query chunks per (image, head): trade the slab reload (S rows) against wave quantisation
// over 2 resident CTAs per SM
static int slab_nsplit(int BM_, int S, int Lq) {
  int best = 1;
  double best_cost = 1e30;
  for (int ns = 1; ns <= 16; ++ns) {
    const int q = poet_ceil_div(Lq, ns);
    const double waves = (double)poet_ceil_div((int64_t)BM_ * ns, 2 * POET_NUM_SMS);
    const double cost = waves * ((double)S + 80.0 * q);
    if (cost < best_cost) { best_cost = cost; best = ns; }
  }
  return best;
}

// POET_OK if the slab kernel ran, POET_ERR_UNSUPPORTED if the shape does not qualify (caller falls back to the
// general kernel: same results up to summation order).
static int try_slab_fwd(const MsdaArgs& a, int mode, cudaStream_t s) {
  static const int enabled = []() { const char* e = getenv("POET_MSDA_SLAB"); return e ? atoi(e) : 1; }();
  if (!enabled || a.D != 16 || a.L != 4 || a.P != 4) return POET_ERR_UNSUPPORTED;
  const int n_boxes = poet_ceil_div(a.S, kSlabBoxRows);
  const size_t slab_bytes = (size_t)n_boxes * kSlabBoxRows * 64 + 128 + kSlabWarps * kRecBytesPerWarp;
  if (slab_bytes > 112 * 1024) return POET_ERR_UNSUPPORTED;                 // two CTAs per SM
  if ((int64_t)a.Lq * 4 < a.S) return POET_ERR_UNSUPPORTED;                 // decoder rows: the reload would dominate
  if ((int64_t)a.B * a.S >= ((int64_t)1 << 31)) return POET_ERR_UNSUPPORTED;
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  // value viewed as [B*S rows, M*D floats]; box = one head's 16 floats x 64 rows
  if (!poet_tma::encode_2d(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, a.value, (uint64_t)a.M * a.D, (uint64_t)a.B * a.S,
                           (uint64_t)a.M * a.D * 4, 16u, (uint32_t)kSlabBoxRows, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B))
    return POET_ERR_UNSUPPORTED;
  const int nsplit = slab_nsplit(a.B * a.M, a.S, a.Lq);
  const int q_per_cta = poet_ceil_div(a.Lq, nsplit);
  const int grid = a.B * a.M * nsplit;
  auto launch = [&](auto kern) -> int {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)slab_bytes);
    if (e != cudaSuccess) return (int)e;
    poet_launch(kern, dim3(grid), dim3(kSlabThreads), slab_bytes, s, a, tm, nsplit, q_per_cta, n_boxes);
    return poet_launch_status();
  };
  return mode ? launch(msda_fwd_slab_kernel<4, 4, true>) : launch(msda_fwd_slab_kernel<4, 4, false>);
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void red_add4(float* p, float w, float4 v) {
  // one 128-bit reduction per corner (sm_90+): red.global.add.v4.f32.  No "memory" clobber: grad_value is write-only
  // in these kernels, and the clobber would pin every later load behind the reduction
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x * w), "f"(v.y * w), "f"(v.z * w),
               "f"(v.w * w));
}

// predicated form: no branch around the reduction (the compiler cannot predicate a volatile asm statement itself)
__device__ __forceinline__ void red_add4_pred(bool on, float* p, float w, float4 v) {
  asm volatile("{\n\t.reg .pred pp;\n\tsetp.ne.u32 pp, %5, 0;\n\t@pp red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n\t}"
               ::"l"(p), "f"(v.x * w), "f"(v.y * w), "f"(v.z * w), "f"(v.w * w), "r"((uint32_t)on));
}

__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// Register diet (the kernel is latency-bound, so occupancy matters): location gradients leave the
// thread level by level (8 floats per level = two float4 stores split over the group's lanes); only
// the L*P attention gradients stay live until the softmax backward at the end.
template <int L, int P, int G, bool FUSED>
__global__ void __launch_bounds__(256, 3) msda_bwd_kernel(const MsdaArgs p) {
  poet_pdl_entry();
  constexpr int LP = L * P, D = 4 * G;
  static_assert(P % 2 == 0, "points per level must be even (float4 location stores)");
  const int64_t total = (int64_t)p.B * p.Lq * p.M * G;
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = t < total;
  if (!live) t = total - 1;                                  // keep the warp converged for the shuffles
  const int c4 = (int)(t % G);
  const int64_t bqm = t / G;
  const int m = (int)(bqm % p.M);
  const int64_t bq = bqm / p.M;
  const int b = (int)(bq / p.Lq);

  float aw[LP];
  load_row<LP>(p.w + bq * p.ldw + m * LP, aw);
  if (FUSED) softmax_inplace<LP>(aw);

  const float* arow = p.a + bq * p.lda + m * LP * 2;
  float* garow = p.grad_a + bq * p.lda + m * LP * 2;
  const int vstride = p.M * D;
  const int64_t voff = ((int64_t)b * p.S * p.M + m) * D + c4 * 4;
  const float4 go = ldg4(p.grad_out + t * 4);

  float g_attn[LP];
#pragma unroll
  for (int l = 0; l < L; ++l) {
    const int H = p.lv.H[l], W = p.lv.W[l];
    const float* vl = p.value + voff + (int64_t)p.lv.start[l] * vstride;
    float* gl = p.grad_value + voff + (int64_t)p.lv.start[l] * vstride;
    float rx = 0.f, ry = 0.f;
    if (FUSED) { float2 r = __ldg(reinterpret_cast<const float2*>(p.ref + (bq * L + l) * 2)); rx = r.x; ry = r.y; }
    float xy[2 * P], gxy[2 * P];
#pragma unroll
    for (int i = 0; i < 2 * P; i += 4) {
      float4 v = ldg4(arow + l * 2 * P + i);
      xy[i] = v.x; xy[i + 1] = v.y; xy[i + 2] = v.z; xy[i + 3] = v.w;
    }
#pragma unroll
    for (int s = 0; s < P; ++s) {
      float lx = xy[2 * s], ly = xy[2 * s + 1];
      if (FUSED) { lx = rx + lx * p.lv.inv_W[l]; ly = ry + ly * p.lv.inv_H[l]; }
      const float x = lx * (float)W - 0.5f, y = ly * (float)H - 0.5f;
      float ga = 0.f, gx = 0.f, gy = 0.f;
      if (x > -1.f && y > -1.f && x < (float)W && y < (float)H) {
        const float xf = floorf(x), yf = floorf(y);
        const int x0 = (int)xf, y0 = (int)yf;
        const float fx = x - xf, fy = y - yf;
        const float a = aw[l * P + s];
        const bool xl = x0 >= 0, xh = x0 + 1 < W, yl = y0 >= 0, yh = y0 + 1 < H;
        const int i00 = (y0 * W + x0) * vstride;
        const bool scatter = live && l < p.red_levels;        // the dense kernel owns the other levels' grad_value
        // the four corner loads of a point are issued together (predicated), the dots and the reductions follow: a load
        // placed behind a reduction would wait for it (the reduction is a volatile asm statement the compiler orders
        // other memory operations around), which made every corner pay its own L2 round trip
        float4 v00 = make_float4(0.f, 0.f, 0.f, 0.f), v01 = v00, v10 = v00, v11 = v00;
        if (yl && xl) v00 = ldg4(vl + i00);
        if (yl && xh) v01 = ldg4(vl + i00 + vstride);
        if (yh && xl) v10 = ldg4(vl + i00 + W * vstride);
        if (yh && xh) v11 = ldg4(vl + i00 + (W + 1) * vstride);
        const float d00 = dot4(go, v00), d01 = dot4(go, v01), d10 = dot4(go, v10), d11 = dot4(go, v11);
        if (scatter) {
          if (yl && xl) red_add4(gl + i00, a * (1.f - fy) * (1.f - fx), go);
          if (yl && xh) red_add4(gl + i00 + vstride, a * (1.f - fy) * fx, go);
          if (yh && xl) red_add4(gl + i00 + W * vstride, a * fy * (1.f - fx), go);
          if (yh && xh) red_add4(gl + i00 + (W + 1) * vstride, a * fy * fx, go);
        }
        ga = (1.f - fy) * (1.f - fx) * d00 + (1.f - fy) * fx * d01 + fy * (1.f - fx) * d10 + fy * fx * d11;
        // d sampled / d x (pixels) and / d y, times attention weight; pixels = loc * size
        gx = a * ((1.f - fy) * (d01 - d00) + fy * (d11 - d10));
        gy = a * ((1.f - fx) * (d10 - d00) + fx * (d11 - d01));
        // d pixel / d (offset) = W * (1/W) = 1 in fused mode; d pixel / d loc = W in core mode
        if (!FUSED) { gx *= (float)W; gy *= (float)H; }
      }
      g_attn[l * P + s] = group_sum<G>(ga);
      gxy[2 * s] = group_sum<G>(gx);
      gxy[2 * s + 1] = group_sum<G>(gy);
    }
    // this level's location gradients: 2P floats = P/2 float4 chunks, chunk j written by lane j % G
    if (live) {
#pragma unroll
      for (int j = 0; j < P / 2; ++j)
        if ((l * (P / 2) + j) % G == c4)
          st4(garow + l * 2 * P + j * 4, make_float4(gxy[4 * j], gxy[4 * j + 1], gxy[4 * j + 2], gxy[4 * j + 3]));
    }
  }

  if (FUSED) {
    // softmax backward: g_logit_i = a_i (g_attn_i - sum_j a_j g_attn_j)
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < LP; ++i) dot += aw[i] * g_attn[i];
#pragma unroll
    for (int i = 0; i < LP; ++i) g_attn[i] = aw[i] * (g_attn[i] - dot);
  }
  if (!live) return;
  float* gw = p.grad_w + bq * p.ldw + m * LP;
#pragma unroll
  for (int j = 0; j < LP / 4; ++j)
    if (j % G == c4) st4(gw + j * 4, make_float4(g_attn[j * 4], g_attn[j * 4 + 1], g_attn[j * 4 + 2], g_attn[j * 4 + 3]));
}


// ------------------------------------------------------------------------------------------
// backward, L = P = 4: the per-point arithmetic shared by the lanes of a (q,m) group
// ------------------------------------------------------------------------------------------
// msda_bwd_kernel lets each of the G lanes of a (b,q,m) group recompute the softmax, the sampling location, the cell and
// the bilinear fractions of all 16 points -- ~35 instructions per point and lane that differ only in the channel group
// (tools/msda_red_bisect.sh: the kernel is half gather-side arithmetic).  Here lane l of a group prepares the four points
// of level l once (one logits float4, two location float4s, its level's constants) and the group reads them by shuffle
// while it walks the 16 points: 4 shuffles per point instead of the redundant arithmetic, four live attention / location
// gradients per lane instead of sixteen (64 registers -> 4 CTAs per SM).  Corner loads, dot products and the
// red.global.add.v4.f32 scatter are unchanged, so are the results (same operations per value, group sums in the same order).
template <int G, bool FUSED>
__global__ void __launch_bounds__(256, 4) msda_bwd_shared_kernel(const MsdaArgs p) {
  poet_pdl_entry();
  constexpr int L = 4, P = 4, LP = 16, D = 4 * G;
  static_assert(G >= 4, "one preparing lane per level");
  constexpr int BIAS = 1 << 20;                                // pixel offsets travel as (px + BIAS) << 4 | corner flags
  const int64_t total = (int64_t)p.B * p.Lq * p.M * G;
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = t < total;
  if (!live) t = total - 1;                                    // keep the warp converged for the shuffles
  const int c4 = (int)(t % G);
  const int64_t bqm = t / G;
  const int m = (int)(bqm % p.M);
  const int64_t bq = bqm / p.M;
  const int b = (int)(bq / p.Lq);
  const int lane = threadIdx.x & 31;
  const int grp = lane & ~(G - 1);                             // first lane of this (b,q,m)
  const int lvl = c4 & 3;                                      // the level whose points this lane prepares

  // ---- this lane's four points ----
  float aw[P];
  {
    const float4 lg = ldg4(p.w + bq * p.ldw + m * LP + lvl * P);
    aw[0] = lg.x; aw[1] = lg.y; aw[2] = lg.z; aw[3] = lg.w;
  }
  if (FUSED) {                                                 // softmax over the 16 logits of the group (lanes 0..3 of each quad hold them)
    float mx = fmaxf(fmaxf(aw[0], aw[1]), fmaxf(aw[2], aw[3]));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < P; ++k) { aw[k] = __expf(aw[k] - mx); sum += aw[k]; }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const float inv = __fdividef(1.f, sum);
#pragma unroll
    for (int k = 0; k < P; ++k) aw[k] *= inv;
  }
  int code[P];
  float fxv[P], fyv[P];
  {
    const int Hl = p.lv.H[lvl], Wl = p.lv.W[lvl];
    const float* arow = p.a + bq * p.lda + m * LP * 2 + lvl * 2 * P;
    const float4 xa = ldg4(arow), xb = ldg4(arow + 4);
    const float xy[2 * P] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
    float rx = 0.f, ry = 0.f, iw = 0.f, ih = 0.f;
    if (FUSED) {
      const float2 r = __ldg(reinterpret_cast<const float2*>(p.ref + (bq * L + lvl) * 2));
      rx = r.x; ry = r.y; iw = p.lv.inv_W[lvl]; ih = p.lv.inv_H[lvl];
    }
#pragma unroll
    for (int k = 0; k < P; ++k) {
      float lx = xy[2 * k], ly = xy[2 * k + 1];
      if (FUSED) { lx = rx + lx * iw; ly = ry + ly * ih; }
      const float x = lx * (float)Wl - 0.5f, y = ly * (float)Hl - 0.5f;
      const bool inside = x > -1.f && y > -1.f && x < (float)Wl && y < (float)Hl;
      const float xf = floorf(x), yf = floorf(y);
      fxv[k] = x - xf; fyv[k] = y - yf;
      // clamp through float so that wild locations cannot overflow the int conversion
      const int x0 = (int)fminf(fmaxf(xf, -1.f), (float)Wl), y0 = (int)fminf(fmaxf(yf, -1.f), (float)Hl);
      const bool xl = x0 >= 0, xh = x0 + 1 < Wl, yl = y0 >= 0, yh = y0 + 1 < Hl;
      const int flags = inside ? ((yl && xl) ? 1 : 0) | ((yl && xh) ? 2 : 0) | ((yh && xl) ? 4 : 0) | ((yh && xh) ? 8 : 0) : 0;
      code[k] = ((y0 * Wl + x0 + BIAS) << 4) | flags;
    }
  }

  // ---- the group walks the 16 points ----
  const int vstride = p.M * D;
  const int64_t voff = ((int64_t)b * p.S * p.M + m) * D + c4 * 4;
  const float4 go = ldg4(p.grad_out + t * 4);
  float ga_own[P], gx_own[P], gy_own[P];
#pragma unroll
  for (int k = 0; k < P; ++k) { ga_own[k] = 0.f; gx_own[k] = 0.f; gy_own[k] = 0.f; }
#pragma unroll
  for (int l = 0; l < L; ++l) {
    const int W = p.lv.W[l];
    const float* vl = p.value + voff + (int64_t)p.lv.start[l] * vstride;
    float* gl = p.grad_value + voff + (int64_t)p.lv.start[l] * vstride;
    const bool scatter = live && l < p.red_levels;            // the dense kernel owns the other levels' grad_value
#pragma unroll
    for (int s = 0; s < P; ++s) {
      const int cd = __shfl_sync(0xffffffffu, code[s], grp + l);
      const float a = __shfl_sync(0xffffffffu, aw[s], grp + l);
      const float fx = __shfl_sync(0xffffffffu, fxv[s], grp + l);
      const float fy = __shfl_sync(0xffffffffu, fyv[s], grp + l);
      float ga = 0.f, gx = 0.f, gy = 0.f;
      if (cd & 15) {
        const int i00 = ((cd >> 4) - BIAS) * vstride;
        const bool c00 = cd & 1, c01 = cd & 2, c10 = cd & 4, c11 = cd & 8;
        // the four corner loads of a point are issued together (predicated), the dots and the reductions follow
        float4 v00 = make_float4(0.f, 0.f, 0.f, 0.f), v01 = v00, v10 = v00, v11 = v00;
        if (c00) v00 = ldg4(vl + i00);
        if (c01) v01 = ldg4(vl + i00 + vstride);
        if (c10) v10 = ldg4(vl + i00 + W * vstride);
        if (c11) v11 = ldg4(vl + i00 + (W + 1) * vstride);
        const float d00 = dot4(go, v00), d01 = dot4(go, v01), d10 = dot4(go, v10), d11 = dot4(go, v11);
        if (scatter) {
          if (c00) red_add4(gl + i00, a * (1.f - fy) * (1.f - fx), go);
          if (c01) red_add4(gl + i00 + vstride, a * (1.f - fy) * fx, go);
          if (c10) red_add4(gl + i00 + W * vstride, a * fy * (1.f - fx), go);
          if (c11) red_add4(gl + i00 + (W + 1) * vstride, a * fy * fx, go);
        }
        ga = (1.f - fy) * (1.f - fx) * d00 + (1.f - fy) * fx * d01 + fy * (1.f - fx) * d10 + fy * fx * d11;
        // d sampled / d x (pixels) and / d y, times attention weight; pixels = loc * size
        gx = a * ((1.f - fy) * (d01 - d00) + fy * (d11 - d10));
        gy = a * ((1.f - fx) * (d10 - d00) + fx * (d11 - d01));
        // d pixel / d (offset) = W * (1/W) = 1 in fused mode; d pixel / d loc = W in core mode
        if (!FUSED) { gx *= (float)W; gy *= (float)p.lv.H[l]; }
      }
      ga = group_sum<G>(ga); gx = group_sum<G>(gx); gy = group_sum<G>(gy);
      if (c4 == l) { ga_own[s] = ga; gx_own[s] = gx; gy_own[s] = gy; }
    }
  }

  // ---- lane l < 4 of the group owns level l's gradients ----
  float dotp = 0.f;
  if (FUSED) {
    // softmax backward: g_logit_i = a_i (g_attn_i - sum_j a_j g_attn_j)
#pragma unroll
    for (int k = 0; k < P; ++k) dotp += aw[k] * ga_own[k];
    dotp += __shfl_xor_sync(0xffffffffu, dotp, 1);
    dotp += __shfl_xor_sync(0xffffffffu, dotp, 2);
#pragma unroll
    for (int k = 0; k < P; ++k) ga_own[k] = aw[k] * (ga_own[k] - dotp);
  }
  if (!live || c4 >= L) return;
  float* garow = p.grad_a + bq * p.lda + m * LP * 2 + c4 * 2 * P;
  st4(garow, make_float4(gx_own[0], gy_own[0], gx_own[1], gy_own[1]));
  st4(garow + 4, make_float4(gx_own[2], gy_own[2], gx_own[3], gy_own[3]));
  st4(p.grad_w + bq * p.ldw + m * LP + c4 * P, make_float4(ga_own[0], ga_own[1], ga_own[2], ga_own[3]));
}

// ------------------------------------------------------------------------------------------
// backward, grad_value of the low-resolution levels as a dense tensor-core product
// ------------------------------------------------------------------------------------------
// The scatter formulation issues one 16-byte global reduction per (corner, 4 channels): 64 per (q,m), and three
// quarters of them land on the few hundred pixels of levels 1..3 (REF pyramid: 300 + 80 + 20 of 1600).  The SM's
// reduction issue rate (0.86 cycles per lane-op) is what bounds that kernel.  For those levels the same sum is a
// small dense product per (image, head):
//
//     grad_value[px, :] = sum_q W[px, q] * grad_out[q, :],   W[px, q] = sum over the sampling points of q at that
//                                                            level of attention * bilinear corner weight at px
//
// W is built in shared memory, one COLUMN per query owned by one thread per level (its own points are serialised in
// the thread, columns of different queries never meet: no atomics), and multiplied on the tensor cores with
// mma.sync m16n8k16: M = pixels, N = D = 16 channels, K = queries.  Operands are split into fp16 hi/lo pairs
// (x = hi + lo to 2^-22; grad_out is scaled by a power of two per work item so that fp16's range fits), three MMAs per
// product with fp32 accumulation: fp32-grade like the split-bf16 GEMMs.  The accumulators of a work item
// (image, head, query chunk) live in registers across its query tiles and leave as one reduction per element at the
// end.  This is legacy-path tensor work on purpose: 3 x 2 x 16 x 400 x 16 flops per query tile is latency-, not
// throughput-bound, and tcgen05's 128-row tiles would be 97 % padding on the N = 16 side.
namespace dense {

constexpr int KQ = 32;              // queries per tile (K of the product)
constexpr int QS = 40;              // row stride of W in floats: QS % 32 == 8 makes the A-fragment loads conflict-free
constexpr int THREADS = 128;        // 4 warps; thread = (query of the tile, role): roles 0..2 build one level each, role 3 stages grad_out
constexpr int MAX_MT = 7;           // m-tiles (16 pixels) per warp: 4 * 7 * 16 = 448 dense pixels at most
constexpr int GS = 20;              // row stride (32-bit words) of the transposed fp16 grad_out tile [d][q/2]

__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void split_half2(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x, y);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(x - hf.x, y - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

template <int L, int P, bool FUSED>
__global__ void __launch_bounds__(THREADS, 3)
msda_bwd_dense_kernel(const MsdaArgs p, int lvl0, int npx, int n_mt, int nchunk, int q_per_chunk, int n_items) {
  poet_pdl_entry();
  constexpr int LP = L * P, D = 16;
  extern __shared__ __align__(16) uint8_t dense_smem[];
  float* Wt = reinterpret_cast<float*>(dense_smem);                                  // [n_mt * 16][QS]
  uint32_t* gh = reinterpret_cast<uint32_t*>(Wt + (size_t)n_mt * 16 * QS);           // [D][GS] fp16 pairs (q even | q odd)
  uint32_t* gl = gh + D * GS;
  __shared__ float s_max[THREADS / 32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int ql = tid & 31, role = tid >> 5;                    // build-phase identity
  const int lvl = lvl0 + role;                                 // level this thread builds (role 3 or lvl >= L: none)
  const bool builder = lvl < L;
  const int pix0 = p.lv.start[lvl0];                           // first dense pixel in the flattened map

  for (int i = tid; i < n_mt * 16 * QS; i += THREADS) Wt[i] = 0.f;
  __syncthreads();

  int H = 1, Wd = 1, lstart = 0;
  float inv_W = 0.f, inv_H = 0.f;
  if (builder) { H = p.lv.H[lvl]; Wd = p.lv.W[lvl]; lstart = p.lv.start[lvl] - pix0; inv_W = p.lv.inv_W[lvl]; inv_H = p.lv.inv_H[lvl]; }

  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int chunk = item % nchunk, bm = item / nchunk;
    const int m = bm % p.M, b = bm / p.M;
    const int q_beg = chunk * q_per_chunk, q_end = min(p.Lq, q_beg + q_per_chunk);
    if (q_beg >= q_end) continue;
    // ---- power-of-two scale of this item's grad_out rows (fp16 range) ----
    float mx = 0.f;
    for (int i = tid; i < (q_end - q_beg) * (D / 4); i += THREADS) {
      const int q = q_beg + i / (D / 4), c = i % (D / 4);
      const float4 v = ldg4(p.grad_out + (((int64_t)b * p.Lq + q) * p.M + m) * D + c * 4);
      mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    __syncthreads();                                            // previous item's readers of s_max are done
    if (lane == 0) s_max[warp] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(s_max[0], s_max[1]), fmaxf(s_max[2], s_max[3]));
    if (!(mx > 0.f) || !(mx < INFINITY)) continue;             // all-zero gradient: nothing to add (uniform over the CTA)
    int e;
    (void)frexpf(mx, &e);                                       // mx = f * 2^e, f in [0.5, 1)
    const float scale = ldexpf(1.f, 10 - e), inv_scale = ldexpf(1.f, e - 10);   // scaled maximum in [512, 1024)

    float acc[MAX_MT][2][4];
#pragma unroll
    for (int i = 0; i < MAX_MT; ++i)
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[i][n][k] = 0.f;

    for (int q0 = q_beg; q0 < q_end; q0 += KQ) {
      const int q = q0 + ql;
      const bool q_ok = q < q_end;
      uint32_t touched[2 * P];                                  // W word offsets written by this thread (pairs packed), 0xffff: none
#pragma unroll
      for (int i = 0; i < 2 * P; ++i) touched[i] = 0xffffffffu;
      if (builder && q_ok) {
        // ---- W columns: this thread's level of query q ----
        const int64_t bq = (int64_t)b * p.Lq + q;
        float aw[P];
        if (FUSED) {
          float lg[LP];
          load_row<LP>(p.w + bq * p.ldw + m * LP, lg);
          softmax_inplace<LP>(lg);
#pragma unroll
          for (int s = 0; s < P; ++s) {                         // level index is runtime: select without dynamic indexing
            float v = lg[s];
#pragma unroll
            for (int l2 = 1; l2 < L; ++l2) v = (lvl == l2) ? lg[l2 * P + s] : v;
            aw[s] = v;
          }
        } else {
          const float4 v = ldg4(p.w + bq * p.ldw + m * LP + lvl * P);
          aw[0] = v.x; aw[1] = v.y; aw[2] = v.z; aw[3] = v.w;
          static_assert(P == 4, "attention row load assumes four points per level");
        }
        float xy[2 * P];
        const float* arow = p.a + bq * p.lda + (m * LP + lvl * P) * 2;
#pragma unroll
        for (int i = 0; i < 2 * P; i += 4) {
          const float4 v = ldg4(arow + i);
          xy[i] = v.x; xy[i + 1] = v.y; xy[i + 2] = v.z; xy[i + 3] = v.w;
        }
        float rx = 0.f, ry = 0.f;
        if (FUSED) { const float2 r = __ldg(reinterpret_cast<const float2*>(p.ref + (bq * L + lvl) * 2)); rx = r.x; ry = r.y; }
#pragma unroll
        for (int s = 0; s < P; ++s) {
          float lx = xy[2 * s], ly = xy[2 * s + 1];
          if (FUSED) { lx = rx + lx * inv_W; ly = ry + ly * inv_H; }
          const float x = lx * (float)Wd - 0.5f, y = ly * (float)H - 0.5f;
          if (x > -1.f && y > -1.f && x < (float)Wd && y < (float)H) {
            const float xf = floorf(x), yf = floorf(y);
            const int x0 = (int)xf, y0 = (int)yf;
            const float fx = x - xf, fy = y - yf;
            const float a = aw[s];
            const bool xl = x0 >= 0, xh = x0 + 1 < Wd, yl = y0 >= 0, yh = y0 + 1 < H;
            const int r00 = (lstart + y0 * Wd + x0) * QS + ql;   // W word of corner (y0, x0) in this query's column
            // same weight expressions as the scatter kernel; the thread's own points are serialised: plain RMW
            if (yl && xl) Wt[r00] += a * (1.f - fy) * (1.f - fx);
            if (yl && xh) Wt[r00 + QS] += a * (1.f - fy) * fx;
            if (yh && xl) Wt[r00 + Wd * QS] += a * fy * (1.f - fx);
            if (yh && xh) Wt[r00 + (Wd + 1) * QS] += a * fy * fx;
            const uint32_t top = (yl ? (uint32_t)(lstart + y0 * Wd + x0 + (xl ? 0 : 1)) : 0xffffu);       // first valid pixel of the row pair
            const uint32_t bot = (yh ? (uint32_t)(lstart + (y0 + 1) * Wd + x0 + (xl ? 0 : 1)) : 0xffffu);
            const uint32_t both = (xl && xh) ? 0x8000u : 0u;    // two x-adjacent pixels in each valid row
            touched[2 * s] = top | both;
            touched[2 * s + 1] = bot | both;
          }
        }
      } else if (role == 3) {
        // ---- grad_out tile: scaled, split into fp16 hi / lo, transposed to [d][q] ----
        float v[D];
        if (q_ok) {
          const float* gp = p.grad_out + (((int64_t)b * p.Lq + q) * p.M + m) * D;
#pragma unroll
          for (int c = 0; c < D; c += 4) { const float4 x = ldg4(gp + c); v[c] = x.x; v[c + 1] = x.y; v[c + 2] = x.z; v[c + 3] = x.w; }
        } else {
#pragma unroll
          for (int c = 0; c < D; ++c) v[c] = 0.f;
        }
        __half* ghh = reinterpret_cast<__half*>(gh);
        __half* glh = reinterpret_cast<__half*>(gl);
#pragma unroll
        for (int c = 0; c < D; ++c) {
          const float x = v[c] * scale;
          const __half h = __float2half_rn(x);
          ghh[(c * GS) * 2 + ql] = h;
          glh[(c * GS) * 2 + ql] = __float2half_rn(x - __half2float(h));
        }
      }
      __syncthreads();
      // ---- acc[px, d] += W[px, q-tile] . grad_out[q-tile, d] ----
#pragma unroll
      for (int ks = 0; ks < KQ / 16; ++ks) {
        uint32_t bh[2][2], bl[2][2];
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          const int w0 = (n * 8 + g) * GS + ks * 8 + t;
          bh[n][0] = gh[w0]; bh[n][1] = gh[w0 + 4];
          bl[n][0] = gl[w0]; bl[n][1] = gl[w0 + 4];
        }
#pragma unroll
        for (int i = 0; i < MAX_MT; ++i) {
          const int mt = warp + 4 * i;
          if (mt < n_mt) {                                      // warp-uniform
            const float* wp = Wt + (mt * 16 + g) * QS + ks * 16 + 2 * t;
            const float2 x0 = *reinterpret_cast<const float2*>(wp);
            const float2 x1 = *reinterpret_cast<const float2*>(wp + 8 * QS);
            const float2 x2 = *reinterpret_cast<const float2*>(wp + 8);
            const float2 x3 = *reinterpret_cast<const float2*>(wp + 8 * QS + 8);
            uint32_t ah[4], al[4];
            split_half2(x0.x, x0.y, ah[0], al[0]);
            split_half2(x1.x, x1.y, ah[1], al[1]);
            split_half2(x2.x, x2.y, ah[2], al[2]);
            split_half2(x3.x, x3.y, ah[3], al[3]);
#pragma unroll
            for (int n = 0; n < 2; ++n) {
              mma_f16(acc[i][n], ah, bl[n][0], bl[n][1]);       // small cross terms first
              mma_f16(acc[i][n], al, bh[n][0], bh[n][1]);
              mma_f16(acc[i][n], ah, bh[n][0], bh[n][1]);
            }
          }
        }
      }
      __syncthreads();
      // ---- clear exactly the words this thread wrote (its column is private) ----
      if (builder && q_ok) {
#pragma unroll
        for (int i = 0; i < 2 * P; ++i) {
          const uint32_t tw = touched[i];
          if ((tw & 0x7fffu) != 0x7fffu) {
            const int r = (int)(tw & 0x7fffu) * QS + ql;
            Wt[r] = 0.f;
            if (tw & 0x8000u) Wt[r + QS] = 0.f;
          }
        }
      }
    }
    // ---- flush: one 8-byte reduction per accumulator pair ----
    float* gv = p.grad_value + (((int64_t)b * p.S + pix0) * p.M + m) * D;
    const int64_t vstride = (int64_t)p.M * D;
#pragma unroll
    for (int i = 0; i < MAX_MT; ++i) {
      const int mt = warp + 4 * i;
      if (mt < n_mt) {
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          const int px0 = mt * 16 + g, px1 = px0 + 8, d = n * 8 + 2 * t;
          if (px0 < npx)
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(gv + px0 * vstride + d), "f"(acc[i][n][0] * inv_scale),
                         "f"(acc[i][n][1] * inv_scale) : "memory");
          if (px1 < npx)
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(gv + px1 * vstride + d), "f"(acc[i][n][2] * inv_scale),
                         "f"(acc[i][n][3] * inv_scale) : "memory");
        }
      }
    }
  }
}

}  // namespace dense

// ------------------------------------------------------------------------------------------
// backward, many queries (the encoder): one CTA per (image, head, run of 64-query tiles); the low-resolution
// levels leave the memory pipes
// ------------------------------------------------------------------------------------------
// msda_bwd_shared_kernel spends its time in the load/store pipe: 64 corner loads and 64 16-byte global reductions per
// (q,m), and half of the reductions of the REF pyramid land on the 100 pixels of levels 2-3 (same-address traffic in L2).
// For the trailing levels that hold <= 104 pixels together (the "dense" levels) both directions are small dense products
// per (image, head) over a tile of 64 queries, done with 3xTF32 mma.sync (x = hi + lo, hi.lo + lo.hi + hi.hi, fp32
// accumulation: ~2^-21 relative, the same grade as the split-bf16 GEMMs):
//
//   gather side    G[q, px]  = sum_c grad_out[q, c] * value[px, c]         (M = 64 queries, N = pixels, K = D = 16)
//                  -> a corner's dot product is ONE shared-memory word G[q, px(corner)] instead of a 64-byte load and
//                     16 FMAs spread over 4 lanes + 2 shuffles;
//   scatter side   grad_value[px, c] += sum_q W[q, px] * grad_out[q, c]    (M = pixels, N = 16, K = 64 queries)
//                  with W[q, px] = sum over q's points of attention * bilinear weight at px, accumulated in shared
//                  memory by the ONE lane that owns (query, level): rows of different queries and pixel ranges of
//                  different levels never meet, a lane's own points are serialised -> plain read-modify-write, no atomics.
//                  The accumulators stay in registers over the CTA's tiles and leave as one 8-byte reduction per pair.
//
// The high-resolution ("sparse") levels keep the gather / red.global.add.v4.f32 scheme of msda_bwd_shared_kernel (same
// per-point record sharing by shuffle).  With one head per CTA, lane c4 of a (q,m) group owns level c4: the dense levels
// need no cross-lane traffic at all.  Per launch at cfg2 this removes half of the corner loads and half of the global
// reductions (52 M of 105 M lane-ops; what remains of them for the dense levels is ~1 M 8-byte reductions).
namespace tile {

constexpr int QT = 64;               // queries per tile at D = 16 (eligibility threshold: Lq >= 2 * QT)
constexpr int THREADS = 256;         // G = D/4 lanes per (q,m): lane cg = channel group in the sparse levels, lanes 0..3 own one level each
constexpr int NPX = 104;             // dense pixels at most (13 n-tiles of 8)
constexpr int LDW = 104;             // row stride of G and W (floats): 104 % 32 == 8 -> conflict-free fragment accesses
// per head size D = 4 G (16 or 32): tile of 256 / G queries, grad_out / value tiles with D + 4 floats per row
// ((D + 4) g + t distinct mod 32 for g < 8, t < 4: conflict-free A / B fragment loads in P1)
template <int G>
struct Cfg {
  static constexpr int D = 4 * G;
  static constexpr int QT = THREADS / G;                 // 64 / 32 queries per tile
  static constexpr int LDC = D + 4;                      // row stride of the grad_out and value tiles
  static constexpr int KS1 = D / 8;                      // k-steps of P1 (reduction over channels)
  static constexpr int NT3 = D / 8;                      // n-tiles of P3 (8 channels each)
  static constexpr int MT1 = QT / 16;                    // m-tiles of P1 (16 queries each)
  static constexpr int NPW1 = (13 + 8 / MT1 - 1) / (8 / MT1);   // n-tiles of P1 per warp (8 warps over MT1 x 13 tiles)
  static constexpr int SMEM_FLOATS = QT * LDW + (QT * LDW + 8) + QT * LDC + NPX * LDC;
};

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// x = hi + lo with hi = the 19 leading bits (a tf32 operand as the tensor core reads it), lo = the exact remainder
// (the tensor core reads ITS 19 leading bits: 22 mantissa bits of x survive)
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_3x(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0,
                                       uint32_t bh1, uint32_t bl0, uint32_t bl1) {
  mma_tf32(c, ah, bl0, bl1);                                   // small cross terms first
  mma_tf32(c, al, bh0, bh1);
  mma_tf32(c, ah, bh0, bh1);
}

template <int G, bool FUSED>
__global__ void __launch_bounds__(THREADS, 3)
msda_bwd_tile_kernel(const MsdaArgs p, int ld, int npx, int nchunk, int tiles_per_chunk, int n_tiles) {
  poet_pdl_entry();
  using C = Cfg<G>;
  constexpr int L = 4, P = 4, LP = 16, D = C::D, QT = C::QT, LDC = C::LDC;
  constexpr int BIAS = 1 << 20;                                // pixel offsets travel as (px + BIAS) << 4 | corner flags
  extern __shared__ __align__(16) float tile_smem[];
  float* Gs = tile_smem;                                       // [QT][LDW]  dot(grad_out[q], value[px])
  float* Ws = Gs + QT * LDW;                                   // [QT][LDW]  (+8: the last m-tile's fragment rows run past px 103)
  float* gos = Ws + QT * LDW + 8;                              // [QT][LDC] grad_out rows of the tile
  float* Vs = gos + QT * LDC;                                  // [NPX][LDC] value rows of the dense levels, this (image, head)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;                       // mma fragment coordinates
  const int c4 = tid % G, ql = tid / G;                        // point-walk identity: lane of the (q,m) group (channel group), query of the tile
  const int grp = lane & ~(G - 1);
  const bool owner = c4 < L;                                   // lanes 0..3 of a group prepare / own one level each
  const int chunk = blockIdx.x % nchunk, bm = blockIdx.x / nchunk;
  const int m = bm % p.M, b = bm / p.M;
  const int tile_beg = chunk * tiles_per_chunk, tile_end = min(n_tiles, tile_beg + tiles_per_chunk);
  if (tile_beg >= tile_end) return;
  const int pix0 = p.lv.start[ld];                             // first dense pixel of the flattened map
  const int n_nt = (npx + 7) >> 3, n_mt = (npx + 15) >> 4;
  const int vstride = p.M * D;

  // ---- value rows of the dense levels (zero rows behind npx) ----
  for (int i = tid; i < NPX * G; i += THREADS) {
    const int px = i / G, c = i % G;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (px < npx) v = ldg4(p.value + ((int64_t)b * p.S + pix0 + px) * vstride + m * D + c * 4);
    *reinterpret_cast<float4*>(Vs + px * LDC + c * 4) = v;
  }

  // this lane's level (point preparation; owner of the level's gradients)
  const int lvl = c4 & 3;                                      // lanes 4.. of a wide group repeat the preparation (never read)
  const int Hl = p.lv.H[lvl], Wl = p.lv.W[lvl];
  const int dense_off = p.lv.start[lvl] - pix0;                // this level's first pixel in G / W columns (dense levels only)
  float rx_iw = 0.f, ry_ih = 0.f;
  if (FUSED) { rx_iw = p.lv.inv_W[lvl]; ry_ih = p.lv.inv_H[lvl]; }

  float acc[C::NT3][4];                                        // P3 accumulators: m-tile = warp, D/8 n-tiles of 8 channels
#pragma unroll
  for (int n = 0; n < C::NT3; ++n)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[n][k] = 0.f;

  for (int tile = tile_beg; tile < tile_end; ++tile) {
    const int q = tile * QT + ql;
    const bool live = q < p.Lq;
    const int64_t bq = (int64_t)b * p.Lq + (live ? q : p.Lq - 1);

    // ---- P0: parameters of this lane's four points, grad_out row, W = 0 ----
    float aw[P];
    {
      const float4 lg = ldg4(p.w + bq * p.ldw + m * LP + lvl * P);
      aw[0] = lg.x; aw[1] = lg.y; aw[2] = lg.z; aw[3] = lg.w;
    }
    const float* arow = p.a + bq * p.lda + m * LP * 2 + lvl * 2 * P;
    const float4 xa = ldg4(arow), xb = ldg4(arow + 4);
    float2 rf = make_float2(0.f, 0.f);
    if (FUSED) rf = __ldg(reinterpret_cast<const float2*>(p.ref + (bq * L + lvl) * 2));
    float4 go = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) go = ldg4(p.grad_out + (bq * p.M + m) * D + c4 * 4);
    *reinterpret_cast<float4*>(gos + ql * LDC + c4 * 4) = go;
    for (int i = tid; i < QT * LDW / 4; i += THREADS) reinterpret_cast<float4*>(Ws)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    // ---- P1: G[q, px] = grad_out[q, :] . value[px, :] ----
    {
      const int mt = warp % C::MT1, nt_beg = (warp / C::MT1) * C::NPW1, nt_end = min(n_nt, nt_beg + C::NPW1);
      uint32_t ah[C::KS1][4], al[C::KS1][4];
#pragma unroll
      for (int ks = 0; ks < C::KS1; ++ks) {
        const float* ap = gos + (mt * 16 + g) * LDC + ks * 8 + t;
        split_tf32(ap[0], ah[ks][0], al[ks][0]);
        split_tf32(ap[8 * LDC], ah[ks][1], al[ks][1]);
        split_tf32(ap[4], ah[ks][2], al[ks][2]);
        split_tf32(ap[8 * LDC + 4], ah[ks][3], al[ks][3]);
      }
      for (int nt = nt_beg; nt < nt_end; ++nt) {
        float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < C::KS1; ++ks) {
          const float* bp = Vs + (nt * 8 + g) * LDC + ks * 8 + t;
          uint32_t bh0, bl0, bh1, bl1;
          split_tf32(bp[0], bh0, bl0);
          split_tf32(bp[4], bh1, bl1);
          mma_3x(c, ah[ks], al[ks], bh0, bh1, bl0, bl1);
        }
        float* gp = Gs + (mt * 16 + g) * LDW + nt * 8 + 2 * t;
        *reinterpret_cast<float2*>(gp) = make_float2(c[0], c[1]);
        *reinterpret_cast<float2*>(gp + 8 * LDW) = make_float2(c[2], c[3]);
      }
    }

    // ---- this lane's four points: softmax, location, cell, fractions (as msda_bwd_shared_kernel) ----
    if (FUSED) {
      float mx = fmaxf(fmaxf(aw[0], aw[1]), fmaxf(aw[2], aw[3]));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < P; ++k) { aw[k] = __expf(aw[k] - mx); sum += aw[k]; }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      const float inv = __fdividef(1.f, sum);
#pragma unroll
      for (int k = 0; k < P; ++k) aw[k] *= inv;
    }
    int code[P];
    float fxv[P], fyv[P];
    {
      const float xy[2 * P] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
      for (int k = 0; k < P; ++k) {
        float lx = xy[2 * k], ly = xy[2 * k + 1];
        if (FUSED) { lx = rf.x + lx * rx_iw; ly = rf.y + ly * ry_ih; }
        const float x = lx * (float)Wl - 0.5f, y = ly * (float)Hl - 0.5f;
        const bool inside = x > -1.f && y > -1.f && x < (float)Wl && y < (float)Hl;
        const float xf = floorf(x), yf = floorf(y);
        fxv[k] = x - xf; fyv[k] = y - yf;
        // clamp through float so that wild locations cannot overflow the int conversion
        const int x0 = (int)fminf(fmaxf(xf, -1.f), (float)Wl), y0 = (int)fminf(fmaxf(yf, -1.f), (float)Hl);
        const bool xl = x0 >= 0, xh = x0 + 1 < Wl, yl = y0 >= 0, yh = y0 + 1 < Hl;
        const int flags = inside ? ((yl && xl) ? 1 : 0) | ((yl && xh) ? 2 : 0) | ((yh && xl) ? 4 : 0) | ((yh && xh) ? 8 : 0) : 0;
        code[k] = ((y0 * Wl + x0 + BIAS) << 4) | flags;
      }
    }
    float ga_own[P];                                           // d loss / d attention of this lane's level (softmax backward needs all 16)
#pragma unroll
    for (int k = 0; k < P; ++k) ga_own[k] = 0.f;
    float* garow = p.grad_a + bq * p.lda + m * LP * 2 + c4 * 2 * P;   // this lane's level: 4 points x (x, y)

    // ---- P2w: dense levels, the owner lane adds its four points into its W row (zeroed before the barrier above; read by
    // P3 after the next one).  Nothing here depends on G, so it sits in front of the sparse walk.
    if (owner && c4 >= ld) {
      float* Wrow = Ws + ql * LDW + dense_off;
#pragma unroll
      for (int s2 = 0; s2 < P; ++s2) {
        const int cd = code[s2];
        const int i00 = (cd >> 4) - BIAS;
        const float a = aw[s2], fx = fxv[s2], fy = fyv[s2];
        // the lane's points are serialised (two of them may share a pixel): plain read-modify-write of its own row
        if (cd & 1) Wrow[i00] += a * (1.f - fy) * (1.f - fx);
        if (cd & 2) Wrow[i00 + 1] += a * (1.f - fy) * fx;
        if (cd & 4) Wrow[i00 + Wl] += a * fy * (1.f - fx);
        if (cd & 8) Wrow[i00 + Wl + 1] += a * fy * fx;
      }
    }

    // ---- P2a: sparse levels, the group walks the points together (global gathers and reductions) ----
    // One point in flight per warp.  Issuing the corner loads of point i+1 before point i is consumed was measured
    // (it needs 128 registers -> 2 CTAs per SM): slower, 350 vs 314 us -- the kernel is not bound by load latency.
    {
      const int64_t voff = ((int64_t)b * p.S * p.M + m) * D + c4 * 4;
#pragma unroll
      for (int l = 0; l < L; ++l) {
        if (l < ld) {                                          // uniform over the grid
          const int W = p.lv.W[l];
          float gxl[P], gyl[P];
#pragma unroll
          for (int s2 = 0; s2 < P; ++s2) {
            const int cd = __shfl_sync(0xffffffffu, code[s2], grp + l);
            const float a = __shfl_sync(0xffffffffu, aw[s2], grp + l);
            const float fx = __shfl_sync(0xffffffffu, fxv[s2], grp + l);
            const float fy = __shfl_sync(0xffffffffu, fyv[s2], grp + l);
            const int64_t o00 = voff + ((int64_t)p.lv.start[l] + ((cd >> 4) - BIAS)) * vstride;
            const float* vp = p.value + o00;
            float* gp = p.grad_value + o00;
            // the four corner loads of a point are issued together (predicated); corners that are switched off carry
            // v = 0 -> d = 0, and their reductions are predicated off: no branches in the point loop
            float4 v00 = make_float4(0.f, 0.f, 0.f, 0.f), v01 = v00, v10 = v00, v11 = v00;
            if (cd & 1) v00 = ldg4(vp);
            if (cd & 2) v01 = ldg4(vp + vstride);
            if (cd & 4) v10 = ldg4(vp + W * vstride);
            if (cd & 8) v11 = ldg4(vp + (W + 1) * vstride);
            const float d00 = dot4(go, v00), d01 = dot4(go, v01), d10 = dot4(go, v10), d11 = dot4(go, v11);
            const float w00 = (1.f - fy) * (1.f - fx), w01 = (1.f - fy) * fx, w10 = fy * (1.f - fx), w11 = fy * fx;
            red_add4_pred(live && (cd & 1), gp, a * w00, go);
            red_add4_pred(live && (cd & 2), gp + vstride, a * w01, go);
            red_add4_pred(live && (cd & 4), gp + W * vstride, a * w10, go);
            red_add4_pred(live && (cd & 8), gp + (W + 1) * vstride, a * w11, go);
            float ga = w00 * d00 + w01 * d01 + w10 * d10 + w11 * d11;
            float gx = a * ((1.f - fy) * (d01 - d00) + fy * (d11 - d10));
            float gy = a * ((1.f - fx) * (d10 - d00) + fx * (d11 - d01));
            if (!FUSED) { gx *= (float)W; gy *= (float)p.lv.H[l]; }
            ga = group_sum<G>(ga); gx = group_sum<G>(gx); gy = group_sum<G>(gy);
            if (c4 == l) ga_own[s2] = ga;
            gxl[s2] = gx; gyl[s2] = gy;
          }
          if (c4 == l && live) {
            st4(garow, make_float4(gxl[0], gyl[0], gxl[1], gyl[1]));
            st4(garow + 4, make_float4(gxl[2], gyl[2], gxl[3], gyl[3]));
          }
        }
      }
    }
    __syncthreads();                                           // G and W complete

    // ---- P2b: dense levels, the owner lane alone: corner dot products are single words of G ----
    if (owner && c4 >= ld) {
      const float* Grow = Gs + ql * LDW + dense_off;
      float gxl[P], gyl[P];
#pragma unroll
      for (int s2 = 0; s2 < P; ++s2) {
        const int cd = code[s2];
        const int i00 = (cd >> 4) - BIAS;
        const float a = aw[s2], fx = fxv[s2], fy = fyv[s2];
        const float d00 = (cd & 1) ? Grow[i00] : 0.f, d01 = (cd & 2) ? Grow[i00 + 1] : 0.f;
        const float d10 = (cd & 4) ? Grow[i00 + Wl] : 0.f, d11 = (cd & 8) ? Grow[i00 + Wl + 1] : 0.f;
        float gx = a * ((1.f - fy) * (d01 - d00) + fy * (d11 - d10));
        float gy = a * ((1.f - fx) * (d10 - d00) + fx * (d11 - d01));
        if (!FUSED) { gx *= (float)Wl; gy *= (float)Hl; }
        ga_own[s2] = (1.f - fy) * (1.f - fx) * d00 + (1.f - fy) * fx * d01 + fy * (1.f - fx) * d10 + fy * fx * d11;
        gxl[s2] = gx; gyl[s2] = gy;
      }
      if (live) {
        st4(garow, make_float4(gxl[0], gyl[0], gxl[1], gyl[1]));
        st4(garow + 4, make_float4(gxl[2], gyl[2], gxl[3], gyl[3]));
      }
    }

    // ---- softmax backward over the 16 points of the group, gradient of the logits ----
    if (FUSED) {
      float dotp = 0.f;
#pragma unroll
      for (int k = 0; k < P; ++k) dotp += aw[k] * ga_own[k];
      dotp += __shfl_xor_sync(0xffffffffu, dotp, 1);
      dotp += __shfl_xor_sync(0xffffffffu, dotp, 2);
#pragma unroll
      for (int k = 0; k < P; ++k) ga_own[k] = aw[k] * (ga_own[k] - dotp);
    }
    if (live && owner) st4(p.grad_w + bq * p.ldw + m * LP + c4 * P, make_float4(ga_own[0], ga_own[1], ga_own[2], ga_own[3]));

    // ---- P3: acc[px, c] += W[q-tile, px]^T . grad_out[q-tile, c] ----
    if (warp < n_mt) {
#pragma unroll 2
      for (int ks = 0; ks < QT / 8; ++ks) {
        const float* ap = Ws + (ks * 8 + t) * LDW + warp * 16 + g;
        uint32_t ah[4], al[4];
        split_tf32(ap[0], ah[0], al[0]);
        split_tf32(ap[8], ah[1], al[1]);
        split_tf32(ap[4 * LDW], ah[2], al[2]);
        split_tf32(ap[4 * LDW + 8], ah[3], al[3]);
#pragma unroll
        for (int n = 0; n < C::NT3; ++n) {
          const float* bp = gos + (ks * 8 + t) * LDC + n * 8 + g;
          uint32_t bh0, bl0, bh1, bl1;
          split_tf32(bp[0], bh0, bl0);
          split_tf32(bp[4 * LDC], bh1, bl1);
          mma_3x(acc[n], ah, al, bh0, bh1, bl0, bl1);
        }
      }
    }
    __syncthreads();                                           // W and the grad_out tile may be overwritten
  }

  // ---- flush: one 8-byte reduction per accumulator pair ----
  if (warp < n_mt) {
    float* gv = p.grad_value + ((int64_t)b * p.S + pix0) * vstride + m * D;
#pragma unroll
    for (int n = 0; n < C::NT3; ++n) {
      const int px0 = warp * 16 + g, px1 = px0 + 8, d = n * 8 + 2 * t;
      if (px0 < npx)
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(gv + (int64_t)px0 * vstride + d), "f"(acc[n][0]), "f"(acc[n][1]) : "memory");
      if (px1 < npx)
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(gv + (int64_t)px1 * vstride + d), "f"(acc[n][2]), "f"(acc[n][3]) : "memory");
    }
  }
}

// Forward on the same decomposition (sparse levels gathered by the 4-lane groups, dense levels as out = W . value) was built
// and measured twice -- corners from global memory: 148 us, corners from a TMA-staged slab: 151 us -- against 112 us for
// msda_fwd_slab_kernel (profiles/r02c_msda_tile_notes.txt): the 4-lane group with shuffle-broadcast records issues ~1.5x the
// instructions per sampling point of the slab kernel's record / quarter-warp scheme.  Removed; the forward stays on the slab.

}  // namespace tile

// ------------------------------------------------------------------------------------------
// few queries (decoder: Lq = 10): one WARP per (b,q,m)
// ------------------------------------------------------------------------------------------
// The general kernels walk the L*P sampling points serially in each thread; with a handful of queries the
// grid is a fraction of one wave and the launch is pure latency (43 us backward at cfg2).  Here the 32 lanes
// of a warp split one (b,q,m): lane = (corner group cg, 4-channel group c4), cg owns PPL = L*P*G/32 sampling
// points with all four corners, so every lane issues 4*PPL independent 128-bit loads at once and the points
// are folded with log2(32/G) shuffle stages.
template <int LP, int G>
struct WarpMap {
  static constexpr int NCG = 32 / G;          // corner groups per warp
  static constexpr int PPL = LP / NCG;        // sampling points per lane
  static_assert(LP % NCG == 0 && PPL >= 1, "points must divide over the corner groups");
};

template <int NCG, int G>
__device__ __forceinline__ float cg_sum(float v) {      // sum over the corner groups (lanes with equal c4)
#pragma unroll
  for (int o = G; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int NCG, int G>
__device__ __forceinline__ float cg_max(float v) {
#pragma unroll
  for (int o = G; o < 32; o <<= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <int L, int P, int G, bool FUSED, bool BWD>
__global__ void __launch_bounds__(256) msda_warp_kernel(const MsdaArgs p) {
  poet_pdl_entry();
  constexpr int LP = L * P, D = 4 * G;
  using WM = WarpMap<LP, G>;
  constexpr int PPL = WM::PPL, NCG = WM::NCG;
  const int lane = threadIdx.x & 31;
  const int64_t bqm = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (bqm >= (int64_t)p.B * p.Lq * p.M) return;                  // warp-uniform
  const int c4 = lane % G, cg = lane / G;
  const int m = (int)(bqm % p.M);
  const int64_t bq = bqm / p.M;
  const int b = (int)(bq / p.Lq);
  const int vstride = p.M * D;
  const int64_t voff = ((int64_t)b * p.S * p.M + m) * D + c4 * 4;

  // this lane's sampling points: pt = cg * PPL + j
  float a[PPL], lx[PPL], ly[PPL];
#pragma unroll
  for (int j = 0; j < PPL; ++j) {
    const int pt = cg * PPL + j;
    a[j] = __ldg(p.w + bq * p.ldw + m * LP + pt);
    const float2 xy = __ldg(reinterpret_cast<const float2*>(p.a + bq * p.lda + (m * LP + pt) * 2));
    lx[j] = xy.x; ly[j] = xy.y;
  }
  if (FUSED) {                                                   // softmax over the L*P logits of (b,q,m)
    float mx = a[0];
#pragma unroll
    for (int j = 1; j < PPL; ++j) mx = fmaxf(mx, a[j]);
    mx = cg_max<NCG, G>(mx);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < PPL; ++j) { a[j] = __expf(a[j] - mx); sum += a[j]; }
    sum = cg_sum<NCG, G>(sum);
    const float inv = __fdividef(1.f, sum);
#pragma unroll
    for (int j = 0; j < PPL; ++j) a[j] *= inv;
  }
  float4 go = make_float4(0.f, 0.f, 0.f, 0.f), acc = go;
  if (BWD) go = ldg4(p.grad_out + bqm * D + c4 * 4);
  float g_a[PPL], g_x[PPL], g_y[PPL];
#pragma unroll
  for (int j = 0; j < PPL; ++j) {
    const int pt = cg * PPL + j, l = pt / P;
    const int H = p.lv.H[l], W = p.lv.W[l];
    float px = lx[j], py = ly[j];
    if (FUSED) {
      const float2 r = __ldg(reinterpret_cast<const float2*>(p.ref + (bq * L + l) * 2));
      px = r.x + px * p.lv.inv_W[l]; py = r.y + py * p.lv.inv_H[l];
    }
    const float x = px * (float)W - 0.5f, y = py * (float)H - 0.5f;
    g_a[j] = 0.f; g_x[j] = 0.f; g_y[j] = 0.f;
    if (x > -1.f && y > -1.f && x < (float)W && y < (float)H) {
      const float xf = floorf(x), yf = floorf(y);
      const int x0 = (int)xf, y0 = (int)yf;
      const float fx = x - xf, fy = y - yf;
      const bool xl = x0 >= 0, xh = x0 + 1 < W, yl = y0 >= 0, yh = y0 + 1 < H;
      const int64_t i00 = voff + ((int64_t)p.lv.start[l] + y0 * W + x0) * vstride;
      float4 v00 = make_float4(0.f, 0.f, 0.f, 0.f), v01 = v00, v10 = v00, v11 = v00;
      if (yl && xl) v00 = ldg4(p.value + i00);
      if (yl && xh) v01 = ldg4(p.value + i00 + vstride);
      if (yh && xl) v10 = ldg4(p.value + i00 + (int64_t)W * vstride);
      if (yh && xh) v11 = ldg4(p.value + i00 + (int64_t)(W + 1) * vstride);
      const float w00 = (1.f - fy) * (1.f - fx), w01 = (1.f - fy) * fx, w10 = fy * (1.f - fx), w11 = fy * fx;
      if (!BWD) {
        const float aw = a[j];
        acc.x += aw * (w00 * v00.x + w01 * v01.x + w10 * v10.x + w11 * v11.x);
        acc.y += aw * (w00 * v00.y + w01 * v01.y + w10 * v10.y + w11 * v11.y);
        acc.z += aw * (w00 * v00.z + w01 * v01.z + w10 * v10.z + w11 * v11.z);
        acc.w += aw * (w00 * v00.w + w01 * v01.w + w10 * v10.w + w11 * v11.w);
      } else {
        const float aw = a[j];
        if (yl && xl) red_add4(p.grad_value + i00, aw * w00, go);
        if (yl && xh) red_add4(p.grad_value + i00 + vstride, aw * w01, go);
        if (yh && xl) red_add4(p.grad_value + i00 + (int64_t)W * vstride, aw * w10, go);
        if (yh && xh) red_add4(p.grad_value + i00 + (int64_t)(W + 1) * vstride, aw * w11, go);
        const float d00 = dot4(go, v00), d01 = dot4(go, v01), d10 = dot4(go, v10), d11 = dot4(go, v11);
        g_a[j] = w00 * d00 + w01 * d01 + w10 * d10 + w11 * d11;
        g_x[j] = aw * ((1.f - fy) * (d01 - d00) + fy * (d11 - d10));
        g_y[j] = aw * ((1.f - fx) * (d10 - d00) + fx * (d11 - d01));
        if (!FUSED) { g_x[j] *= (float)W; g_y[j] *= (float)H; }
      }
    }
  }
  if (!BWD) {
    acc.x = cg_sum<NCG, G>(acc.x); acc.y = cg_sum<NCG, G>(acc.y);
    acc.z = cg_sum<NCG, G>(acc.z); acc.w = cg_sum<NCG, G>(acc.w);
    if (cg == 0) st4(p.out + bqm * D + c4 * 4, acc);
    return;
  }
  // backward: fold the channel groups (partial dot products), then the softmax backward over all points
  float dot = 0.f;
#pragma unroll
  for (int j = 0; j < PPL; ++j) {
    g_a[j] = group_sum<G>(g_a[j]); g_x[j] = group_sum<G>(g_x[j]); g_y[j] = group_sum<G>(g_y[j]);
    dot += a[j] * g_a[j];
  }
  if (FUSED) {
    dot = cg_sum<NCG, G>(dot);
#pragma unroll
    for (int j = 0; j < PPL; ++j) g_a[j] = a[j] * (g_a[j] - dot);
  }
  if (c4 == 0) {
#pragma unroll
    for (int j = 0; j < PPL; ++j) {
      const int pt = cg * PPL + j;
      p.grad_w[bq * p.ldw + m * LP + pt] = g_a[j];
      *reinterpret_cast<float2*>(p.grad_a + bq * p.lda + (m * LP + pt) * 2) = make_float2(g_x[j], g_y[j]);
    }
  }
}

template <bool BWD>
static int try_warp_kernel(const MsdaArgs& a, int mode, cudaStream_t s) {
  static const int enabled = []() { const char* e = getenv("POET_MSDA_WARP"); return e ? atoi(e) : 1; }();
  const int64_t warps = (int64_t)a.B * a.Lq * a.M;
  if (!enabled || a.L != 4 || a.P != 4 || warps * 32 > (int64_t)POET_NUM_SMS * 2048) return POET_ERR_UNSUPPORTED;
  const int grid = poet_ceil_div(warps * 32, 256);
#define POET_WARP_LAUNCH(GG)                                                                  \
  do {                                                                                        \
    if (mode) poet_launch(msda_warp_kernel<4, 4, GG, true, BWD>, dim3(grid), dim3(256), 0, s, a);                  \
    else poet_launch(msda_warp_kernel<4, 4, GG, false, BWD>, dim3(grid), dim3(256), 0, s, a);                      \
    return poet_launch_status();                                                              \
  } while (0)
  switch (a.D) {
    case 8: POET_WARP_LAUNCH(2);
    case 16: POET_WARP_LAUNCH(4);
    case 32: POET_WARP_LAUNCH(8);
    default: return POET_ERR_UNSUPPORTED;
  }
#undef POET_WARP_LAUNCH
}

// First level (>= 1) from which the remaining levels hold at most 448 pixels together, or L if the dense backward does
// not apply (POET_MSDA_DENSE=0, other head sizes, too few queries to amortise the per-item flush).
static int dense_first_level(const MsdaArgs& a) {
  // Off by default: measured on B200 (profiles/r02_msda_bwd_dense.txt) the pair (scatter kernel with level-0 reductions
  // only + dense kernel) is not faster than the scatter kernel alone, because gathers and reductions share the LSU pipe
  // and the level-0 reductions already cost as much as the gather.  POET_MSDA_DENSE=1 selects it (parity-tested).
  static const int enabled = []() { const char* e = getenv("POET_MSDA_DENSE"); return e ? atoi(e) : 0; }();
  if (!enabled || a.D != 16 || a.L != 4 || a.P != 4 || a.Lq < 256) return a.L;
  for (int l0 = 1; l0 < a.L; ++l0) {
    const int npx = a.S - a.lv.start[l0];
    if (npx <= dense::MAX_MT * 4 * 16) return l0;
  }
  return a.L;
}

static int launch_dense_bwd(const MsdaArgs& a, int mode, int lvl0, cudaStream_t s) {
  const int npx = a.S - a.lv.start[lvl0];
  const int n_mt = poet_ceil_div(npx, 16);
  const size_t smem = (size_t)n_mt * 16 * dense::QS * 4 + 2 * 16 * dense::GS * 4;
  const int slots = 3 * POET_NUM_SMS;                      // three resident CTAs per SM (about 70 KB of shared memory each)
  int best = 1;
  double best_cost = 1e30;
  for (int nc = 1; nc <= 16; ++nc) {                       // query chunks per (image, head): makespan in queries, plus the flush
    const int qpc = poet_ceil_div(a.Lq, nc);
    if (nc > 1 && qpc < 96) break;
    const double rounds = (double)poet_ceil_div((int64_t)a.B * a.M * nc, slots);
    const double cost = rounds * (qpc + 24.0);
    if (cost < best_cost) { best_cost = cost; best = nc; }
  }
  const int nchunk = best, q_per_chunk = poet_ceil_div(a.Lq, nchunk);
  const int n_items = a.B * a.M * nchunk;
  const int grid = n_items < slots ? n_items : slots;
  auto launch = [&](auto kern) -> int {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    poet_launch(kern, dim3(grid), dim3(dense::THREADS), smem, s, a, lvl0, npx, n_mt, nchunk, q_per_chunk, n_items);
    return poet_launch_status();
  };
  return mode ? launch(dense::msda_bwd_dense_kernel<4, 4, true>) : launch(dense::msda_bwd_dense_kernel<4, 4, false>);
}

// Tile backward: POET_OK if it ran, POET_ERR_UNSUPPORTED if the shape does not qualify (the caller falls back to the
// one-thread-per-(b,q,m,channel group) kernels: same results up to summation order).
static int try_tile_bwd(const MsdaArgs& a, int mode, cudaStream_t s) {
  static const int enabled = []() { const char* e = getenv("POET_MSDA_TILE"); return e ? atoi(e) : 1; }();
  if (!enabled || (a.D != 16 && a.D != 32) || a.L != 4 || a.P != 4 || a.Lq < 2 * tile::QT) return POET_ERR_UNSUPPORTED;
  int ld = a.L;                                              // first of the trailing levels that fit the dense tile together
  for (int l0 = 0; l0 < a.L; ++l0)
    if (a.S - a.lv.start[l0] <= tile::NPX) { ld = l0; break; }
  // at least two dense levels: with one (the 1280x960 pyramid, 80 pixels) the tile phases cost more than the quarter of
  // the reductions they remove (cfg5: 2.52 vs 2.12 ms per launch, profiles/r02c_msda_tile_notes.txt)
  static const int max_ld = []() { const char* e = getenv("POET_MSDA_TILE_MAX_LD"); return e ? atoi(e) : 2; }();
  if (ld > max_ld) return POET_ERR_UNSUPPORTED;
  const int npx = a.S - a.lv.start[ld];
  const int n_tiles = poet_ceil_div(a.Lq, a.D == 16 ? tile::Cfg<4>::QT : tile::Cfg<8>::QT);
  const int slots = 3 * POET_NUM_SMS;                        // three resident CTAs per SM
  int best = 1;
  double best_cost = 1e30;
  for (int nc = 1; nc <= n_tiles; ++nc) {                    // chunks per (image, head): makespan in tiles + staging / flush
    const int tpc = poet_ceil_div(n_tiles, nc);
    if ((int64_t)(nc - 1) * tpc >= n_tiles) continue;        // would leave an empty chunk
    const double rounds = (double)poet_ceil_div((int64_t)a.B * a.M * nc, slots);
    const double cost = rounds * (tpc + 0.5);
    if (cost < best_cost) { best_cost = cost; best = nc; }
  }
  static const int force_nc = []() { const char* e = getenv("POET_MSDA_TILE_NCHUNK"); return e ? atoi(e) : 0; }();   // A/B only
  if (force_nc > 0 && force_nc <= n_tiles) best = force_nc;
  const int nchunk = best, tiles_per_chunk = poet_ceil_div(n_tiles, nchunk);
  const int64_t grid = (int64_t)a.B * a.M * nchunk;
  if (grid >= ((int64_t)1 << 31)) return POET_ERR_UNSUPPORTED;
  const size_t smem = (size_t)(a.D == 16 ? tile::Cfg<4>::SMEM_FLOATS : tile::Cfg<8>::SMEM_FLOATS) * sizeof(float);
  auto launch = [&](auto kern) -> int {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    poet_launch(kern, dim3((unsigned)grid), dim3(tile::THREADS), smem, s, a, ld, npx, nchunk, tiles_per_chunk, n_tiles);
    return poet_launch_status();
  };
  if (a.D == 16) return mode ? launch(tile::msda_bwd_tile_kernel<4, true>) : launch(tile::msda_bwd_tile_kernel<4, false>);
  return mode ? launch(tile::msda_bwd_tile_kernel<8, true>) : launch(tile::msda_bwd_tile_kernel<8, false>);
}

int fill_args(MsdaArgs& a, const int32_t* shapes_host, int B, int S, int Lq, int M, int D, int L, int P,
              const float* value, const float* aa, int64_t lda, const float* w, int64_t ldw, const float* ref, int mode) {
  POET_REQUIRE(value && aa && w && shapes_host, POET_ERR_NULL_POINTER);
  POET_REQUIRE(mode == 0 || (mode == 1 && ref != nullptr), POET_ERR_NULL_POINTER);
  POET_REQUIRE(B > 0 && S > 0 && Lq > 0 && M > 0, POET_ERR_BAD_SHAPE);
  POET_REQUIRE(L >= 1 && L <= kMaxLevels && (L * P) % 4 == 0 && L * P <= kMaxLP, POET_ERR_UNSUPPORTED);
  POET_REQUIRE(D == 8 || D == 16 || D == 32 || D == 64, POET_ERR_UNSUPPORTED);   // D/4 lanes must divide a warp
  POET_REQUIRE(lda % 4 == 0 && ldw % 4 == 0 && lda >= (int64_t)M * L * P * 2 && ldw >= (int64_t)M * L * P,
               POET_ERR_BAD_ALIGNMENT);
  POET_REQUIRE(poet_aligned16(value) && poet_aligned16(aa) && poet_aligned16(w), POET_ERR_BAD_ALIGNMENT);
  POET_REQUIRE((int64_t)S * M * D < (int64_t)1 << 31, POET_ERR_BAD_SHAPE);        // per-image offsets are 32-bit
  int start = 0;
  for (int l = 0; l < L; ++l) {
    a.lv.H[l] = shapes_host[2 * l]; a.lv.W[l] = shapes_host[2 * l + 1]; a.lv.start[l] = start;
    POET_REQUIRE(a.lv.H[l] > 0 && a.lv.W[l] > 0, POET_ERR_BAD_SHAPE);
    a.lv.inv_H[l] = 1.0f / (float)a.lv.H[l];
    a.lv.inv_W[l] = 1.0f / (float)a.lv.W[l];
    start += a.lv.H[l] * a.lv.W[l];
  }
  POET_REQUIRE(start == S, POET_ERR_BAD_SHAPE);
  a.value = value; a.a = aa; a.lda = lda; a.w = w; a.ldw = ldw; a.ref = ref;
  a.B = B; a.S = S; a.Lq = Lq; a.M = M; a.D = D; a.L = L; a.P = P;
  a.red_levels = L;
  return POET_OK;
}

template <bool BWD, int LL, int PP, int GG>
void launch_one(const MsdaArgs& a, int mode, int grid, cudaStream_t s) {
  if constexpr (BWD && LL == 4 && PP == 4 && GG >= 4) {     // every PoET config: per-point arithmetic shared by the group
    static const int shared = []() { const char* e = getenv("POET_MSDA_BWD_SHARED"); return e ? atoi(e) : 1; }();
    if (shared) {
      if (mode) poet_launch(msda_bwd_shared_kernel<GG, true>, dim3(grid), dim3(256), 0, s, a);
      else poet_launch(msda_bwd_shared_kernel<GG, false>, dim3(grid), dim3(256), 0, s, a);
      return;
    }
  }
  if (BWD) {
    if (mode) poet_launch(msda_bwd_kernel<LL, PP, GG, true>, dim3(grid), dim3(256), 0, s, a);
    else poet_launch(msda_bwd_kernel<LL, PP, GG, false>, dim3(grid), dim3(256), 0, s, a);
  } else {
    if (mode) poet_launch(msda_fwd_kernel<LL, PP, GG, true>, dim3(grid), dim3(256), 0, s, a);
    else poet_launch(msda_fwd_kernel<LL, PP, GG, false>, dim3(grid), dim3(256), 0, s, a);
  }
}

template <bool BWD, int LL, int PP>
int launch_g(const MsdaArgs& a, int mode, cudaStream_t s) {
  const int64_t total = (int64_t)a.B * a.Lq * a.M * (a.D / 4);
  const int grid = poet_ceil_div(total, 256);
  switch (a.D) {
    case 8: launch_one<BWD, LL, PP, 2>(a, mode, grid, s); break;
    case 16: launch_one<BWD, LL, PP, 4>(a, mode, grid, s); break;
    case 32: launch_one<BWD, LL, PP, 8>(a, mode, grid, s); break;
    case 64: launch_one<BWD, LL, PP, 16>(a, mode, grid, s); break;
    default: return POET_ERR_UNSUPPORTED;
  }
  return poet_launch_status();
}

template <bool BWD>
int dispatch(const MsdaArgs& a, int mode, cudaStream_t s) {
  if (a.L == 4 && a.P == 4) return launch_g<BWD, 4, 4>(a, mode, s);      // every PoET config
  if (a.L == 4 && a.P == 2) return launch_g<BWD, 4, 2>(a, mode, s);
  if (a.L == 2 && a.P == 2) return launch_g<BWD, 2, 2>(a, mode, s);      // upstream test.py shape
  if (a.L == 2 && a.P == 4) return launch_g<BWD, 2, 4>(a, mode, s);
  if (a.L == 1 && a.P == 4) return launch_g<BWD, 1, 4>(a, mode, s);
  return POET_ERR_UNSUPPORTED;
}

}  // namespace

extern "C" int poet_msda_fwd(const float* value, const float* a, int64_t lda, const float* w, int64_t ldw,
                             const float* ref, float* out, const int32_t* shapes_host, int B, int S, int Lq,
                             int M, int D, int L, int P, int mode, poet_stream_t stream) {
  MsdaArgs args{};
  int rc = fill_args(args, shapes_host, B, S, Lq, M, D, L, P, value, a, lda, w, ldw, ref, mode);
  if (rc) return rc;
  POET_REQUIRE(out != nullptr, POET_ERR_NULL_POINTER);
  POET_REQUIRE(poet_aligned16(out), POET_ERR_BAD_ALIGNMENT);
  args.out = out;
  const int rc_warp = try_warp_kernel<false>(args, mode, (cudaStream_t)stream);
  if (rc_warp != POET_ERR_UNSUPPORTED) return rc_warp;
  const int rc_slab = try_slab_fwd(args, mode, (cudaStream_t)stream);
  if (rc_slab != POET_ERR_UNSUPPORTED) return rc_slab;
  return dispatch<false>(args, mode, (cudaStream_t)stream);
}

extern "C" int poet_msda_bwd(const float* value, const float* a, int64_t lda, const float* w, int64_t ldw,
                             const float* ref, const float* grad_out, float* grad_value, float* grad_a,
                             float* grad_w, const int32_t* shapes_host, int B, int S, int Lq, int M, int D,
                             int L, int P, int mode, poet_stream_t stream) {
  MsdaArgs args{};
  int rc = fill_args(args, shapes_host, B, S, Lq, M, D, L, P, value, a, lda, w, ldw, ref, mode);
  if (rc) return rc;
  POET_REQUIRE(grad_out && grad_value && grad_a && grad_w, POET_ERR_NULL_POINTER);
  POET_REQUIRE(poet_aligned16(grad_out) && poet_aligned16(grad_value) && poet_aligned16(grad_a) &&
               poet_aligned16(grad_w), POET_ERR_BAD_ALIGNMENT);
  args.grad_out = grad_out; args.grad_value = grad_value; args.grad_a = grad_a; args.grad_w = grad_w;
  const int rc_warp = try_warp_kernel<true>(args, mode, (cudaStream_t)stream);
  if (rc_warp != POET_ERR_UNSUPPORTED) return rc_warp;
  // many queries (the encoder): one CTA per (image, head, run of query tiles), low-resolution levels as dense products
  const int rc_tile = try_tile_bwd(args, mode, (cudaStream_t)stream);
  if (rc_tile != POET_ERR_UNSUPPORTED) return rc_tile;
  // legacy experiment (POET_MSDA_DENSE=1): grad_value of the low-resolution levels from a separate dense kernel, the
  // scatter kernel keeps its global reductions for the high-resolution level(s) only
  const int lvl0 = dense_first_level(args);
  args.red_levels = lvl0;
  static const int red_dbg = []() { const char* e = getenv("POET_MSDA_RED_LEVELS"); return e ? atoi(e) : -1; }();   // timing bisection only (tools/msda_red_bisect.sh)
  if (red_dbg >= 0) args.red_levels = red_dbg;
  static const int dbg = []() { const char* e = getenv("POET_MSDA_DENSE_DEBUG"); return e ? atoi(e) : 0; }();   // timing bisection: 1 scatter part only, 2 dense part only
  const int rc_scatter = (dbg == 2 && lvl0 < args.L) ? POET_OK : dispatch<true>(args, mode, (cudaStream_t)stream);
  if (rc_scatter != POET_OK || lvl0 >= args.L || dbg == 1) return rc_scatter;
  return launch_dense_bwd(args, mode, lvl0, (cudaStream_t)stream);
}
