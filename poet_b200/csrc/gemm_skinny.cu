// Exact-fp32 GEMM for the launch-latency-bound contractions of the decoder / heads (B*Q = 160 rows)
// and their weight gradients (K = 160): a few MFLOP each, ~170 of them per training step.
//
// A tensor-core pipeline (TMEM allocation, tensor maps, mbarrier ring) costs more than these
// problems take; what matters here is a short dependency chain.  Each CTA owns a 16 x 32 output tile,
// streams K in chunks of 128 with every global load of a chunk issued before the first use (one
// memory round trip per chunk), keeps the operands k-major in shared memory (A read by broadcast,
// B read conflict-free) and finishes with the same fused epilogue as the big kernels
// (bias / ReLU / ReLU-gate / row mask / accumulate).  All three layouts (forward NT, dgrad NN,
// wgrad TN) without transposes in HBM.
#include "common.cuh"

namespace {

constexpr int TM = 16, TN = 32, KC = 128, THREADS = 256;

struct SkinnyArgs {
  const float* A; int64_t lda;
  const float* B; int64_t ldb;
  float* C; int64_t ldc;
  int M, N, K;
  float alpha;
  const float* bias; const float* gate; const uint8_t* row_mask;
  int flags;
};

template <bool AK, bool BK_>
__global__ void __launch_bounds__(THREADS) gemm_skinny_kernel(const SkinnyArgs p) {
  // k-major tiles; B rows padded so that both the transposing scalar stores (BK_) and the float4
  // stores (!BK_) stay cheap and the compute-phase reads Bs[k][lane] are conflict-free
  constexpr int BP = BK_ ? TN + 1 : TN + 4;
  __shared__ __align__(16) float As[KC][TM];
  __shared__ __align__(16) float Bs[KC][BP];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  float acc0 = 0.f, acc1 = 0.f;                        // rows m0 + w and m0 + w + 8, column n0 + lane

  for (int k0 = 0; k0 < p.K; k0 += KC) {
    const int kc = min(KC, p.K - k0);
    // ---- stage A [TM x kc] ----
    if (AK) {                                           // A[m][k], k contiguous: float4 along k, transposed store
#pragma unroll
      for (int idx = tid; idx < TM * (KC / 4); idx += THREADS) {
        const int i = idx / (KC / 4), k = (idx % (KC / 4)) * 4;
        const int m = m0 + i;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < p.M && k < kc) {
          const float* g = p.A + (int64_t)m * p.lda + k0 + k;
          if (k + 3 < kc && ((reinterpret_cast<uintptr_t>(g) & 15) == 0)) v = ldg4(g);
          else {
            v.x = __ldg(g);
            if (k + 1 < kc) v.y = __ldg(g + 1);
            if (k + 2 < kc) v.z = __ldg(g + 2);
            if (k + 3 < kc) v.w = __ldg(g + 3);
          }
        }
        As[k][i] = v.x; As[k + 1][i] = v.y; As[k + 2][i] = v.z; As[k + 3][i] = v.w;
      }
    } else {                                            // A stored [K][M], m contiguous
#pragma unroll
      for (int idx = tid; idx < KC * TM; idx += THREADS) {
        const int k = idx / TM, i = idx % TM;
        const int m = m0 + i;
        As[k][i] = (k < kc && m < p.M) ? __ldg(p.A + (int64_t)(k0 + k) * p.lda + m) : 0.f;
      }
    }
    // ---- stage B [TN x kc] ----
    if (BK_) {                                          // B[n][k] (nn.Linear weight), k contiguous
#pragma unroll
      for (int idx = tid; idx < TN * (KC / 4); idx += THREADS) {
        const int j = idx / (KC / 4), k = (idx % (KC / 4)) * 4;
        const int n = n0 + j;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < p.N && k < kc) {
          const float* g = p.B + (int64_t)n * p.ldb + k0 + k;
          if (k + 3 < kc && ((reinterpret_cast<uintptr_t>(g) & 15) == 0)) v = ldg4(g);
          else {
            v.x = __ldg(g);
            if (k + 1 < kc) v.y = __ldg(g + 1);
            if (k + 2 < kc) v.z = __ldg(g + 2);
            if (k + 3 < kc) v.w = __ldg(g + 3);
          }
        }
        Bs[k][j] = v.x; Bs[k + 1][j] = v.y; Bs[k + 2][j] = v.z; Bs[k + 3][j] = v.w;
      }
    } else {                                            // B stored [K][N], n contiguous
#pragma unroll
      for (int idx = tid; idx < KC * (TN / 4); idx += THREADS) {
        const int k = idx / (TN / 4), j = (idx % (TN / 4)) * 4;
        const int n = n0 + j;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < kc && n < p.N) {
          const float* g = p.B + (int64_t)(k0 + k) * p.ldb + n;
          if (n + 3 < p.N && ((reinterpret_cast<uintptr_t>(g) & 15) == 0)) v = ldg4(g);
          else {
            v.x = __ldg(g);
            if (n + 1 < p.N) v.y = __ldg(g + 1);
            if (n + 2 < p.N) v.z = __ldg(g + 2);
            if (n + 3 < p.N) v.w = __ldg(g + 3);
          }
        }
        *reinterpret_cast<float4*>(&Bs[k][j]) = v;
      }
    }
    __syncthreads();
#pragma unroll 16
    for (int k = 0; k < KC; ++k) {                      // rows beyond kc are zero-filled
      const float b = Bs[k][lane];
      acc0 = fmaf(As[k][w], b, acc0);
      acc1 = fmaf(As[k][w + 8], b, acc1);
    }
    __syncthreads();
  }

  const int n = n0 + lane;
  if (n >= p.N) return;
  const float bias = p.bias ? __ldg(p.bias + n) : 0.f;
  const bool relu = p.flags & POET_GEMM_RELU, accum = p.flags & POET_GEMM_ACCUMULATE;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int m = m0 + w + 8 * h;
    if (m >= p.M) continue;
    float x = p.alpha * (h ? acc1 : acc0) + bias;
    if (relu) x = fmaxf(x, 0.f);
    float* cp = p.C + (int64_t)m * p.ldc + n;
    if (p.gate && !(__ldg(p.gate + (int64_t)m * p.ldc + n) > 0.f)) x = 0.f;
    if (p.row_mask && p.row_mask[m]) x = 0.f;
    *cp = accum ? *cp + x : x;
  }
}

}  // namespace

int poet_gemm_skinny(const float* A, int64_t lda, int a_kcontig, const float* Bm, int64_t ldb, int b_kcontig, float* C,
                     int64_t ldc, int M, int N, int K, float alpha, const float* bias, const float* gate,
                     const uint8_t* row_mask, int flags, cudaStream_t s) {
  SkinnyArgs a;
  a.A = A; a.lda = lda; a.B = Bm; a.ldb = ldb; a.C = C; a.ldc = ldc; a.M = M; a.N = N; a.K = K; a.alpha = alpha;
  a.bias = bias; a.gate = gate; a.row_mask = row_mask; a.flags = flags;
  dim3 grid(poet_ceil_div(N, TN), poet_ceil_div(M, TM));
  if (a_kcontig && b_kcontig) gemm_skinny_kernel<true, true><<<grid, THREADS, 0, s>>>(a);
  else if (a_kcontig && !b_kcontig) gemm_skinny_kernel<true, false><<<grid, THREADS, 0, s>>>(a);
  else if (!a_kcontig && b_kcontig) gemm_skinny_kernel<false, true><<<grid, THREADS, 0, s>>>(a);
  else gemm_skinny_kernel<false, false><<<grid, THREADS, 0, s>>>(a);
  return poet_launch_status();
}
