// FP32 SIMT GEMM with fused epilogues: the exact-fp32 contraction path of poet_gemm().
// Used for the small (Q-row) decoder / head GEMMs, for ragged shapes and as the bit-faithful
// fp32 yardstick of the tcgen05 split-bf16 path (gemm_tc.cu).
//
// C[M,N] = epi(alpha * op(A) . op(B)); operands may be k-contiguous or m/n-contiguous so the same
// kernel serves nn.Linear forward (NT), dgrad (NN) and wgrad (TN, split-K with atomics).
#include "common.cuh"

namespace {

struct GemmArgs {
  const float* A; int64_t lda;
  const float* B; int64_t ldb;
  float* C; int64_t ldc;
  int M, N, K;
  float alpha;
  const float* bias; const float* gate; const uint8_t* row_mask;
  int flags;
  int k_per_split;
  int splits;
  int vecA, vecB, vecC;
};

// Load a ROWS x BK operand tile into registers (zero-filled outside the matrix / k-range).
template <int ROWS, int BK, int NT, bool KCONTIG>
__device__ __forceinline__ void tile_load(const float* __restrict__ G, int64_t ld, int row0, int nrows,
                                          int k0, int kend, bool vec, float4 (&reg)[ROWS * BK / 4 / NT]) {
  constexpr int NV = ROWS * BK / 4 / NT;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int f = threadIdx.x + i * NT;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (KCONTIG) {
      int r = f / (BK / 4), k = k0 + (f % (BK / 4)) * 4;
      int row = row0 + r;
      if (row < nrows) {
        const float* p = G + (int64_t)row * ld + k;
        if (vec && k + 3 < kend) v = ldg4(p);
        else {
          if (k + 0 < kend) v.x = __ldg(p + 0);
          if (k + 1 < kend) v.y = __ldg(p + 1);
          if (k + 2 < kend) v.z = __ldg(p + 2);
          if (k + 3 < kend) v.w = __ldg(p + 3);
        }
      }
    } else {
      int k = k0 + f / (ROWS / 4), r = row0 + (f % (ROWS / 4)) * 4;
      if (k < kend) {
        const float* p = G + (int64_t)k * ld + r;
        if (vec && r + 3 < nrows) v = ldg4(p);
        else {
          if (r + 0 < nrows) v.x = __ldg(p + 0);
          if (r + 1 < nrows) v.y = __ldg(p + 1);
          if (r + 2 < nrows) v.z = __ldg(p + 2);
          if (r + 3 < nrows) v.w = __ldg(p + 3);
        }
      }
    }
    reg[i] = v;
  }
}

template <int ROWS, int BK, int NT, bool KCONTIG>
__device__ __forceinline__ void tile_store(float (*sm)[ROWS + 4], const float4 (&reg)[ROWS * BK / 4 / NT]) {
  constexpr int NV = ROWS * BK / 4 / NT;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int f = threadIdx.x + i * NT;
    if (KCONTIG) {
      int r = f / (BK / 4), k = (f % (BK / 4)) * 4;
      sm[k + 0][r] = reg[i].x; sm[k + 1][r] = reg[i].y; sm[k + 2][r] = reg[i].z; sm[k + 3][r] = reg[i].w;
    } else {
      int k = f / (ROWS / 4), r = (f % (ROWS / 4)) * 4;
      *reinterpret_cast<float4*>(&sm[k][r]) = reg[i];
    }
  }
}

template <int BM, int BN, int BK, int TM, int TN, bool AK, bool BKC>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
sgemm_kernel(const GemmArgs p) {
  poet_pdl_entry();
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int TX = BN / TN;
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * p.k_per_split;
  const int kend = min(p.K, kbeg + p.k_per_split);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra[BM * BK / 4 / NT], rb[BN * BK / 4 / NT];
  tile_load<BM, BK, NT, AK>(p.A, p.lda, m0, p.M, kbeg, kend, p.vecA, ra);
  tile_load<BN, BK, NT, BKC>(p.B, p.ldb, n0, p.N, kbeg, kend, p.vecB, rb);
  tile_store<BM, BK, NT, AK>(As[0], ra);
  tile_store<BN, BK, NT, BKC>(Bs[0], rb);
  __syncthreads();

  int buf = 0;
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    const bool more = k0 + BK < kend;
    if (more) {
      tile_load<BM, BK, NT, AK>(p.A, p.lda, m0, p.M, k0 + BK, kend, p.vecA, ra);
      tile_load<BN, BK, NT, BKC>(p.B, p.ldb, n0, p.N, k0 + BK, kend, p.vecB, rb);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int c = 0; c < TM / 4; ++c) {
        float4 v = *reinterpret_cast<const float4*>(&As[buf][k][c * (BM / (TM / 4)) + ty * 4]);
        a[c * 4 + 0] = v.x; a[c * 4 + 1] = v.y; a[c * 4 + 2] = v.z; a[c * 4 + 3] = v.w;
      }
#pragma unroll
      for (int c = 0; c < TN / 4; ++c) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][c * (BN / (TN / 4)) + tx * 4]);
        b[c * 4 + 0] = v.x; b[c * 4 + 1] = v.y; b[c * 4 + 2] = v.z; b[c * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      tile_store<BM, BK, NT, AK>(As[buf ^ 1], ra);
      tile_store<BN, BK, NT, BKC>(Bs[buf ^ 1], rb);
      __syncthreads();
      buf ^= 1;
    }
  }

  // ---- epilogue ----
  const bool relu = p.flags & POET_GEMM_RELU;
  const bool accum = p.flags & POET_GEMM_ACCUMULATE;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + (i / 4) * (BM / (TM / 4)) + ty * 4 + (i % 4);
    if (m >= p.M) continue;
    const bool dead = p.row_mask != nullptr && p.row_mask[m] != 0;
#pragma unroll
    for (int c = 0; c < TN / 4; ++c) {
      const int n = n0 + c * (BN / (TN / 4)) + tx * 4;
      if (n >= p.N) continue;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = p.alpha * acc[i][c * 4 + j];
      float* cp = p.C + (int64_t)m * p.ldc + n;
      if (p.splits > 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < p.N) {
            float x = v[j];
            if (p.bias != nullptr && blockIdx.z == 0) x += __ldg(p.bias + n + j);
            atomicAdd(cp + j, x);
          }
        continue;
      }
      const float* gp = p.gate ? p.gate + (int64_t)m * p.ldc + n : nullptr;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (n + j >= p.N) continue;
        float x = v[j];
        if (p.bias) x += __ldg(p.bias + n + j);
        if (relu) x = fmaxf(x, 0.f);
        if (gp) x = (__ldg(gp + j) > 0.f) ? x : 0.f;
        if (dead) x = 0.f;
        v[j] = x;
      }
      if (accum) {
        // beta = 1 goes through atomics: micro-batches on different streams accumulate into the same gradient
        // (one writer per element inside a launch, so a lone launch stays bitwise deterministic)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < p.N) atomicAdd(cp + j, v[j]);
      } else if (p.vecC && n + 3 < p.N) {
        st4(cp, make_float4(v[0], v[1], v[2], v[3]));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < p.N) cp[j] = v[j];
      }
    }
  }
}

template <int BM, int BN, int BK, int TM, int TN>
void launch_cfg(const GemmArgs& a, int a_k, int b_k, cudaStream_t s) {
  dim3 grid(poet_ceil_div(a.N, BN), poet_ceil_div(a.M, BM), a.splits);
  dim3 block((BM / TM) * (BN / TN));
  if (a_k && b_k) poet_launch(sgemm_kernel<BM, BN, BK, TM, TN, true, true>, dim3(grid), dim3(block), 0, s, a);
  else if (a_k && !b_k) poet_launch(sgemm_kernel<BM, BN, BK, TM, TN, true, false>, dim3(grid), dim3(block), 0, s, a);
  else if (!a_k && b_k) poet_launch(sgemm_kernel<BM, BN, BK, TM, TN, false, true>, dim3(grid), dim3(block), 0, s, a);
  else poet_launch(sgemm_kernel<BM, BN, BK, TM, TN, false, false>, dim3(grid), dim3(block), 0, s, a);
}

}  // namespace

// Entry used by poet_gemm() (api.cu) for precision == POET_GEMM_FP32.
int poet_gemm_simt(const float* A, int64_t lda, int a_kcontig, const float* Bm, int64_t ldb, int b_kcontig,
                   float* C, int64_t ldc, int M, int N, int K, float alpha, const float* bias,
                   const float* gate, const uint8_t* row_mask, int flags, cudaStream_t s) {
  GemmArgs a;
  a.A = A; a.lda = lda; a.B = Bm; a.ldb = ldb; a.C = C; a.ldc = ldc;
  a.M = M; a.N = N; a.K = K; a.alpha = alpha; a.bias = bias; a.gate = gate; a.row_mask = row_mask;
  a.flags = flags;
  a.vecA = poet_aligned16(A) && (lda % 4 == 0);
  a.vecB = poet_aligned16(Bm) && (ldb % 4 == 0);
  a.vecC = poet_aligned16(C) && (ldc % 4 == 0) && (gate == nullptr || poet_aligned16(gate));

  const bool big = (int64_t)M * N >= (int64_t)128 * 128 * POET_NUM_SMS / 2;
  const int bm = big ? 128 : 64, bn = big ? 128 : 64;
  const int64_t tiles = (int64_t)poet_ceil_div(M, bm) * poet_ceil_div(N, bn);
  int splits = 1;
  const bool linear_epi = !(flags & POET_GEMM_RELU) && gate == nullptr && row_mask == nullptr;
  // split-K (atomic accumulation) only for the weight-gradient shape (A = dY^T): forward and dgrad
  // stay bitwise deterministic
  if (linear_epi && !a_kcontig && tiles < POET_NUM_SMS && K >= 512) {
    splits = (int)((2 * POET_NUM_SMS + tiles - 1) / tiles);
    int max_splits = K / 128;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  int kps = poet_ceil_div(K, splits);
  kps = (kps + 15) / 16 * 16;
  splits = poet_ceil_div(K, kps);
  a.k_per_split = kps;
  a.splits = splits;
  if (splits > 1 && !(flags & POET_GEMM_ACCUMULATE)) {
    cudaError_t e = cudaMemset2DAsync(C, ldc * sizeof(float), 0, (size_t)N * sizeof(float), M, s);
    if (e != cudaSuccess) return (int)e;
  }
  if (big) launch_cfg<128, 128, 16, 8, 8>(a, a_kcontig, b_kcontig, s);
  else launch_cfg<64, 64, 16, 4, 4>(a, a_kcontig, b_kcontig, s);
  return poet_launch_status();
}
