// Latency-optimised fp32 GEMM for the query-row contractions of the decoder and the pose heads.
//
// The decoder chain (reference models/deformable_transformer.py:275-292) and the pose heads
// (pose_estimation_transformer.py:357-393, 677-689) are ~40 DEPENDENT GEMMs per decoder layer and direction on
// B*Q = 160 rows: a few MFLOP each.  On the persistent tcgen05 kernel such a launch costs 8-9 us (barrier / TMEM
// set-up, a 2-3 stage operand pipeline that pays the L2 latency once per 64-wide k-block, TMA-store drain) and uses
// 2-8 SMs; the chain is pure latency.  This kernel is built for that regime instead:
//   * 32 x 32 output tiles -> 40-320 CTAs for one GEMM, every operand byte requested up front with cp.async
//     (LDGSTS, zero-fill outside the matrix) into a ring of 128-wide k stages: one L2 round trip, not one per k-block;
//   * 256 threads = 4 k-groups x (8 x 8 threads with 4 x 4 register tiles): the k range of a stage is split over the
//     groups, partial tiles meet in shared memory once at the end;
//   * exact fp32 FFMA (these rows carry the 1e-4 translation budget; no operand splitting needed at this size),
//     weights either as fp32 or as the bf16 hi/lo planes of the step's arena (hi + lo, the same 2^-17 operand);
//   * the epilogues of poet_gemm: alpha, bias, ReLU, ReLU gate, beta = 1 accumulation, bias-gradient column sums.
//   * deep reductions (K >= 512: linear2 forward, linear1 / sampling-offset dgrad) are split over a thread-block
//     CLUSTER of 2-4 CTAs along K: each CTA has its whole k range in flight after one L2 round trip (as in the K = 256
//     case), the partial tiles meet in the first CTA's epilogue through distributed shared memory -- no workspace, no
//     zero-fill, no atomics;
// All four operand layouts (forward NT, dgrad NN, wgrad TN) are template variants of the shared-memory fetch.
#include <cuda_bf16.h>
#include "common.cuh"

namespace small {

constexpr int BM = 32, BN = 32, BKS = 128;          // output tile, k per stage
constexpr int THREADS = 256, KG = 4;                // 4 k-groups of 64 threads
constexpr int KPG = BKS / KG;                       // k per group per stage
constexpr int LDK = BKS + 4;                        // row stride (floats) of a k-contiguous tile [row][k]
constexpr int LDR = BM + 4;                         // row stride (floats) of a row-contiguous tile [k][row]; 36 % 32 == 4
constexpr int LDKH = BKS + 8;                       // bf16 plane tiles: row strides in bf16 elements (16-byte aligned rows,
constexpr int LDRH = BN + 8;                        // conflict-free 8-byte fetches)
constexpr int PLANE_ELEMS = BN * LDKH > BKS * LDRH ? BN * LDKH : BKS * LDRH;   // one plane (hi or lo) of one stage
constexpr int TILE_FLOATS = PLANE_ELEMS;            // = 2 planes x 2 bytes; also >= the fp32 tile (32 x 132 or 128 x 36 floats)
static_assert(TILE_FLOATS >= BM * LDK && TILE_FLOATS >= BKS * LDR, "fp32 tile must fit the per-operand area");
constexpr int MAX_STAGES = 3;

struct Args {
  const float* A; int64_t lda;
  const float* B; const __nv_bfloat16* Bhi; const __nv_bfloat16* Blo; int64_t ldb;
  float* C; int64_t ldc;
  int M, N, K;
  float alpha;
  const float* bias; const float* gate;
  float* a_colsum;
  int flags;
  int stages;
  int ksplit;                                        // CTAs of a cluster (grid z) that share one output tile's reduction
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  const int bytes = valid ? 16 : 0;                  // src-size 0: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One operand tile of one stage.  KC: the operand is k-contiguous in global memory ([rows, K]) -> smem [row][k];
// else it is row-contiguous ([K, rows]) -> smem [k][row].  `base` points at fp32 (4 bytes / element).
template <bool KC>
__device__ __forceinline__ void load_tile_f32(float* s, const float* __restrict__ G, int64_t ld, int row0, int nrows,
                                              int k0, int kend, int tid) {
  if (KC) {                                           // 32 rows x 128 k: 32 chunks of 16 bytes per row
#pragma unroll
    for (int i = 0; i < BM * (BKS / 4) / THREADS; ++i) {
      const int c = tid + i * THREADS;
      const int r = c / (BKS / 4), k = (c % (BKS / 4)) * 4;
      const bool ok = row0 + r < nrows && k0 + k < kend;          // K % 4 == 0: a chunk is entirely inside or outside
      cp_async16(s + r * LDK + k, G + (int64_t)(ok ? row0 + r : 0) * ld + (ok ? k0 + k : 0), ok);
    }
  } else {                                            // 128 k x 32 rows: 8 chunks per k
#pragma unroll
    for (int i = 0; i < BKS * (BM / 4) / THREADS; ++i) {
      const int c = tid + i * THREADS;
      const int k = c / (BM / 4), r = (c % (BM / 4)) * 4;
      const bool ok = k0 + k < kend && row0 + r < nrows;          // rows % 4 == 0 (checked on the host)
      cp_async16(s + k * LDR + r, G + (int64_t)(ok ? k0 + k : 0) * ld + (ok ? row0 + r : 0), ok);
    }
  }
}

// bf16 plane tile (hi or lo): the same logical layout, 2 bytes / element, 8 elements per 16-byte chunk; the two planes
// of a stage share the operand's tile area (PLANE_ELEMS bf16 each).
template <bool KC>
__device__ __forceinline__ void load_tile_bf16(__nv_bfloat16* s, const __nv_bfloat16* __restrict__ G, int64_t ld, int row0,
                                               int nrows, int k0, int kend, int tid) {
  if (KC) {
#pragma unroll
    for (int i = 0; i < BN * (BKS / 8) / THREADS; ++i) {
      const int c = tid + i * THREADS;
      const int r = c / (BKS / 8), k = (c % (BKS / 8)) * 8;
      const bool ok = row0 + r < nrows && k0 + k < kend;          // K % 8 == 0 for plane operands
      cp_async16(s + r * LDKH + k, G + (int64_t)(ok ? row0 + r : 0) * ld + (ok ? k0 + k : 0), ok);
    }
  } else {
    constexpr int CH = BKS * (BN / 8);                // 512 chunks
#pragma unroll
    for (int i = 0; i < CH / THREADS; ++i) {
      const int c = tid + i * THREADS;
      const int k = c / (BN / 8), r = (c % (BN / 8)) * 8;
      const bool ok = k0 + k < kend && row0 + r < nrows;          // rows % 8 == 0 for plane operands
      cp_async16(s + k * LDRH + r, G + (int64_t)(ok ? k0 + k : 0) * ld + (ok ? row0 + r : 0), ok);
    }
  }
}

__device__ __forceinline__ float bf16_to_f32(uint32_t bits16) { return __uint_as_float(bits16 << 16); }

// a[i][kk] / b[j][kk] for this thread's 4 rows / 4 columns at k = kb .. kb+3 of the stage.
// k-contiguous tiles: rows {t + 8 i} (one 128-bit load per row); row-contiguous tiles: rows {4 t + i} (one per k).
template <bool KC>
__device__ __forceinline__ void fetch_f32(const float* s, int t, int kb, float (&v)[4][4]) {
  if (KC) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 x = *reinterpret_cast<const float4*>(s + (t + 8 * i) * LDK + kb);
      v[i][0] = x.x; v[i][1] = x.y; v[i][2] = x.z; v[i][3] = x.w;
    }
  } else {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float4 x = *reinterpret_cast<const float4*>(s + (kb + kk) * LDR + 4 * t);
      v[0][kk] = x.x; v[1][kk] = x.y; v[2][kk] = x.z; v[3][kk] = x.w;
    }
  }
}

template <bool KC>
__device__ __forceinline__ void fetch_planes(const __nv_bfloat16* hi, const __nv_bfloat16* lo, int t, int kb, float (&v)[4][4]) {
  if (KC) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint2 h = *reinterpret_cast<const uint2*>(hi + (t + 8 * i) * LDKH + kb);
      const uint2 l = *reinterpret_cast<const uint2*>(lo + (t + 8 * i) * LDKH + kb);
      v[i][0] = bf16_to_f32(h.x & 0xffffu) + bf16_to_f32(l.x & 0xffffu);
      v[i][1] = bf16_to_f32(h.x >> 16) + bf16_to_f32(l.x >> 16);
      v[i][2] = bf16_to_f32(h.y & 0xffffu) + bf16_to_f32(l.y & 0xffffu);
      v[i][3] = bf16_to_f32(h.y >> 16) + bf16_to_f32(l.y >> 16);
    }
  } else {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const uint2 h = *reinterpret_cast<const uint2*>(hi + (kb + kk) * LDRH + 4 * t);
      const uint2 l = *reinterpret_cast<const uint2*>(lo + (kb + kk) * LDRH + 4 * t);
      v[0][kk] = bf16_to_f32(h.x & 0xffffu) + bf16_to_f32(l.x & 0xffffu);
      v[1][kk] = bf16_to_f32(h.x >> 16) + bf16_to_f32(l.x >> 16);
      v[2][kk] = bf16_to_f32(h.y & 0xffffu) + bf16_to_f32(l.y & 0xffffu);
      v[3][kk] = bf16_to_f32(h.y >> 16) + bf16_to_f32(l.y >> 16);
    }
  }
}

template <bool AK, bool BK_, bool PLANES>
__global__ void __launch_bounds__(THREADS) gemm_small_kernel(const Args p) {
  poet_pdl_launch_dependents();
  extern __shared__ __align__(16) float sm[];        // stages x {A tile, B tile (fp32, or bf16 hi | lo in its two halves)}
  const int tid = threadIdx.x;
  const int kg = tid >> 6, t64 = tid & 63, tx = t64 & 7, ty = t64 >> 3;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int n_stage_all = (p.K + BKS - 1) / BKS;
  // cluster split-K: CTA z of the cluster owns stages [s_beg, s_end) of the reduction (ksplit == 1: all of them)
  const int spp = (n_stage_all + p.ksplit - 1) / p.ksplit;
  const int s_beg = (int)blockIdx.z * spp, n_stage = min(n_stage_all, s_beg + spp);
  const bool do_colsum = !AK && p.a_colsum != nullptr && blockIdx.x == 0;

  auto stage_ptr = [&](int s, int which) { return sm + ((size_t)(s % p.stages) * 2 + which) * TILE_FLOATS; };
  auto issue_b = [&](int s) {
    if (s >= n_stage) return;
    const int k0 = s * BKS;
    if (PLANES) {
      __nv_bfloat16* bt = reinterpret_cast<__nv_bfloat16*>(stage_ptr(s, 1));
      load_tile_bf16<BK_>(bt, p.Bhi, p.ldb, n0, p.N, k0, p.K, tid);
      load_tile_bf16<BK_>(bt + PLANE_ELEMS, p.Blo, p.ldb, n0, p.N, k0, p.K, tid);        // second half of the tile area
    } else {
      load_tile_f32<BK_>(stage_ptr(s, 1), p.B, p.ldb, n0, p.N, k0, p.K, tid);
    }
  };
  auto issue = [&](int s, bool with_b) {
    if (s < n_stage) {
      load_tile_f32<AK>(stage_ptr(s, 0), p.A, p.lda, m0, p.M, s * BKS, p.K, tid);
      if (with_b) issue_b(s);
    }
    cp_async_commit();                                // one group per stage slot, empty when past the end
  };
  // A parameter operand (POET_GEMM_B_STABLE) is requested before the wait for the preceding kernel: its L2 round trip
  // overlaps that kernel's tail.  The early requests join the first committed group, which stage 0 waits for anyway.
  const bool b_early = (p.flags & POET_GEMM_B_STABLE) != 0;
  if (b_early)
    for (int s = s_beg; s < s_beg + p.stages; ++s) issue_b(s);    // every slot of the ring is free at this point
  poet_pdl_wait();

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float csum = 0.f;

  for (int s = s_beg; s < s_beg + p.stages - 1; ++s) issue(s, !b_early);    // prologue: every stage of a short K is in flight at once
  for (int s = s_beg; s < n_stage; ++s) {
    issue(s + p.stages - 1, !(b_early && s == s_beg));
    if (p.stages == 3) cp_async_wait<2>(); else if (p.stages == 2) cp_async_wait<1>(); else cp_async_wait<0>();
    __syncthreads();
    const float* as = stage_ptr(s, 0);
    const float* bs = stage_ptr(s, 1);
    const int klim = min(BKS, p.K - s * BKS);          // zero-filled beyond: no guard needed in the math
    if (kg * KPG < klim) {
#pragma unroll 2
      for (int kb = kg * KPG; kb < kg * KPG + KPG; kb += 4) {
        float a[4][4], b[4][4];
        fetch_f32<AK>(as, ty, kb, a);
        if (PLANES) {
          const __nv_bfloat16* bh = reinterpret_cast<const __nv_bfloat16*>(bs);
          fetch_planes<BK_>(bh, bh + PLANE_ELEMS, tx, kb, b);
        } else {
          fetch_f32<BK_>(bs, tx, kb, b);
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i][kk], b[j][kk], acc[i][j]);
      }
    }
    if (do_colsum && tid < BM) {                      // bias gradient: column sums of the dY tile ([k][m] layout), tile column 0 only
      for (int k = 0; k < klim; ++k) csum += as[k * LDR + tid];
    }
    __syncthreads();                                  // the slot is refilled by the next iteration's issue()
  }
  cp_async_wait<0>();

  // ---- fold the four k-groups, then the epilogue on [32 x 32] with one float4 per thread ----
  float* red = sm;                                    // 4 x 32 x 33 floats, the stage memory is free now
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = AK ? ty + 8 * i : 4 * ty + i, n = BK_ ? tx + 8 * j : 4 * tx + j;
      red[(kg * BM + m) * (BN + 1) + n] = acc[i][j];
    }
  __syncthreads();
  const int m = tid >> 3, nq = (tid & 7) * 4;
  const int gm = m0 + m, gn = n0 + nq;
  if (do_colsum && tid < BM && m0 + tid < p.M) atomicAdd(p.a_colsum + m0 + tid, csum);
  float v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float x = 0.f;
#pragma unroll
    for (int g = 0; g < KG; ++g) x += red[(g * BM + m) * (BN + 1) + nq + j];
    v[j] = x;
  }
  if (p.ksplit > 1) {
    // the partial tiles of the cluster's CTAs meet in CTA 0: each CTA publishes its folded tile in its own shared memory
    // (behind the k-group partials), CTA 0 reads the others' through the distributed-shared-memory window
    float* part = sm + KG * BM * (BN + 1);             // [32][36] floats
    *reinterpret_cast<float4*>(part + m * 36 + nq) = make_float4(v[0], v[1], v[2], v[3]);
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (blockIdx.z == 0) {
      const uint32_t local = (uint32_t)__cvta_generic_to_shared(part + m * 36 + nq);
      for (int r = 1; r < p.ksplit; ++r) {
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(r));
        float4 t;
        asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"(remote));
        v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
      }
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");   // remote tiles stay alive until read
    if (blockIdx.z != 0) return;
  }
  if (gm >= p.M || gn >= p.N) return;
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] *= p.alpha;
  const bool relu = p.flags & POET_GEMM_RELU, accum = p.flags & POET_GEMM_ACCUMULATE;
  float* cp = p.C + (int64_t)gm * p.ldc + gn;
  const float* gp = p.gate ? p.gate + (int64_t)gm * p.ldc + gn : nullptr;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (gn + j >= p.N) continue;
    float x = v[j];
    if (p.bias) x += __ldg(p.bias + gn + j);
    if (relu) x = fmaxf(x, 0.f);
    if (gp) x = (__ldg(gp + j) > 0.f) ? x : 0.f;
    v[j] = x;
  }
  if (accum) {                                        // beta = 1 through atomics (micro-batches / side streams share the gradient)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (gn + j < p.N) atomicAdd(cp + j, v[j]);
  } else if (gn + 3 < p.N && (p.ldc & 3) == 0) {
    st4(cp, make_float4(v[0], v[1], v[2], v[3]));
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (gn + j < p.N) cp[j] = v[j];
  }
}

template <bool AK, bool BK_, bool PLANES>
static int launch(const Args& a, cudaStream_t s) {
  size_t smem = (size_t)a.stages * 2 * TILE_FLOATS * sizeof(float);
  const size_t epi = (size_t)(KG * BM * (BN + 1) + BM * 36) * sizeof(float);      // k-group partials + the cluster's tile
  if (smem < epi) smem = epi;
  auto kern = gemm_small_kernel<AK, BK_, PLANES>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const dim3 grid(poet_ceil_div(a.N, BN), poet_ceil_div(a.M, BM), a.ksplit);
  if (a.ksplit == 1) {
    poet_launch(kern, grid, dim3(THREADS), smem, s, a);
    return poet_launch_status();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = poet_pdl_enabled() ? 1 : 0;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = 1; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = a.ksplit;
  cfg.attrs = attr; cfg.numAttrs = 2;
  (void)cudaLaunchKernelEx(&cfg, kern, a);
  return poet_launch_status();
}

}  // namespace small

// Shapes this kernel is meant for: few output tiles (the tcgen05 kernel would run on a handful of SMs and be pure
// latency), operands addressable in 16-byte chunks.  M = output rows, N = output columns, K = reduction length.
bool poet_gemm_small_supported(int M, int N, int K, int a_kcontig, int b_kcontig, int64_t lda, int64_t ldb, int64_t ldc,
                               const void* A, const void* B, const void* Bhi, const void* Blo, bool relu_or_gate) {
  static const int enabled = []() { const char* e = getenv("POET_GEMM_SMALL"); return e ? atoi(e) : 1; }();
  if (!enabled) return false;
  (void)ldc; (void)relu_or_gate;
  const int64_t tiles = (int64_t)poet_ceil_div(M, small::BM) * poet_ceil_div(N, small::BN);
  if (tiles > 2048 || (int64_t)M * N * K > ((int64_t)1 << 28)) return false;       // big problems belong to the tensor cores
  if (K % 4 != 0 || lda % 4 != 0 || !poet_aligned16(A)) return false;
  if (!a_kcontig && M % 4 != 0) return false;
  const bool planes = B == nullptr;
  if (planes) {
    if (!Bhi || !Blo || K % 8 != 0 || ldb % 8 != 0 || !poet_aligned16(Bhi) || !poet_aligned16(Blo)) return false;
    if (!b_kcontig && N % 8 != 0) return false;
  } else {
    if (ldb % 4 != 0 || !poet_aligned16(B)) return false;
    if (!b_kcontig && N % 4 != 0) return false;
  }
  return true;
}

int poet_gemm_small(const float* A, int64_t lda, int a_kcontig, const float* Bm, const void* b_hi, const void* b_lo,
                    int64_t ldb, int b_kcontig, float* C, int64_t ldc, int M, int N, int K, float alpha, const float* bias,
                    const float* gate, float* a_colsum, int flags, cudaStream_t s) {
  small::Args a;
  a.A = A; a.lda = lda; a.B = Bm; a.Bhi = reinterpret_cast<const __nv_bfloat16*>(b_hi);
  a.Blo = reinterpret_cast<const __nv_bfloat16*>(b_lo); a.ldb = ldb; a.C = C; a.ldc = ldc; a.M = M; a.N = N; a.K = K;
  a.alpha = alpha; a.bias = bias; a.gate = gate; a.a_colsum = a_colsum; a.flags = flags;
  const int n_stage_all = poet_ceil_div(K, small::BKS);
  // deep reductions over few output tiles: a cluster of CTAs along K (POET_GEMM_SMALL_KSPLIT=0 disables)
  static const int ksplit_on = []() { const char* e = getenv("POET_GEMM_SMALL_KSPLIT"); return e ? atoi(e) : 1; }();
  const int64_t tiles = (int64_t)poet_ceil_div(M, small::BM) * poet_ceil_div(N, small::BN);
  a.ksplit = 1;
  static const int ksplit_min = []() { const char* e = getenv("POET_GEMM_SMALL_KSPLIT_MIN"); return e ? atoi(e) : 4; }();   // stages; A/B only
  if (ksplit_on && n_stage_all >= ksplit_min && n_stage_all >= 2 && tiles <= 96 && a_colsum == nullptr && !(flags & POET_GEMM_ACCUMULATE))
    a.ksplit = n_stage_all >= 8 ? 4 : (n_stage_all >= 6 ? 3 : 2);
  const int n_stage = poet_ceil_div(n_stage_all, a.ksplit);
  a.stages = n_stage < small::MAX_STAGES ? (n_stage < 1 ? 1 : n_stage) : small::MAX_STAGES;
  const bool planes = Bm == nullptr;
#define POET_SMALL(AKV, BKV)                                                                         \
  return planes ? small::launch<AKV, BKV, true>(a, s) : small::launch<AKV, BKV, false>(a, s)
  if (a_kcontig && b_kcontig) { POET_SMALL(true, true); }
  if (a_kcontig && !b_kcontig) { POET_SMALL(true, false); }
  if (!a_kcontig && b_kcontig) { POET_SMALL(false, true); }
  POET_SMALL(false, false);
#undef POET_SMALL
}
