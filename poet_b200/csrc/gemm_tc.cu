// tcgen05 (5th-gen tensor core) GEMM on fp32 data with split-bf16 operands.
//
//   C[M,N] = epi( alpha * op(A)[M,K] . op(B)[K,N] ),  fp32 in / fp32 out, accumulators in TMEM.
//
// Why split precision: the reference computes every nn.Linear in fp32 and the parity budget is
// 1e-4 abs on the translation head; plain TF32/BF16 operands miss it (1.9e-4 / 1.8e-3).
// Each fp32 operand x is split exactly into hi = bf16(x), lo = bf16(x - hi) and the product is
// accumulated as hi*lo + lo*hi + hi*hi in fp32 (POET_GEMM_BF16X3, relative error ~2^-17 per product,
// fp32-grade after accumulation); POET_GEMM_BF16 issues only hi*hi (throughput mode).
//
// One 128 x BN output tile per CTA, BK = 64, 2 smem stages, warp-specialised:
//   warps [0,PW)  A producers (and B producers when B is an fp32 activation): 128-bit global loads
//                 (the next k-block is prefetched into registers while the current one is converted),
//                 split/convert, st.shared into the UMMA canonical SWIZZLE_128B layout (K-major or
//                 MN-major: forward / dgrad / wgrad need no transposes), fence.proxy.async, arrive on
//                 full[stage].  Afterwards the epilogue: tcgen05.ld 32 columns at a time (one accumulator
//                 row per thread), bias / ReLU / ReLU-gate / row-mask / accumulate / split-K reduction.
//   warp PW       TMEM alloc + single-thread tcgen05.mma issue; tcgen05.commit releases smem stages.
//   warp PW+1     TMA: when B is a weight it was split to bf16 hi/lo planes once per step
//                 (poet_split_bf16) and is fetched by cp.async.bulk.tensor straight into the swizzled
//                 stage (complete_tx on the same full[stage] barrier): no per-CTA re-conversion of W.
#include <cuda.h>
#include <cuda_bf16.h>
#include "common.cuh"

namespace tc {

constexpr int BM = 128, BK = 64, STAGES = 2;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}

// ---- TMA ----------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

// ---- tcgen05 ------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row (lane base + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- descriptors --------------------------------------------------------------------------
// shared-memory matrix descriptor, SWIZZLE_128B, sm_100 version bit set (cute/arch/mma_sm100_desc.hpp)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;          // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;          // LayoutType::SWIZZLE_128B
  return d;
}
// instruction descriptor, kind::f16: D=f32, A=B=bf16, M=128, N=BN, majorness per operand
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

struct Args {
  const float* A; int64_t lda;
  const float* B; int64_t ldb;            // fp32 B (nullptr when B arrives through the tensor maps)
  float* C; int64_t ldc;
  int M, N, K;
  float alpha;
  const float* bias; const float* gate; const uint8_t* row_mask;
  int flags;
  int kb_per_split;       // k-blocks (of 64) per grid.z slice
  int splits;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// A tile of ROWS "rows" x (SEGS*64) contiguous fp32 elements <-> SEGS blocks of [ROWS x 128 B] bf16 in
// the SWIZZLE_128B canonical layout.
//   K-major operand : row = m (or n) index, contiguous = k   -> ROWS = tile extent, SEGS = 1
//   MN-major operand: row = k index,        contiguous = m/n -> ROWS = 64,          SEGS = extent/64
// Each thread owns CH 16-byte chunks (8 elements); chunk id = tid + i*NT.
template <int ROWS, int SEGS, int NT>
struct Tile {
  static constexpr int CHUNKS = ROWS * SEGS * 8;
  static constexpr int CH = CHUNKS / NT;
  static_assert(CHUNKS % NT == 0, "tile / thread mismatch");

  __device__ static __forceinline__ void load(const float* __restrict__ G, int64_t ld, int row0, int row_end, int col0,
                                              int col_end, int tid, float4 (&v)[CH][2]) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int ch = tid + i * NT;
      const int c = ch & 7, rs = ch >> 3;
      const int row = rs % ROWS, seg = rs / ROWS;
      const int grow = row0 + row, gcol = col0 + seg * 64 + c * 8;
      v[i][0] = make_float4(0.f, 0.f, 0.f, 0.f);
      v[i][1] = v[i][0];
      if (grow < row_end && gcol < col_end) {                    // col_end % 8 == 0 (checked on the host)
        const float* p = G + (int64_t)grow * ld + gcol;
        v[i][0] = ldg4(p);
        v[i][1] = ldg4(p + 4);
      }
    }
  }

  template <bool WITH_LO>
  __device__ static __forceinline__ void store(uint8_t* s_hi, uint8_t* s_lo, int tid, const float4 (&v)[CH][2]) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int ch = tid + i * NT;
      const int c = ch & 7, rs = ch >> 3;
      const int row = rs % ROWS, seg = rs / ROWS;
      const uint32_t off = (uint32_t)seg * (ROWS * 128) + (uint32_t)(row >> 3) * 1024 + (uint32_t)(row & 7) * 128 +
                           (uint32_t)((c ^ (row & 7)) << 4);
      const float x[8] = {v[i][0].x, v[i][0].y, v[i][0].z, v[i][0].w, v[i][1].x, v[i][1].y, v[i][1].z, v[i][1].w};
      float h[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) h[j] = __bfloat162float(__float2bfloat16_rn(x[j]));
      uint4 hi;
      hi.x = pack_bf16(h[0], h[1]); hi.y = pack_bf16(h[2], h[3]); hi.z = pack_bf16(h[4], h[5]); hi.w = pack_bf16(h[6], h[7]);
      *reinterpret_cast<uint4*>(s_hi + off) = hi;
      if (WITH_LO) {
        uint4 lo;
        lo.x = pack_bf16(x[0] - h[0], x[1] - h[1]); lo.y = pack_bf16(x[2] - h[2], x[3] - h[3]);
        lo.z = pack_bf16(x[4] - h[4], x[5] - h[5]); lo.w = pack_bf16(x[6] - h[6], x[7] - h[7]);
        *reinterpret_cast<uint4*>(s_lo + off) = lo;
      }
    }
  }
};

template <int BN, bool A_MN, bool B_MN, bool X3, bool B_TMA, int PW>
__global__ void __launch_bounds__((PW + 2) * 32, 1)
gemm_tc_kernel(const Args p, const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo) {
  constexpr int NT = PW * 32;                                        // producer / epilogue threads
  constexpr int PLANES = X3 ? 2 : 1;
  constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;         // one bf16 plane
  constexpr int STAGE_BYTES = (A_BYTES + B_BYTES) * PLANES;
  using TA = Tile<A_MN ? BK : BM, A_MN ? BM / 64 : 1, NT>;
  using TB = Tile<B_MN ? BK : BN, B_MN ? BN / 64 : 1, NT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bars[2 * STAGES + 1];
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int total_kb = (p.K + BK - 1) / BK;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_end = min(total_kb, kb_begin + p.kb_per_split);
  const int nkb = kb_end - kb_begin;

  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[STAGES]), accum_bar = smem_u32(&bars[2 * STAGES]);
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, NT + (B_TMA ? 1 : 0)); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == PW) tmem_alloc(smem_u32(&s_tmem), BN);
  if (B_TMA && warp == PW + 1 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_hi) : "memory");
    if (X3) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_lo) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;

  auto a_hi_of = [&](int s) { return smem + s * STAGE_BYTES; };
  auto b_hi_of = [&](int s) { return smem + s * STAGE_BYTES + A_BYTES * PLANES; };
  auto loadA = [&](int i, float4 (&v)[TA::CH][2]) {
    const int k0 = (kb_begin + i) * BK;
    if (A_MN) TA::load(p.A, p.lda, k0, p.K, m0, p.M, tid, v);
    else      TA::load(p.A, p.lda, m0, p.M, k0, p.K, tid, v);
  };
  auto publish = [&](int i, const float4 (&va)[TA::CH][2]) {
    const int s = i % STAGES;
    if constexpr (B_TMA) {
      if (i >= STAGES) mbar_wait(empty0 + 8 * s, ((i / STAGES) - 1) & 1);
      TA::template store<X3>(a_hi_of(s), a_hi_of(s) + A_BYTES, tid, va);
    } else {                                                  // fp32 activation B: loads issued before the stage wait
      float4 vb[TB::CH][2];
      const int k0 = (kb_begin + i) * BK;
      if (B_MN) TB::load(p.B, p.ldb, k0, p.K, n0, p.N, tid, vb);
      else      TB::load(p.B, p.ldb, n0, p.N, k0, p.K, tid, vb);
      if (i >= STAGES) mbar_wait(empty0 + 8 * s, ((i / STAGES) - 1) & 1);
      TA::template store<X3>(a_hi_of(s), a_hi_of(s) + A_BYTES, tid, va);
      TB::template store<X3>(b_hi_of(s), b_hi_of(s) + B_BYTES, tid, vb);
    }
    fence_proxy_async();                                      // generic-proxy smem writes -> visible to the tensor core
    mbar_arrive(full0 + 8 * s);
  };

  if (warp < PW) {
    // ===================== producers =====================
    float4 ra[TA::CH][2], rb[TA::CH][2];
    if (nkb > 0) loadA(0, ra);
    for (int i = 0; i < nkb; i += 2) {
      if (i + 1 < nkb) loadA(i + 1, rb);                      // prefetch the next k-block into registers
      publish(i, ra);
      if (i + 1 < nkb) {
        if (i + 2 < nkb) loadA(i + 2, ra);
        publish(i + 1, rb);
      }
    }
  } else if (warp == PW) {
    if (lane == 0) {
      // ===================== MMA issuer (one thread) =====================
      constexpr uint32_t idesc = make_idesc(BN, A_MN, B_MN);
      // K-major: 8-row groups 1024 B apart, a 16-wide k-step is 32 B inside the 128 B swizzle row.
      // MN-major: 64-element m/n groups ROWS*128 = 8192 B apart (LBO), 8-row k groups 1024 B apart (SBO),
      //           a 16-wide k-step is two k groups = 2048 B.
      constexpr uint32_t A_LBO = A_MN ? BK * 128 : 16, A_STEP = A_MN ? 2048 : 32;
      constexpr uint32_t B_LBO = B_MN ? BK * 128 : 16, B_STEP = B_MN ? 2048 : 32;
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        mbar_wait(full0 + 8 * s, (i / STAGES) & 1);
        tc_fence_after();
        const uint32_t st = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t a_hi = st, a_lo = st + A_BYTES, b_hi = st + A_BYTES * PLANES, b_lo = b_hi + B_BYTES;
#pragma unroll
        for (int j = 0; j < BK / 16; ++j) {
          const uint64_t dah = make_desc(a_hi + j * A_STEP, A_LBO, 1024);
          const uint64_t dbh = make_desc(b_hi + j * B_STEP, B_LBO, 1024);
          const uint32_t first = (i | j) ? 1u : 0u;
          if (X3) {
            const uint64_t dal = make_desc(a_lo + j * A_STEP, A_LBO, 1024);
            const uint64_t dbl = make_desc(b_lo + j * B_STEP, B_LBO, 1024);
            umma_bf16(tmem_base, dah, dbl, idesc, first);        // small cross terms first
            umma_bf16(tmem_base, dal, dbh, idesc, 1u);
            umma_bf16(tmem_base, dah, dbh, idesc, 1u);
          } else {
            umma_bf16(tmem_base, dah, dbh, idesc, first);
          }
        }
        umma_commit(empty0 + 8 * s);                             // stage reusable once these MMAs retire
      }
      umma_commit(accum_bar);                                    // accumulator complete
    }
  } else if (B_TMA && lane == 0) {
    // ===================== TMA: pre-split bf16 weight planes =====================
    for (int i = 0; i < nkb; ++i) {
      const int s = i % STAGES;
      if (i >= STAGES) mbar_wait(empty0 + 8 * s, ((i / STAGES) - 1) & 1);
      const uint32_t bar = full0 + 8 * s;
      mbar_arrive_expect_tx(bar, B_BYTES * PLANES);
      const int k0 = (kb_begin + i) * BK;
      const uint32_t b_hi = smem_u32(b_hi_of(s)), b_lo = b_hi + B_BYTES;
      if (B_MN) {
#pragma unroll
        for (int g = 0; g < BN / 64; ++g) {                        // box = 64 (n, contiguous) x 64 (k rows)
          tma_load_2d(b_hi + g * (BK * 128), &tm_hi, n0 + g * 64, k0, bar);
          if (X3) tma_load_2d(b_lo + g * (BK * 128), &tm_lo, n0 + g * 64, k0, bar);
        }
      } else {                                                     // box = 64 (k, contiguous) x BN (n rows)
        tma_load_2d(b_hi, &tm_hi, k0, n0, bar);
        if (X3) tma_load_2d(b_lo, &tm_lo, k0, n0, bar);
      }
    }
  }

  // ===================== epilogue (producer warps) =====================
  if (warp < PW) {
    if (nkb > 0) {
      mbar_wait(accum_bar, 0);
      tc_fence_after();
    }
    // TMEM hands each thread one accumulator ROW (32 consecutive columns per tcgen05.ld).  Storing that
    // directly would make every warp store touch 32 different rows; instead each warp transposes its
    // 32x32 block through a private padded smem tile (the operand stages are free once accum_bar fired)
    // so that lane = column and every global load / store / reduction is one coalesced 128-byte row.
    constexpr int GROUPS = PW / 4, COLS = BN / GROUPS;
    const int quarter = warp & 3, group = warp >> 2;
    const bool relu = p.flags & POET_GEMM_RELU, accum = p.flags & POET_GEMM_ACCUMULATE;
    float* tile = reinterpret_cast<float*>(smem) + warp * (32 * 33);
    const int mrow0 = m0 + quarter * 32;
    // rows of this warp's quarter that the padding mask zeroes (bit r = row mrow0 + r)
    const uint32_t dead_rows = __ballot_sync(0xffffffffu, p.row_mask != nullptr && mrow0 + lane < p.M &&
                                                              p.row_mask[min(mrow0 + lane, p.M - 1)] != 0);
#pragma unroll 1
    for (int cc = 0; cc < COLS; cc += 32) {
      const int col = group * COLS + cc;
      float v[32];
      if (nkb > 0) tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)col, v);
      else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 32; ++j) tile[lane * 33 + j] = v[j];
      __syncwarp();
      const int n = n0 + col + lane;                           // this lane's output column
      if (n >= p.N) continue;
      const float bias = (p.bias != nullptr && (p.splits == 1 || blockIdx.z == 0)) ? __ldg(p.bias + n) : 0.f;
      const int rows = min(32, p.M - mrow0);
      float* cbase = p.C + (int64_t)mrow0 * p.ldc + n;
      const float alpha = p.alpha;
      if (p.splits > 1) {
#pragma unroll 8
        for (int r = 0; r < rows; ++r)
          asm volatile("red.global.add.f32 [%0], %1;" ::"l"(cbase + (int64_t)r * p.ldc), "f"(alpha * tile[r * 33 + lane] + bias) : "memory");
      } else if (p.gate == nullptr && !accum) {
        // common case: bias (+ReLU) (+row mask); nothing to load, stores are fire-and-forget
#pragma unroll 8
        for (int r = 0; r < rows; ++r) {
          float x = alpha * tile[r * 33 + lane] + bias;
          if (relu) x = fmaxf(x, 0.f);
          if ((dead_rows >> r) & 1u) x = 0.f;
          cbase[(int64_t)r * p.ldc] = x;
        }
      } else {
        // ReLU-gate (dgrad through the FFN activation) and/or accumulate: batches of 16 rows so the
        // dependent global loads overlap
        const float* gbase = p.gate ? p.gate + (int64_t)mrow0 * p.ldc + n : nullptr;
#pragma unroll 1
        for (int r0 = 0; r0 < rows; r0 += 16) {
          float g[16], old[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int r = min(r0 + i, rows - 1);
            g[i] = gbase ? __ldg(gbase + (int64_t)r * p.ldc) : 1.f;
            old[i] = accum ? cbase[(int64_t)r * p.ldc] : 0.f;
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int r = r0 + i;
            if (r < rows) {
              float x = alpha * tile[r * 33 + lane] + bias;
              if (relu) x = fmaxf(x, 0.f);
              if (!(g[i] > 0.f) || ((dead_rows >> r) & 1u)) x = 0.f;
              cbase[(int64_t)r * p.ldc] = x + old[i];
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == PW) tmem_dealloc(tmem_base, BN);
}

// ---- fp32 -> bf16 hi/lo planes (weights, once per step) --------------------------------------
__global__ void __launch_bounds__(256) split_bf16_kernel(const float4* __restrict__ src, uint2* __restrict__ hi,
                                                         uint2* __restrict__ lo, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 x = src[i];
    const float h0 = __bfloat162float(__float2bfloat16_rn(x.x)), h1 = __bfloat162float(__float2bfloat16_rn(x.y));
    const float h2 = __bfloat162float(__float2bfloat16_rn(x.z)), h3 = __bfloat162float(__float2bfloat16_rn(x.w));
    hi[i] = make_uint2(pack_bf16(h0, h1), pack_bf16(h2, h3));
    if (lo) lo[i] = make_uint2(pack_bf16(x.x - h0, x.y - h1), pack_bf16(x.z - h2, x.w - h3));
  }
}

// ---- host side ------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// 2-D bf16 tensor map: `inner` contiguous elements per row, `rows` rows of stride ld elements,
// box = 64 x box_rows, SWIZZLE_128B (the box row is exactly one 128-byte swizzle span).
static int make_map(CUtensorMap* map, const void* base, int64_t inner, int64_t rows, int64_t ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return POET_ERR_UNSUPPORTED;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? POET_OK : POET_ERR_UNSUPPORTED;
}

template <int BN, bool A_MN, bool B_MN, bool X3, bool B_TMA, int PW>
int launch(const Args& a, const CUtensorMap& mh, const CUtensorMap& ml, cudaStream_t s) {
  constexpr int PLANES = X3 ? 2 : 1;
  constexpr size_t stage_smem = (size_t)STAGES * (BM * BK * 2 + BN * BK * 2) * PLANES;
  constexpr size_t epi_smem = (size_t)PW * 32 * 33 * sizeof(float);          // per-warp transpose tiles
  constexpr size_t smem = (stage_smem > epi_smem ? stage_smem : epi_smem) + 1024;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, X3, B_TMA, PW>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  dim3 grid(a.N / BN, poet_ceil_div(a.M, BM), a.splits);
  kern<<<grid, (PW + 2) * 32, smem, s>>>(a, mh, ml);
  return poet_launch_status();
}

template <int BN, bool X3>
int dispatch(const Args& a, bool a_mn, bool b_mn, bool b_tma, const CUtensorMap& mh, const CUtensorMap& ml, cudaStream_t s) {
  if (b_tma) {                                      // B is a pre-split weight: A is k-contiguous (forward / dgrad)
    if (a_mn) return POET_ERR_UNSUPPORTED;
    return b_mn ? launch<BN, false, true, X3, true, 8>(a, mh, ml, s) : launch<BN, false, false, X3, true, 8>(a, mh, ml, s);
  }
  if (!a_mn && !b_mn) return launch<BN, false, false, X3, false, 16>(a, mh, ml, s);
  if (!a_mn && b_mn) return launch<BN, false, true, X3, false, 16>(a, mh, ml, s);
  if (a_mn && b_mn) return launch<BN, true, true, X3, false, 16>(a, mh, ml, s);
  return launch<BN, true, false, X3, false, 16>(a, mh, ml, s);
}

}  // namespace tc

bool poet_gemm_tc_supported(int M, int N, int K, int a_kcontig, int b_kcontig, int64_t lda, int64_t ldb, int64_t ldc) {
  if (N % 128 != 0 || K % 8 != 0 || M % 8 != 0) return false;
  if (lda % 4 != 0 || ldb % 4 != 0 || ldc % 4 != 0) return false;
  // very small problems (head outputs, K < 64) stay on the exact-fp32 SIMT kernel
  if (M < 64 || K < 64) return false;
  (void)a_kcontig; (void)b_kcontig;
  return true;
}

size_t poet_gemm_tc_workspace_bytes(int, int, int, int, int, int) { return 0; }

// b_hi / b_lo != nullptr: B was pre-split into bf16 planes (same logical layout / ldb as the fp32 B).
int poet_gemm_tc(const float* A, int64_t lda, int a_kcontig, const float* Bm, const void* b_hi, const void* b_lo,
                 int64_t ldb, int b_kcontig, float* C, int64_t ldc, int M, int N, int K, float alpha, const float* bias,
                 const float* gate, const uint8_t* row_mask, int flags, int precision, cudaStream_t s) {
  POET_REQUIRE(poet_aligned16(A) && poet_aligned16(C), POET_ERR_BAD_ALIGNMENT);
  POET_REQUIRE(!bias || poet_aligned16(bias), POET_ERR_BAD_ALIGNMENT);
  POET_REQUIRE(!gate || poet_aligned16(gate), POET_ERR_BAD_ALIGNMENT);
  const bool x3 = precision == POET_GEMM_BF16X3;
  const bool b_tma = b_hi != nullptr && (!x3 || b_lo != nullptr) && a_kcontig && (ldb % 8 == 0) &&
                     poet_aligned16(b_hi) && (!x3 || poet_aligned16(b_lo));
  POET_REQUIRE(b_tma || (Bm != nullptr && poet_aligned16(Bm)), POET_ERR_NULL_POINTER);
  tc::Args a;
  a.A = A; a.lda = lda; a.B = b_tma ? nullptr : Bm; a.ldb = ldb; a.C = C; a.ldc = ldc; a.M = M; a.N = N; a.K = K;
  a.alpha = alpha; a.bias = bias; a.gate = gate; a.row_mask = row_mask; a.flags = flags;
  // 128 x 256 tiles when that still fills the machine, else 128 x 128 (decoder rows: more CTAs in flight)
  int bn = (N % 256 == 0) ? 256 : 128;
  if (bn == 256 && (int64_t)(N / 256) * poet_ceil_div(M, tc::BM) < POET_NUM_SMS / 2) bn = 128;
  const int64_t tiles = (int64_t)(N / bn) * poet_ceil_div(M, tc::BM);
  const int total_kb = poet_ceil_div(K, tc::BK);
  int splits = 1;
  const bool linear_epi = !(flags & POET_GEMM_RELU) && gate == nullptr && row_mask == nullptr;
  if (linear_epi && !a_kcontig && tiles < POET_NUM_SMS && total_kb >= 8) {       // weight-gradient shape
    splits = (int)(POET_NUM_SMS / tiles);
    if (splits > total_kb / 4) splits = total_kb / 4;
    if (splits < 1) splits = 1;
  }
  a.kb_per_split = poet_ceil_div(total_kb, splits);
  a.splits = poet_ceil_div(total_kb, a.kb_per_split);
  if (a.splits > 1 && !(flags & POET_GEMM_ACCUMULATE)) {
    cudaError_t e = cudaMemset2DAsync(C, ldc * sizeof(float), 0, (size_t)N * sizeof(float), M, s);
    if (e != cudaSuccess) return (int)e;
  }
  const bool a_mn = !a_kcontig, b_mn = !b_kcontig;
  CUtensorMap mh, ml;
  memset(&mh, 0, sizeof(mh));
  memset(&ml, 0, sizeof(ml));
  if (b_tma) {
    // K-major weight [N,K]: inner = K, rows = N, box rows = BN.  MN-major weight [K,N]: inner = N, rows = K, box rows = 64.
    const int64_t inner = b_mn ? N : K, rows = b_mn ? K : N;
    const int box_rows = b_mn ? tc::BK : bn;
    int rc = tc::make_map(&mh, b_hi, inner, rows, ldb, box_rows);
    if (rc) return rc;
    if (x3) { rc = tc::make_map(&ml, b_lo, inner, rows, ldb, box_rows); if (rc) return rc; }
  }
  if (bn == 256) return x3 ? tc::dispatch<256, true>(a, a_mn, b_mn, b_tma, mh, ml, s) : tc::dispatch<256, false>(a, a_mn, b_mn, b_tma, mh, ml, s);
  return x3 ? tc::dispatch<128, true>(a, a_mn, b_mn, b_tma, mh, ml, s) : tc::dispatch<128, false>(a, a_mn, b_mn, b_tma, mh, ml, s);
}

int poet_split_bf16_impl(const float* src, void* hi, void* lo, int64_t n, cudaStream_t s) {
  POET_REQUIRE(src && hi, POET_ERR_NULL_POINTER);
  POET_REQUIRE(n > 0 && n % 4 == 0, POET_ERR_BAD_SHAPE);
  POET_REQUIRE(poet_aligned16(src) && (reinterpret_cast<uintptr_t>(hi) % 8 == 0) &&
               (!lo || reinterpret_cast<uintptr_t>(lo) % 8 == 0), POET_ERR_BAD_ALIGNMENT);
  int grid = poet_ceil_div(n / 4, 256);
  if (grid > POET_NUM_SMS * 8) grid = POET_NUM_SMS * 8;
  tc::split_bf16_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<uint2*>(hi),
                                             reinterpret_cast<uint2*>(lo), n / 4);
  return poet_launch_status();
}
