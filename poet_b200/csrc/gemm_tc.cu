// tcgen05 (5th-gen tensor core) GEMM on fp32 data with split-bf16 operands: persistent,
// warp-specialised, TMEM double-buffered.
//
//   C[M,N] = epi( alpha * op(A)[M,K] . op(B)[K,N] ),  fp32 in / fp32 out, accumulators in TMEM.
//
// Why split precision: the reference computes every nn.Linear in fp32 and the parity budget is
// 1e-4 abs on the translation head; plain TF32/BF16 operands miss it (measured: bf16 single pass
// gives 3e-3..7e-3 on the translation head).  Each fp32 operand x is split exactly into
// hi = bf16(x), lo = bf16(x - hi) and the product is accumulated as hi*lo + lo*hi + hi*hi in fp32
// (POET_GEMM_BF16X3: ~2^-17 per product, fp32-grade after accumulation, measured <= 3e-5 rel. vs fp64);
// POET_GEMM_BF16 issues only hi*hi (throughput mode).
//
// One CTA per SM walks a static round-robin list of (128 x BN output tile, k-split) work items:
//   warps [0,PW)    producers: 128-bit global loads of the fp32 A tile (and of B when it is an fp32
//                   activation: wgrad), next k-block prefetched into registers, split/convert, st.shared
//                   into the UMMA canonical SWIZZLE_128B layout (K-major or MN-major: forward / dgrad /
//                   wgrad need no transposes), fence.proxy.async, arrive on full[stage].  The smem ring
//                   runs across work items, so the next tile's operands stream in during this tile's epilogue.
//   warp PW         single-thread tcgen05.mma issue into TMEM accumulator buffer (tile & 1);
//                   tcgen05.commit frees smem stages and publishes finished accumulators.
//   warp PW+1       TMA: weights are split to bf16 hi/lo planes once per step (poet_split_bf16) and
//                   fetched by cp.async.bulk.tensor straight into the swizzled stage.
//   warps 12..19    epilogue (two per TMEM lane quarter, alternating 32-column chunks): tcgen05.ld (one accumulator
//                   row per thread), bias / ReLU (+ sign bitmask out) / ReLU-gate (bitmask or fp32) / row mask in
//                   registers, st.shared into a per-warp SWIZZLE_128B 32x32 staging box, then ONE asynchronous TMA
//                   store per box (cp.async.bulk.tensor; cp.reduce...add for beta=1 and split-K).  Runs concurrently
//                   with the next tile's MMAs (other TMEM buffer).
// Weight-gradient shape (both operands fp32 activations, MN-major): BK = 32, both operand tiles are
// register-prefetched one k-block ahead, lines further ahead are pulled into L2 with prefetch hints,
// and the tile is 128 x 256 when N allows (less L2->SM traffic per flop).
#include <cstdlib>
#include <cuda.h>
#include <cuda_bf16.h>
#include "common.cuh"
#include "tma_host.cuh"

// Pipeline-bisection knobs (tools/gemm_bisect.py, profiles/r02_gemm_epilogue_bisect.txt) exist only in builds with
// -DPOET_GEMM_BISECT: the kernel is issue bound next to the MMA stream, and even never-taken `p.debug & bit` tests in
// its loops cost 0.15 ms per cfg2 step (7.62 -> 7.47 ms when three of them were removed).
#ifdef POET_GEMM_BISECT
#define POET_DBG(p, bit) ((p).debug & (bit))
#else
#define POET_DBG(p, bit) 0
#endif

namespace tc {

constexpr int BM = 128;
constexpr int EPI_WARPS = 8;            // two per TMEM lane quarter: they take alternate 32-column chunks of a tile

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

// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
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
  uint32_t* relu_bits;                    // [M, N/32] sign bitmask of the ReLU output (TMA epilogue only)
  const uint32_t* gate_bits;              // [M, N/32] keep-mask for the ReLU backward (TMA epilogue only)
  const uint8_t* a_row_mask;              // rows of the stored A operand (tokens) to read as zeros, or nullptr
  float* a_colsum;                        // weight-gradient shape only: a_colsum[m] += sum_k A[k, m] (the bias gradient), or nullptr
  PoetDropout drop;                       // train-mode dropout of the (activated) output, TMA epilogue only; seed == nullptr: off
  int flags;
  int kb_per_split;       // k-blocks (of BK) per split
  int splits;
  int n_tiles, total_work, total_kb;
  uint64_t r_splits, r_n_tiles, r_tail_slices;   // fast_div reciprocals
  int tail_shift;                                // log2(tail_slices)
  int fast_k;                                    // K % BK == 0: operand loads need no bounds predicates (Tile::load<true>)
  int tail_start, tail_slices;   // work ids >= tail_start are column slices of the last round's tiles (tail_start = total_work: none)
  int grid_limit;         // persistent CTAs at most (POET_NUM_SMS, or fewer for POET_GEMM_BACKGROUND launches)
  int l2_prefetch;        // 1: L2 prefetch hints ahead of the register-prefetched operand loads
  int debug;              // POET_GEMM_DEBUG bit flags (pipeline bisection only): 1 no A loads, 2 no A stores, 4 no TMA, 8 no epilogue stores,
                          // 16 no epilogue math (TMA epilogue), 32 no MMA issue (barriers only), 64 no staging-reuse wait, 128 no epilogue proxy fence,
                          // 256 no TMA store issue, 512 no staging writes (64..512: timing only, results invalid; the skeleton bits 1024..4096 of
                          // profiles/r02_gemm_epilogue_bisect.txt were removed again after the measurement)
};

// work item w -> (m0, n0, split, k-block range); n fastest so concurrent CTAs share A rows in L2
// Tail splitting: with T tiles on G = 148 persistent CTAs the last, partial round leaves most SMs idle for a whole
// tile time ([25600 x 256] outputs: 200 tiles = 1.35 rounds).  The T % G tiles of that round are therefore cut
// into tail_slices column slices of bn / tail_slices columns each (work ids >= tail_start), so the last round is a
// full one of a fraction of the duration.  A slice is the same work item with a narrower MMA N / epilogue width.
struct Work {
  int m0, n0, split, kb0, nkb, bn;
};
// n / d for 0 <= n, d < 2^20 by one 64-bit multiply with the host-computed reciprocal floor(2^40 / d) + 1 (exact while
// n * d < 2^40).  decode() is inlined in every role's loop: four run-time integer divisions each were ~100 instructions.
__device__ __forceinline__ int fast_div(int n, uint64_t recip) { return (int)(((uint64_t)(uint32_t)n * recip) >> 40); }
template <int BK>
__device__ __forceinline__ Work decode(const Args& p, int w, int bn) {
  Work k;
  const int total_kb = p.total_kb;
  if (w >= p.tail_start) {                                   // only set up when splits == 1
    const int t = w - p.tail_start;
    const int q = fast_div(t, p.r_tail_slices);
    const int tile = p.tail_start + q, sl = t - q * p.tail_slices;
    k.bn = bn >> p.tail_shift;
    const int mt = fast_div(tile, p.r_n_tiles);
    k.n0 = (tile - mt * p.n_tiles) * bn + sl * k.bn;
    k.m0 = mt * BM;
    k.split = 0; k.kb0 = 0; k.nkb = total_kb;
    return k;
  }
  const int tile = fast_div(w, p.r_splits);
  k.split = w - tile * p.splits;
  k.bn = bn;
  const int mt = fast_div(tile, p.r_n_tiles);
  k.n0 = (tile - mt * p.n_tiles) * bn;
  k.m0 = mt * BM;
  k.kb0 = k.split * p.kb_per_split;
  k.nkb = min(total_kb, k.kb0 + p.kb_per_split) - k.kb0;
  return k;
}

__device__ __forceinline__ void sts128(uint32_t saddr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// L2 prefetch hints (no architectural effect on results)
__device__ __forceinline__ void l2_prefetch_line(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {   // p 16-byte aligned, bytes % 16 == 0
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// TMA stores of one [32 x 32] fp32 box (smem: 32 rows x 128 B, SWIZZLE_128B)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// A tile of ROWS "rows" x (SEGS*64) contiguous fp32 elements <-> SEGS blocks of [ROWS x 128 B] bf16 in
// the SWIZZLE_128B canonical layout.
//   K-major operand : row = m (or n) index, contiguous = k   -> ROWS = tile extent, SEGS = 1 (BK = 64)
//   MN-major operand: row = k index,        contiguous = m/n -> ROWS = BK,          SEGS = extent/64
// Each thread owns CH 16-byte chunks (8 elements); chunk id = tid + i*NT.
template <int ROWS, int SEGS, int NT>
struct Tile {
  static constexpr int CHUNKS = ROWS * SEGS * 8;
  static constexpr int CH = CHUNKS / NT;
  static_assert(CHUNKS % NT == 0, "tile / thread mismatch");

  // row_mask (nullable): stored rows of the operand with row_mask[row] != 0 are read as zeros (the masked_fill
  // backward of MSDeformAttn.value_proj, applied on the fly instead of by a separate pass over the gradient)
  // FAST: the caller guarantees that no element that must read as zero lies outside the matrix (the reduction length is
  // a multiple of BK); rows / columns past the end are clamped onto the last valid chunk (they only feed output rows /
  // columns that the epilogue's tensor map clips), so the loads carry no predicates and no zero initialisation.
  template <bool FAST = false>
  __device__ static __forceinline__ void load(const float* __restrict__ G, int64_t ld, int row0, int row_end, int col0,
                                              int col_end, int tid, float4 (&v)[CH][2], const uint8_t* __restrict__ row_mask = nullptr) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int ch = tid + i * NT;
      const int c = ch & 7, rs = ch >> 3;
      const int row = rs % ROWS, seg = rs / ROWS;
      const int grow = row0 + row, gcol = col0 + seg * 64 + c * 8;
      if (FAST) {
        const float* p = G + (int64_t)min(grow, row_end - 1) * ld + min(gcol, col_end - 8);
        v[i][0] = ldg4(p);
        v[i][1] = ldg4(p + 4);
        continue;
      }
      v[i][0] = make_float4(0.f, 0.f, 0.f, 0.f);
      v[i][1] = v[i][0];
      if (grow < row_end && gcol < col_end) {                    // col_end % 8 == 0 (checked on the host)
        const float* p = G + (int64_t)grow * ld + gcol;
        v[i][0] = ldg4(p);
        v[i][1] = ldg4(p + 4);
      }
    }
    if (row_mask != nullptr) {                                   // after the data loads are in flight: no added latency
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const int rs = (tid + i * NT) >> 3;
        const int grow = row0 + rs % ROWS;
        if (grow < row_end && __ldg(row_mask + grow) != 0) { v[i][0] = make_float4(0.f, 0.f, 0.f, 0.f); v[i][1] = v[i][0]; }
      }
    }
  }

  // one L2 prefetch hint per 128-byte line of the tile (LINES = ROWS*SEGS*2 <= 2*NT)
  __device__ static __forceinline__ void prefetch(const float* __restrict__ G, int64_t ld, int row0, int row_end, int col0,
                                                  int col_end, int tid) {
    constexpr int LINES = ROWS * SEGS * 2;
#pragma unroll
    for (int i = 0; i < (LINES + NT - 1) / NT; ++i) {
      const int ln = tid + i * NT;
      const int h = ln & 1, rs = ln >> 1;
      const int row = rs % ROWS, seg = rs / ROWS;
      const int grow = row0 + row, gcol = col0 + seg * 64 + h * 32;
      if (ln < LINES && grow < row_end && gcol < col_end) l2_prefetch_line(G + (int64_t)grow * ld + gcol);
    }
  }

  template <bool WITH_LO>
  __device__ static __forceinline__ void store(uint32_t s_hi, uint32_t s_lo, int tid, const float4 (&v)[CH][2]) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int ch = tid + i * NT;
      const int c = ch & 7, rs = ch >> 3;
      const int row = rs % ROWS, seg = rs / ROWS;
      const uint32_t off = (uint32_t)seg * (ROWS * 128) + (uint32_t)(row >> 3) * 1024 + (uint32_t)(row & 7) * 128 +
                           (uint32_t)((c ^ (row & 7)) << 4);
      const float x[8] = {v[i][0].x, v[i][0].y, v[i][0].z, v[i][0].w, v[i][1].x, v[i][1].y, v[i][1].z, v[i][1].w};
      // hi = bf16(x) by the packed convert (F2FP.BF16.F32.PACK_AB, two elements per instruction); the fp32 value of
      // hi needed for lo = bf16(x - hi) is the same 16 bits shifted back up, so no single-element F2F is issued
      uint32_t hp[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) hp[j] = pack_bf16(x[2 * j], x[2 * j + 1]);
      sts128(s_hi + off, make_uint4(hp[0], hp[1], hp[2], hp[3]));
      if (WITH_LO) {
        uint32_t lp[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float h0 = __uint_as_float(hp[j] << 16), h1 = __uint_as_float(hp[j] & 0xffff0000u);
          lp[j] = pack_bf16(x[2 * j] - h0, x[2 * j + 1] - h1);
        }
        uint4 lo = make_uint4(lp[0], lp[1], lp[2], lp[3]);
        sts128(s_lo + off, lo);
      }
    }
  }
};

template <int BN, int BK, bool X3, int STAGES>
struct SmemPlan {
  static constexpr int PLANES = X3 ? 2 : 1;
  static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;         // one bf16 plane
  static constexpr int STAGE_BYTES = (A_BYTES + B_BYTES) * PLANES;
  static constexpr int EPI_BOX_BYTES = 32 * 32 * 4;                           // one TMA store box (1024-byte aligned)
  static constexpr int EPI_WARP_BYTES = EPI_BOX_BYTES;                        // one box per warp (the pair of a quarter alternates)
  static constexpr int EPI_BYTES = EPI_WARPS * EPI_WARP_BYTES;
  static constexpr int BIAS_BYTES = BN * 4;                                    // this tile's bias slice, shared by the epilogue warps
  static constexpr size_t TOTAL = (size_t)STAGES * STAGE_BYTES + EPI_BYTES + BIAS_BYTES + 1024;
  static_assert(STAGE_BYTES % 1024 == 0, "stages must keep the 1024-byte swizzle alignment");
};

// Block = 20 warps: 8 producers (2 warpgroups), MMA, TMA, 2 idle, 8 epilogue (2 warpgroups) -- every warpgroup is
// complete because setmaxnreg is warpgroup-aligned.  The kernel starts at 96 registers/thread; producers raise their
// budget to PRODUCER_REGS (their operand tiles live in registers, DEPTH k-blocks in flight), everyone else drops to
// OTHER_REGS.  Eight epilogue warps because the epilogue, not the MMA, paced the wide-output shapes: with four warps one
// 128 x 256 tile took ~19 k cycles to drain (TMEM load -> bias / ReLU / bitmask -> staging -> TMA store, a serial chain
// per 32-column chunk) against 6.1 k cycles of MMA issue at K = 256 (profiles/r02_gemm_epilogue_bisect.txt).
constexpr int BLOCK_THREADS = 640;
constexpr int EPI_WARP0 = 12;                             // first epilogue warp (warps 12..19)
constexpr int PRODUCER_REGS = 144, OTHER_REGS = 64;      // 256*144 + 384*64 = 61440 = 640*96: the CTA pool is what the launch allocated
__device__ __forceinline__ void regs_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(PRODUCER_REGS)); }
__device__ __forceinline__ void regs_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(OTHER_REGS)); }

// EPI: epilogue specialisation (the epilogue of a lane quarter issues from one SM sub-partition: run-time flag tests per
// 16-column slice are not free).  0 generic (every flag tested at run time); 1 linear: alpha = 1, optional bias, optional
// output row mask, store or reduce-add; 2 bias + ReLU + sign bitmask out (+ optional dropout); 3 ReLU gate from the saved
// bitmask, alpha, no bias.
template <int BN, int BK, bool A_MN, bool B_MN, bool X3, bool B_TMA, int PW, int STAGES, int EPI>
__global__ void __launch_bounds__(BLOCK_THREADS, 1)
gemm_tc_kernel(const Args p, const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
               const __grid_constant__ CUtensorMap tm_c) {
  poet_pdl_launch_dependents();                  // dependents may be scheduled; they wait for our completion themselves
  using SP = SmemPlan<BN, BK, X3, STAGES>;
  constexpr int NT = PW * 32;                                        // producer threads
  constexpr int PLANES = SP::PLANES, A_BYTES = SP::A_BYTES, B_BYTES = SP::B_BYTES, STAGE_BYTES = SP::STAGE_BYTES;
  // both operands fp32 and MN-major (weight gradient): B is register-prefetched like A
  constexpr bool B_PREFETCH = !B_TMA && A_MN && B_MN;
  static_assert(BK == 64 || (A_MN && B_MN), "K-major operands need BK = 64 (one 128-byte swizzle row)");
  using TA = Tile<A_MN ? BK : BM, A_MN ? BM / 64 : 1, NT>;
  using TB = Tile<B_MN ? BK : BN, B_MN ? BN / 64 : 1, NT>;
  constexpr int BCH = B_PREFETCH ? TB::CH : 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bars[2 * STAGES + 4];
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[STAGES]);
  const uint32_t tfull0 = smem_u32(&bars[2 * STAGES]), tempty0 = smem_u32(&bars[2 * STAGES + 2]);
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, PW + (B_TMA ? 1 : 0)); mbar_init(empty0 + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == PW) tmem_alloc(smem_u32(&s_tmem), 2 * BN);             // two accumulator buffers
  if (warp == PW + 1 && lane == 0) {
    if (B_TMA) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_hi) : "memory");
      if (X3) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_lo) : "memory");
    }
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_c) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  poet_pdl_wait();                                                // prologue done; from here on we touch global memory
  if (warp < PW) regs_inc(); else regs_dec();                     // warpgroup-uniform

  if (warp < PW) {
    // ===================== producers =====================
    int it = 0;                                                    // k-blocks published so far (ring position)
    const bool do_colsum = B_PREFETCH && p.a_colsum != nullptr;
    float cs[B_PREFETCH ? TA::CH : 1][8];
#pragma unroll
    for (int i = 0; i < (B_PREFETCH ? TA::CH : 1); ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) cs[i][j] = 0.f;
    auto publish = [&](const Work& wk, int kb, const float4 (&va)[TA::CH][2], const float4 (&vbp)[BCH][2]) {
      const int s = it % STAGES;
      const uint32_t a_hi = smem_u32(smem) + s * STAGE_BYTES;
      const uint32_t b_hi = a_hi + A_BYTES * PLANES;
      if constexpr (B_TMA) {
        if (it >= STAGES) mbar_wait(empty0 + 8 * s, ((it / STAGES) - 1) & 1);
        if (!(POET_DBG(p, 2))) TA::template store<X3>(a_hi, a_hi + A_BYTES, tid, va);
      } else if constexpr (B_PREFETCH) {
        if (do_colsum && wk.n0 == 0) {                          // bias gradient: column sums of the dY tile, for free
#pragma unroll
          for (int i = 0; i < TA::CH; ++i) {
            cs[i][0] += va[i][0].x; cs[i][1] += va[i][0].y; cs[i][2] += va[i][0].z; cs[i][3] += va[i][0].w;
            cs[i][4] += va[i][1].x; cs[i][5] += va[i][1].y; cs[i][6] += va[i][1].z; cs[i][7] += va[i][1].w;
          }
        }
        if (it >= STAGES) mbar_wait(empty0 + 8 * s, ((it / STAGES) - 1) & 1);
        TA::template store<X3>(a_hi, a_hi + A_BYTES, tid, va);
        TB::template store<X3>(b_hi, b_hi + B_BYTES, tid, vbp);
        if (do_colsum && wk.n0 == 0 && kb == wk.nkb - 1) {      // item finished: fold the 4 row groups of the warp, then atomics
#pragma unroll
          for (int i = 0; i < TA::CH; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float v = cs[i][j];
              v += __shfl_xor_sync(0xffffffffu, v, 8);
              v += __shfl_xor_sync(0xffffffffu, v, 16);
              cs[i][j] = 0.f;
              // chunk i of this thread: columns (seg = i) * 64 + (tid & 7) * 8 + j of the tile (Tile<> chunk map, ROWS = BK = 32)
              const int gcol = wk.m0 + i * 64 + (tid & 7) * 8 + j;
              if (lane < 8 && gcol < p.M) atomicAdd(p.a_colsum + gcol, v);
            }
          }
        }
      } else {                                                  // fp32 activation B: loads issued before the stage wait
        float4 vb[TB::CH][2];
        const int k0 = (wk.kb0 + kb) * BK;
        if (B_MN) TB::load(p.B, p.ldb, k0, p.K, wk.n0, p.N, tid, vb);
        else      TB::load(p.B, p.ldb, wk.n0, p.N, k0, p.K, tid, vb);
        if (it >= STAGES) mbar_wait(empty0 + 8 * s, ((it / STAGES) - 1) & 1);
        TA::template store<X3>(a_hi, a_hi + A_BYTES, tid, va);
        TB::template store<X3>(b_hi, b_hi + B_BYTES, tid, vb);
      }
      fence_proxy_async();                                      // generic-proxy smem writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(full0 + 8 * s);                // one arrival per producer warp (256 arrivals on one
      ++it;                                                     // mbarrier cost ~1000 cycles per k-block)
    };
    auto loadAB = [&](const Work& wk, int kb, float4 (&v)[TA::CH][2], float4 (&vbp)[BCH][2]) {
      if (POET_DBG(p, 1)) {
#pragma unroll
        for (int i = 0; i < TA::CH; ++i) { v[i][0] = make_float4(1.f, 1.f, 1.f, 1.f); v[i][1] = v[i][0]; }
      } else {
        const int k0 = (wk.kb0 + kb) * BK;
        if (p.fast_k) {                                          // CTA-uniform
          if (A_MN) TA::template load<true>(p.A, p.lda, k0, p.K, wk.m0, p.M, tid, v, p.a_row_mask);
          else      TA::template load<true>(p.A, p.lda, wk.m0, p.M, k0, p.K, tid, v, p.a_row_mask);
        } else {
          if (A_MN) TA::load(p.A, p.lda, k0, p.K, wk.m0, p.M, tid, v, p.a_row_mask);
          else      TA::load(p.A, p.lda, wk.m0, p.M, k0, p.K, tid, v, p.a_row_mask);
        }
      }
      if constexpr (B_PREFETCH) {
        const int k0 = (wk.kb0 + kb) * BK;
        if (p.fast_k) TB::template load<true>(p.B, p.ldb, k0, p.K, wk.n0, p.N, tid, vbp);
        else          TB::load(p.B, p.ldb, k0, p.K, wk.n0, p.N, tid, vbp);
      }
    };
    // L2 prefetch hints.  K-major A: when a work item starts, the next item's rows (this CTA's next tile) are
    // requested from HBM, a whole tile period before the register prefetch asks for them.
    // MN-major operands (weight gradient): lines PF_DIST k-blocks ahead inside the same item.
    constexpr int PF_DIST = 4;
    auto hints = [&](const Work& wk, int kb, int w) {
      if (!p.l2_prefetch) return;
      if constexpr (!A_MN) {
        if (kb == 0 && tid < BM) {
          const int nw = w + gridDim.x;
          if (nw < p.total_work) {
            const Work nx = decode<BK>(p, nw, BN);
            const int row = nx.m0 + tid, k0 = nx.kb0 * BK;
            if (row < p.M) l2_prefetch_bulk(p.A + (int64_t)row * p.lda + k0, (uint32_t)(min(nx.nkb * BK, p.K - k0) * 4));
          }
        }
      } else {
        const int kpf = kb + PF_DIST;
        if (kpf < wk.nkb) {
          const int k0 = (wk.kb0 + kpf) * BK;
          TA::prefetch(p.A, p.lda, k0, p.K, wk.m0, p.M, tid);
          if constexpr (B_PREFETCH) TB::prefetch(p.B, p.ldb, k0, p.K, wk.n0, p.N, tid);
        }
      }
    };
    // flattened (work item, k-block) sequence; DEPTH k-blocks of operand tiles are in flight in registers
    // (ring of DEPTH+1 buffers, rotation resolved at compile time by unrolling the ring)
    constexpr int DEPTH = B_PREFETCH ? (BN == 256 ? 1 : 2) : 2;
    constexpr int NBUF = DEPTH + 1;
    struct Pos { Work wk; int kb, w; bool valid; };
    auto next_pos = [&](const Pos& c) {
      Pos n = c;
      if (!c.valid) return n;
      n.kb = c.kb + 1;
      if (n.kb == c.wk.nkb) {
        n.kb = 0;
        n.w = c.w + gridDim.x;
        if (n.w < p.total_work) n.wk = decode<BK>(p, n.w, BN); else n.valid = false;
      }
      return n;
    };
    float4 ta[NBUF][TA::CH][2], tb[NBUF][BCH][2];
    Pos pos[NBUF];
    pos[0].w = blockIdx.x; pos[0].kb = 0; pos[0].valid = pos[0].w < p.total_work;
    if (pos[0].valid) pos[0].wk = decode<BK>(p, pos[0].w, BN); else pos[0].wk = Work{0, 0, 0, 0, 1};
#pragma unroll
    for (int d = 1; d < NBUF; ++d) pos[d] = next_pos(pos[d - 1]);
#pragma unroll
    for (int d = 0; d < DEPTH; ++d)
      if (pos[d].valid) loadAB(pos[d].wk, pos[d].kb, ta[d], tb[d]);
    while (pos[0].valid) {
#pragma unroll
      for (int u = 0; u < NBUF; ++u) {
        // slot u holds the tile to publish now; slot (u + DEPTH) % NBUF is free: fetch the tile DEPTH ahead into it
        const int f = (u + DEPTH) % NBUF;
        if (pos[u].valid) {
          if (pos[f].valid) loadAB(pos[f].wk, pos[f].kb, ta[f], tb[f]);
          hints(pos[u].wk, pos[u].kb, pos[u].w);
          publish(pos[u].wk, pos[u].kb, ta[u], tb[u]);
          // slot u becomes the position NBUF ahead
          Pos nx = pos[(u + NBUF - 1) % NBUF];
          pos[u] = next_pos(nx);
        }
      }
    }
  } else if (warp == PW) {
    if (lane == 0) {
      // ===================== MMA issuer (one thread) =====================

      // K-major: 8-row groups 1024 B apart, a 16-wide k-step is 32 B inside the 128 B swizzle row.
      // MN-major: 64-element m/n groups BK*128 B apart (LBO), 8-row k groups 1024 B apart (SBO),
      //           a 16-wide k-step is two k groups = 2048 B.
      constexpr uint32_t A_LBO = A_MN ? BK * 128 : 16, A_STEP = A_MN ? 2048 : 32;
      constexpr uint32_t B_LBO = B_MN ? BK * 128 : 16, B_STEP = B_MN ? 2048 : 32;
      int it = 0, tcnt = 0;
      for (int w = blockIdx.x; w < p.total_work; w += gridDim.x, ++tcnt) {
        const Work wk = decode<BK>(p, w, BN);
        const int ab = tcnt & 1;
        if (tcnt >= 2) mbar_wait(tempty0 + 8 * ab, ((tcnt >> 1) - 1) & 1);    // epilogue drained this buffer
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(ab * BN);
        const uint32_t idesc = make_idesc(wk.bn, A_MN, B_MN);      // MMA N = this item's width (BN, or a tail slice)
        for (int i = 0; i < wk.nkb; ++i, ++it) {
          const int s = it % STAGES;
          mbar_wait(full0 + 8 * s, (it / STAGES) & 1);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t a_hi = st, a_lo = st + A_BYTES, b_hi = st + A_BYTES * PLANES, b_lo = b_hi + B_BYTES;
#pragma unroll
          for (int j = 0; j < BK / 16; ++j) {
            if (POET_DBG(p, 32)) break;
            const uint64_t dah = make_desc(a_hi + j * A_STEP, A_LBO, 1024);
            const uint64_t dbh = make_desc(b_hi + j * B_STEP, B_LBO, 1024);
            const uint32_t first = (i | j) ? 1u : 0u;
            if (X3) {
              const uint64_t dal = make_desc(a_lo + j * A_STEP, A_LBO, 1024);
              const uint64_t dbl = make_desc(b_lo + j * B_STEP, B_LBO, 1024);
              umma_bf16(tacc, dah, dbl, idesc, first);            // small cross terms first
              umma_bf16(tacc, dal, dbh, idesc, 1u);
              umma_bf16(tacc, dah, dbh, idesc, 1u);
            } else {
              umma_bf16(tacc, dah, dbh, idesc, first);
            }
          }
          umma_commit(empty0 + 8 * s);                             // stage reusable once these MMAs retire
        }
        umma_commit(tfull0 + 8 * ab);                              // accumulator complete
      }
    }
  } else if (warp == PW + 1) {
    if (B_TMA && lane == 0) {
      // ===================== TMA: pre-split bf16 weight planes =====================
      int it = 0;
      for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
        const Work wk = decode<BK>(p, w, BN);
        for (int i = 0; i < wk.nkb; ++i, ++it) {
          const int s = it % STAGES;
          if (it >= STAGES) mbar_wait(empty0 + 8 * s, ((it / STAGES) - 1) & 1);
          const uint32_t bar = full0 + 8 * s;
          if (POET_DBG(p, 4)) { mbar_arrive(bar); continue; }
          mbar_arrive_expect_tx(bar, B_BYTES * PLANES);
          const int k0 = (wk.kb0 + i) * BK;
          const uint32_t b_hi = smem_u32(smem + s * STAGE_BYTES + A_BYTES * PLANES), b_lo = b_hi + B_BYTES;
          if (B_MN) {
#pragma unroll
            for (int g = 0; g < BN / 64; ++g) {                      // box = 64 (n, contiguous) x BK (k rows)
              tma_load_2d(b_hi + g * (BK * 128), &tm_hi, wk.n0 + g * 64, k0, bar);
              if (X3) tma_load_2d(b_lo + g * (BK * 128), &tm_lo, wk.n0 + g * 64, k0, bar);
            }
          } else {                                                   // box = 64 (k, contiguous) x BN (n rows)
            tma_load_2d(b_hi, &tm_hi, k0, wk.n0, bar);
            if (X3) tma_load_2d(b_lo, &tm_lo, k0, wk.n0, bar);
          }
        }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue warps =====================
    // TMEM hands each thread one accumulator ROW.  Warp e handles lane quarter (warp & 3) and the 32-column chunks
    // e>>2, e>>2 + 2, ... of the tile, 16 columns at a time (the budget is 64 registers): all epilogue math happens on
    // the row in registers; the row is written into a SWIZZLE_128B staging box (16-byte chunk c of row r at chunk
    // c ^ (r & 7): conflict-free for the quarter-warps) and one lane issues the TMA store of the 32 x 32 box.  Rows
    // beyond M are clipped by the tensor map.  One box per warp: while its store drains, the quarter's other warp works.
    const int ew = warp - EPI_WARP0;                                 // 0..7
    const int quarter = warp & 3, half = ew >> 2;                    // TMEM lane quarter this warp may access; chunk parity
    const uint32_t box = smem_u32(smem) + STAGES * STAGE_BYTES + ew * SP::EPI_WARP_BYTES;
    constexpr bool G = EPI == 0;
    constexpr bool MAY_ALPHA = G || EPI == 3, MAY_BIAS = EPI != 3, MAY_RELU = G || EPI == 2, MAY_DROP = G || EPI == 2;
    constexpr bool MAY_GATE_BITS = G || EPI == 3, MAY_GATE_F32 = G, MAY_DEAD = G || EPI == 1;
    const bool relu = G ? (p.flags & POET_GEMM_RELU) != 0 : EPI == 2;
    const bool accum = p.flags & POET_GEMM_ACCUMULATE;
    const float alpha = p.alpha;
    const bool reduce = accum || p.splits > 1;
    const int words = p.N >> 5;                                      // bitmask words per row
    bool pending = false;                                            // a store of this warp's box may still be reading it
    // The tile's bias slice is staged in shared memory once per tile (a dependent global load per 32-column
    // chunk costs ~500 cycles of epilogue latency).
    const uint32_t bias_s = smem_u32(smem) + STAGES * STAGE_BYTES + SP::EPI_BYTES;
    const int et = tid - EPI_WARP0 * 32;                             // 0..255 over the epilogue warps
    int tcnt = 0;
    for (int w = blockIdx.x; w < p.total_work; w += gridDim.x, ++tcnt) {
      const Work wk = decode<BK>(p, w, BN);
      const int ab = tcnt & 1;
      const int mrow0 = wk.m0 + quarter * 32;
      const int row = mrow0 + lane;
      const bool row_ok = row < p.M;
      const bool dead = MAY_DEAD && p.row_mask != nullptr && row_ok && p.row_mask[row] != 0;
      const bool has_bias = MAY_BIAS && p.bias != nullptr && wk.split == 0;
      if (has_bias) {
        asm volatile("bar.sync 1, 256;" ::: "memory");               // every epilogue warp is done with the previous slice
        if (et < wk.bn) {
          const float b1 = __ldg(p.bias + wk.n0 + et);
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_s + et * 4), "f"(b1) : "memory");
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      mbar_wait(tfull0 + 8 * ab, (tcnt >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int col = half * 32; col < wk.bn; col += 64) {            // wk.bn >= 64: every warp owns at least one chunk
        const int n0c = wk.n0 + col;
        const bool last = col + 64 >= wk.bn;
        const bool live = mrow0 < p.M && !(POET_DBG(p, 8));             // warp-uniform
        uint32_t bits = 0, gbits = 0xffffffffu, keep = 0xffffffffu;
        if (live && !(POET_DBG(p, 16))) {
          if (MAY_GATE_BITS && (EPI == 3 || p.gate_bits != nullptr))
            gbits = row_ok ? __ldg(p.gate_bits + (int64_t)row * words + (n0c >> 5)) : 0u;
          // nn.Dropout on the epilogue's output (reference: dropout2 / dropout3 on relu(linear1(x)),
          // deformable_transformer.py:194,268).  The keep mask is folded into the ReLU sign bitmask, so the backward
          // needs no second mask: the dgrad gates on (pre-activation > 0 AND kept) and scales by 1/(1-p) through alpha.
          if (MAY_DROP && p.drop.seed != nullptr) {
            const PoetDropKey key = poet_drop_key(p.drop);
            const uint64_t pair0 = ((uint64_t)row * (uint64_t)p.N + (uint64_t)n0c) >> 1;      // N and n0c are even
            keep = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) keep |= poet_drop_keep2(key, pair0 + j, p.drop.threshold16) << (2 * j);
          }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float v[16];
          tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ab * BN + col + 16 * h), v);
          if (last && h == 1) {                                      // last read of this accumulator by this warp: release it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * ab);
          }
          if (!live) continue;
          if (!(POET_DBG(p, 16))) {
            if (MAY_ALPHA && alpha != 1.f) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] *= alpha;
            }
            if (has_bias) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 b4 = lds128(bias_s + (col + 16 * h + 4 * j) * 4);  // warp-uniform address: one broadcast wavefront
                v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
              }
            }
            if (MAY_RELU && relu) {
              if (EPI == 2 || p.relu_bits != nullptr) {
#pragma unroll
                for (int i = 0; i < 16; ++i) bits |= (v[i] > 0.f ? 1u : 0u) << (16 * h + i);
              }
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
            }
            if (MAY_DROP && p.drop.seed != nullptr) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = ((keep >> (16 * h + i)) & 1u) ? v[i] * p.drop.scale16 : 0.f;
            }
            if (MAY_GATE_BITS && (EPI == 3 || p.gate_bits != nullptr)) {   // ReLU backward from the saved sign bitmask
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = ((gbits >> (16 * h + i)) & 1u) ? v[i] : 0.f;
            } else if (MAY_GATE_F32 && p.gate != nullptr) {          // ReLU backward from the fp32 activation
              if (row_ok) {
                const float* gp = p.gate + (int64_t)row * p.ldc + n0c + 16 * h;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float4 g4 = ldg4(gp + 4 * j);
                  v[4 * j] = g4.x > 0.f ? v[4 * j] : 0.f; v[4 * j + 1] = g4.y > 0.f ? v[4 * j + 1] : 0.f;
                  v[4 * j + 2] = g4.z > 0.f ? v[4 * j + 2] : 0.f; v[4 * j + 3] = g4.w > 0.f ? v[4 * j + 3] : 0.f;
                }
              }
            }
            if (MAY_DEAD && dead) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = 0.f;
            }
          }
          if (h == 0 && pending && !(POET_DBG(p, 64))) {                // the previous store of this warp has read the box
            if (lane == 0) bulk_wait_read_0();
            __syncwarp();
          }
          if (!(POET_DBG(p, 512))) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              sts128(box + lane * 128 + (((4 * h + j) ^ (lane & 7)) << 4),
                     make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
                                __float_as_uint(v[4 * j + 3])));
          }
        }
        if (!live) continue;
        if (MAY_RELU && relu && (EPI == 2 || p.relu_bits != nullptr) && row_ok && !(POET_DBG(p, 16)))
          p.relu_bits[(int64_t)row * words + (n0c >> 5)] = dead ? 0u : (bits & keep);
        if (!(POET_DBG(p, 128))) fence_proxy_async();                   // generic-proxy writes -> visible to the TMA engine
        __syncwarp();
        if (lane == 0 && !(POET_DBG(p, 256))) {
          if (reduce) tma_reduce_add_2d(&tm_c, box, n0c, mrow0);
          else        tma_store_2d(&tm_c, box, n0c, mrow0);
          bulk_commit();
        }
        pending = true;
      }
    }
    if (lane == 0) bulk_wait_all();                                  // staging smem must outlive the last store
  }
  tc_fence_before();
  __syncthreads();
  if (warp == PW) tmem_dealloc(tmem_base, 2 * BN);
}

// ---- fp32 -> bf16 hi/lo planes (weights, once per step) --------------------------------------
__global__ void __launch_bounds__(256) split_bf16_kernel(const float4* __restrict__ src, uint2* __restrict__ hi,
                                                         uint2* __restrict__ lo, int64_t n4) {
  poet_pdl_wait();          // no early launch of dependents: consumers may prefetch the planes before their own wait (POET_GEMM_B_STABLE)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 x = src[i];
    const float h0 = __bfloat162float(__float2bfloat16_rn(x.x)), h1 = __bfloat162float(__float2bfloat16_rn(x.y));
    const float h2 = __bfloat162float(__float2bfloat16_rn(x.z)), h3 = __bfloat162float(__float2bfloat16_rn(x.w));
    hi[i] = make_uint2(pack_bf16(h0, h1), pack_bf16(h2, h3));
    if (lo) lo[i] = make_uint2(pack_bf16(x.x - h0, x.y - h1), pack_bf16(x.z - h2, x.w - h3));
  }
}

// All weight matrices of a model in ONE launch: table[t] = {src, hi, lo, first_chunk}, chunk = 1024 float4.
struct SplitEntry { const float4* src; uint2* hi; uint2* lo; int64_t n4; int64_t first_chunk; };
__global__ void __launch_bounds__(256) split_bf16_multi_kernel(const SplitEntry* __restrict__ table, int n_tensors) {
  poet_pdl_wait();          // no early launch of dependents: consumers may prefetch the planes before their own wait (POET_GEMM_B_STABLE)
  // binary search of the chunk's tensor (n_tensors is ~100: 7 steps)
  const int64_t chunk = blockIdx.x;
  int lo_i = 0, hi_i = n_tensors - 1;
  while (lo_i < hi_i) {
    const int mid = (lo_i + hi_i + 1) >> 1;
    if (table[mid].first_chunk <= chunk) lo_i = mid; else hi_i = mid - 1;
  }
  const SplitEntry e = table[lo_i];
  const int64_t base = (chunk - e.first_chunk) * 1024;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int64_t i = base + threadIdx.x + j * 256;
    if (i < e.n4) {
      const float4 x = e.src[i];
      const float h0 = __bfloat162float(__float2bfloat16_rn(x.x)), h1 = __bfloat162float(__float2bfloat16_rn(x.y));
      const float h2 = __bfloat162float(__float2bfloat16_rn(x.z)), h3 = __bfloat162float(__float2bfloat16_rn(x.w));
      e.hi[i] = make_uint2(pack_bf16(h0, h1), pack_bf16(h2, h3));
      if (e.lo) e.lo[i] = make_uint2(pack_bf16(x.x - h0, x.y - h1), pack_bf16(x.z - h2, x.w - h3));
    }
  }
}

// ---- host side ------------------------------------------------------------------------------
// 2-D bf16 tensor map: `inner` contiguous elements per row, `rows` rows of stride ld elements,
// box = 64 x box_rows, SWIZZLE_128B (the box row is exactly one 128-byte swizzle span).
static int make_map(CUtensorMap* map, const void* base, int64_t inner, int64_t rows, int64_t ld, int box_rows) {
  return poet_tma::encode_2d(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, (uint64_t)inner, (uint64_t)rows, (uint64_t)ld * 2,
                             64u, (uint32_t)box_rows, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B)
             ? POET_OK : POET_ERR_UNSUPPORTED;
}

// 2-D fp32 output map: N contiguous columns, M rows of stride ldc floats, box = 32 x 32 (128-byte rows), SWIZZLE_128B.
static int make_map_c(CUtensorMap* map, float* base, int64_t N, int64_t M, int64_t ldc) {
  return poet_tma::encode_2d(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, (uint64_t)N, (uint64_t)M, (uint64_t)ldc * 4, 32u,
                             32u, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)
             ? POET_OK : POET_ERR_UNSUPPORTED;
}

struct Maps {
  CUtensorMap hi, lo, c;
};

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

template <int BN, int BK, bool A_MN, bool B_MN, bool X3, bool B_TMA, int STAGES, int EPI = 0>
int launch(Args a, const Maps& m, cudaStream_t s) {
  constexpr int PW = 8;
  constexpr size_t smem = SmemPlan<BN, BK, X3, STAGES>::TOTAL;
  static_assert(smem <= 227 * 1024 - 256, "shared memory plan exceeds the 227 KB per-CTA limit");
  auto kern = gemm_tc_kernel<BN, BK, A_MN, B_MN, X3, B_TMA, PW, STAGES, EPI>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int grid = a.total_work < a.grid_limit ? a.total_work : a.grid_limit;       // persistent: one CTA per SM
  poet_launch(kern, dim3(grid), dim3(BLOCK_THREADS), smem, s, a, m.hi, m.lo, m.c);
  return poet_launch_status();
}

// stage counts: 128 x 256 tiles, BK 64: 2 stages of 96 KB (x3); 128 x 128, BK 64: 3 x 64 KB;
// weight gradient (BK 32): 128 x 128: 6 x 32 KB, 128 x 256: 4 x 48 KB.  (+ 32 KB epilogue staging each)
template <int BN, bool X3>
int dispatch(const Args& a, bool a_mn, bool b_mn, bool b_tma, const Maps& m, cudaStream_t s, int epi) {
  // single-pass bf16 stages are half as large (one plane per operand): twice the pipeline depth in the same shared memory
  constexpr int ST64 = ((BN == 256) ? 2 : 3) * (X3 ? 1 : 2);
  constexpr int ST32 = ((BN == 256) ? 4 : 6);
  if (b_tma) {                                      // B is a pre-split weight: A is k-contiguous (forward / dgrad)
    if (a_mn) return POET_ERR_UNSUPPORTED;
    if constexpr (BN == 256) {                      // the token-row GEMMs of the step: specialised epilogues
      if (epi == 1) return b_mn ? launch<256, 64, false, true, X3, true, ST64, 1>(a, m, s) : launch<256, 64, false, false, X3, true, ST64, 1>(a, m, s);
      if (epi == 2) return b_mn ? launch<256, 64, false, true, X3, true, ST64, 2>(a, m, s) : launch<256, 64, false, false, X3, true, ST64, 2>(a, m, s);
      if (epi == 3) return b_mn ? launch<256, 64, false, true, X3, true, ST64, 3>(a, m, s) : launch<256, 64, false, false, X3, true, ST64, 3>(a, m, s);
    }
    if constexpr (BN == 128) {
      static const int st2 = env_int("POET_GEMM_STAGES128", 3) == 2;      // pipeline-depth experiment
      if (st2) return b_mn ? launch<128, 64, false, true, X3, true, 2>(a, m, s) : launch<128, 64, false, false, X3, true, 2>(a, m, s);
    }
    return b_mn ? launch<BN, 64, false, true, X3, true, ST64>(a, m, s) : launch<BN, 64, false, false, X3, true, ST64>(a, m, s);
  }
  if (a_mn && b_mn) {                               // weight gradient
    if constexpr (BN == 256) { if (epi == 1) return launch<256, 32, true, true, X3, false, ST32, 1>(a, m, s); }
    return launch<BN, 32, true, true, X3, false, ST32>(a, m, s);
  }
  if constexpr (BN == 128) {
    if (!a_mn && !b_mn) return launch<128, 64, false, false, X3, false, 3>(a, m, s);
    if (!a_mn && b_mn) return launch<128, 64, false, true, X3, false, 3>(a, m, s);
    return launch<128, 64, true, false, X3, false, 3>(a, m, s);
  }
  return POET_ERR_UNSUPPORTED;
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

// 1 when relu_bits / gate_bits are honoured (TMA epilogue enabled); the Python side asks before using them.
int poet_gemm_tc_bits_supported() { return 1; }

// b_hi / b_lo != nullptr: B was pre-split into bf16 planes (same logical layout / ldb as the fp32 B).
int poet_gemm_tc(const float* A, int64_t lda, int a_kcontig, const float* Bm, const void* b_hi, const void* b_lo,
                 int64_t ldb, int b_kcontig, float* C, int64_t ldc, int M, int N, int K, float alpha, const float* bias,
                 const float* gate, const uint8_t* row_mask, uint32_t* relu_bits, const uint32_t* gate_bits, float* a_colsum,
                 const uint8_t* a_row_mask, int flags, int precision, cudaStream_t s, const PoetDropout* drop) {
  POET_REQUIRE(poet_aligned16(A) && poet_aligned16(C), POET_ERR_BAD_ALIGNMENT);
  POET_REQUIRE(!bias || poet_aligned16(bias), POET_ERR_BAD_ALIGNMENT);
  POET_REQUIRE(!gate || poet_aligned16(gate), POET_ERR_BAD_ALIGNMENT);
  static const int dbg = tc::env_int("POET_GEMM_DEBUG", 0);
  static const int l2pf = tc::env_int("POET_GEMM_L2_PREFETCH", 1);
  static const int wgrad_bn = tc::env_int("POET_GEMM_WGRAD_BN", 256);
  const bool dropping = drop != nullptr && drop->seed != nullptr;
  POET_REQUIRE(!dropping || N % 32 == 0, POET_ERR_UNSUPPORTED);
  const bool x3 = precision == POET_GEMM_BF16X3;
  const bool b_tma = b_hi != nullptr && (!x3 || b_lo != nullptr) && a_kcontig && (ldb % 8 == 0) &&
                     poet_aligned16(b_hi) && (!x3 || poet_aligned16(b_lo));
  POET_REQUIRE(b_tma || (Bm != nullptr && poet_aligned16(Bm)), POET_ERR_NULL_POINTER);
  const bool a_mn = !a_kcontig, b_mn = !b_kcontig;
  const bool wgrad = !b_tma && a_mn && b_mn;
  POET_REQUIRE(a_colsum == nullptr || wgrad, POET_ERR_UNSUPPORTED);
  const int bk = wgrad ? 32 : 64;
  tc::Args a;
  a.A = A; a.lda = lda; a.B = b_tma ? nullptr : Bm; a.ldb = ldb; a.C = C; a.ldc = ldc; a.M = M; a.N = N; a.K = K;
  a.alpha = alpha; a.bias = bias; a.gate = gate; a.row_mask = row_mask; a.relu_bits = relu_bits; a.gate_bits = gate_bits; a.a_colsum = a_colsum; a.a_row_mask = a_row_mask;
  a.flags = flags;
  memset(&a.drop, 0, sizeof(a.drop));
  if (dropping) a.drop = *drop;
  a.l2_prefetch = l2pf; a.debug = dbg;
  // POET_GEMM_BACKGROUND: this GEMM runs on a side stream next to a chain of latency-bound kernels (the decoder's value
  // projections of `memory` and their backward).  A persistent CTA per SM with ~200 KB of shared memory would keep every
  // small kernel of that chain off the machine for the whole GEMM, so a background launch leaves some SMs free.
  static const int bg_sms = tc::env_int("POET_GEMM_BG_SMS", 116);
  const int sms = (flags & POET_GEMM_BACKGROUND) ? (bg_sms < 8 ? 8 : (bg_sms > POET_NUM_SMS ? POET_NUM_SMS : bg_sms)) : POET_NUM_SMS;
  a.grid_limit = sms;
  const int m_tiles = poet_ceil_div(M, tc::BM);
  const int total_kb = poet_ceil_div(K, bk);
  // Tile width: rounds of the persistent grid x work per round.  128-wide tiles quantise better over 148 SMs
  // but convert every A tile twice as often, hence the 15 % handicap.
  int bn = 128;
  if (N % 256 == 0 && b_tma) {
    // 256-wide tiles convert each A tile once per 256 output columns; measured on the cfg2 shapes they are never
    // slower than 128-wide ones once there is at least one tile per SM (K=1024: 63 vs 77 us), so the 128-wide
    // tile is kept for small problems only (more CTAs in flight).
    if ((int64_t)(N / 256) * m_tiles >= sms) bn = 256;
    static const int force_bn = tc::env_int("POET_GEMM_FORCE_BN", 0);
    if (force_bn == 128 || force_bn == 256) bn = force_bn;
  }
  // weight gradient: the wide tile halves the L2->SM operand traffic per flop; split-K fills the machine
  if (wgrad && N % 256 == 0 && wgrad_bn == 256 && total_kb >= 16) bn = 256;
  const int64_t tiles = (int64_t)(N / bn) * m_tiles;
  int splits = 1;
  const bool linear_epi = !(flags & POET_GEMM_RELU) && gate == nullptr && gate_bits == nullptr && row_mask == nullptr;
  const int min_kb = 256 / bk;                                                     // at least 256 k per split
  if (linear_epi && !a_kcontig && tiles < sms && total_kb >= 2 * min_kb) {  // weight-gradient shape
    splits = (int)(sms / tiles);
    if (splits > total_kb / min_kb) splits = total_kb / min_kb;
    if (splits < 1) splits = 1;
  }
  a.kb_per_split = poet_ceil_div(total_kb, splits);
  a.splits = poet_ceil_div(total_kb, a.kb_per_split);
  a.n_tiles = N / bn;
  a.total_kb = total_kb;
  a.fast_k = (K % bk == 0 && M >= 8 && N >= 8 && K >= 8 && tc::env_int("POET_GEMM_FAST_LOADS", 1)) ? 1 : 0;
  a.total_work = a.n_tiles * m_tiles * a.splits;
  POET_REQUIRE(a.total_work < (1 << 20), POET_ERR_BAD_SHAPE);
  a.tail_start = a.total_work; a.tail_slices = 1;
  static const int tail_split = tc::env_int("POET_GEMM_TAIL_SPLIT", 1);
  if (tail_split && a.splits == 1 && !wgrad && a.total_work > sms) {
    const int rem = a.total_work % sms;
    int sl = 1;
    while (rem > 0 && sl * 2 <= bn / 64 && rem * sl * 2 <= sms + sms / 4) sl *= 2;
    if (sl > 1) {
      a.tail_start = a.total_work - rem;
      a.tail_slices = sl;
      a.total_work = a.tail_start + rem * sl;
    }
  }
  auto recip = [](int d) { return (((uint64_t)1 << 40) / (uint64_t)d) + 1; };
  a.r_splits = recip(a.splits); a.r_n_tiles = recip(a.n_tiles); a.r_tail_slices = recip(a.tail_slices);
  a.tail_shift = 0;
  while ((1 << a.tail_shift) < a.tail_slices) ++a.tail_shift;
  if (a.splits > 1 && !(flags & POET_GEMM_ACCUMULATE)) {
    cudaError_t e = cudaMemset2DAsync(C, ldc * sizeof(float), 0, (size_t)N * sizeof(float), M, s);
    if (e != cudaSuccess) return (int)e;
  }
  tc::Maps m;
  memset(&m, 0, sizeof(m));
  if (b_tma) {
    // K-major weight [N,K]: inner = K, rows = N, box rows = BN.  MN-major weight [K,N]: inner = N, rows = K, box rows = BK.
    const int64_t inner = b_mn ? N : K, rows = b_mn ? K : N;
    const int box_rows = b_mn ? bk : bn;
    int rc = tc::make_map(&m.hi, b_hi, inner, rows, ldb, box_rows);
    if (rc) return rc;
    if (x3) { rc = tc::make_map(&m.lo, b_lo, inner, rows, ldb, box_rows); if (rc) return rc; }
  }
  {
    int rc = tc::make_map_c(&m.c, C, N, M, ldc);
    if (rc) return rc;
  }
  // epilogue specialisation (see gemm_tc_kernel)
  const bool is_relu = flags & POET_GEMM_RELU;
  int epi = 0;
  if (gate_bits && !gate && !row_mask && !is_relu && !bias && !dropping && !relu_bits) epi = 3;
  else if (is_relu && relu_bits && !gate && !gate_bits && !row_mask && alpha == 1.f) epi = 2;
  else if (!is_relu && !gate && !gate_bits && !relu_bits && !dropping && alpha == 1.f) epi = 1;
  static const int epi_spec = tc::env_int("POET_GEMM_EPI_SPEC", 1);
  if (!epi_spec) epi = 0;
  if (bn == 256) return x3 ? tc::dispatch<256, true>(a, a_mn, b_mn, b_tma, m, s, epi) : tc::dispatch<256, false>(a, a_mn, b_mn, b_tma, m, s, epi);
  return x3 ? tc::dispatch<128, true>(a, a_mn, b_mn, b_tma, m, s, epi) : tc::dispatch<128, false>(a, a_mn, b_mn, b_tma, m, s, epi);
}

int poet_split_bf16_multi_impl(const void* table_dev, int n_tensors, int64_t total_chunks, cudaStream_t s) {
  POET_REQUIRE(table_dev != nullptr, POET_ERR_NULL_POINTER);
  POET_REQUIRE(n_tensors > 0 && total_chunks > 0 && total_chunks < ((int64_t)1 << 31), POET_ERR_BAD_SHAPE);
  static_assert(sizeof(tc::SplitEntry) == 40, "table layout is part of the ABI (5 x 8 bytes)");
  poet_launch(tc::split_bf16_multi_kernel, dim3((unsigned)total_chunks), dim3(256), 0, s, reinterpret_cast<const tc::SplitEntry*>(table_dev), n_tensors);
  return poet_launch_status();
}

int poet_split_bf16_impl(const float* src, void* hi, void* lo, int64_t n, cudaStream_t s) {
  POET_REQUIRE(src && hi, POET_ERR_NULL_POINTER);
  POET_REQUIRE(n > 0 && n % 4 == 0, POET_ERR_BAD_SHAPE);
  POET_REQUIRE(poet_aligned16(src) && (reinterpret_cast<uintptr_t>(hi) % 8 == 0) &&
               (!lo || reinterpret_cast<uintptr_t>(lo) % 8 == 0), POET_ERR_BAD_ALIGNMENT);
  int grid = poet_ceil_div(n / 4, 256);
  if (grid > POET_NUM_SMS * 8) grid = POET_NUM_SMS * 8;
  poet_launch(tc::split_bf16_kernel, dim3(grid), dim3(256), 0, s, reinterpret_cast<const float4*>(src), reinterpret_cast<uint2*>(hi),
                                             reinterpret_cast<uint2*>(lo), n / 4);
  return poet_launch_status();
}
