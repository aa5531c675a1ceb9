// Micro-benchmark: how fast can one B200 scatter-add the MSDA backward's grad_value traffic?
// 6.55 M sampling points (B=16, Lq=1600, M=16, L*P=16), each adding a 2x2 pixel footprint of 16 fp32 channels.
//   A  red.global.add.v4.f32, 4 lanes per corner, row-major [B,S,M,D] (the product's scheme: 64-byte segments 1 KB apart)
//   B  the same instructions on a head-major [B,M,S,D] buffer (x-adjacent corners contiguous: 128-byte segments)
//   C  cp.reduce.async.bulk (1-D, 64 bytes = one corner) from shared memory, row-major; one lane per corner
//   D  cp.reduce.async.bulk 128 bytes (a corner pair), head-major; one lane per footprint row
//   E  cp.reduce.async.bulk.tensor.5d, box (16 d, 1 m, 2 x, 2 y, 1 b) = the whole footprint in ONE op, row-major
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/red_micro tools/red_micro.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>

constexpr int B = 16, S = 1600, M = 16, D = 16, LQ = 1600, LP = 16;
constexpr int W0 = 40, H0 = 30;                       // level 0 of the REF pyramid: all points land here (worst spread)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
// footprint origin of point `pt` of (b,q,m): near the query's own pixel (locality like the encoder) + a random offset
__device__ __forceinline__ void origin(int q, int m, int pt, int spread, int& x0, int& y0) {
  const uint32_t h = hash32((uint32_t)(q * 16 + m) * 16u + pt);
  const int qx = q % W0, qy = (q / W0) % H0;
  x0 = min(max(qx + (int)(h % (2 * spread + 1)) - spread, 0), W0 - 2);
  y0 = min(max(qy + (int)((h >> 8) % (2 * spread + 1)) - spread, 0), H0 - 2);
}

template <bool HEAD_MAJOR>
__global__ void __launch_bounds__(256) k_red_v4(float* dst, int spread) {
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int c4 = t & 3;
  const int64_t bqm = t >> 2;
  int m, q, b;
  if (HEAD_MAJOR) { q = bqm % LQ; m = (bqm / LQ) % M; b = bqm / (LQ * M); }      // a warp = 8 consecutive queries of one head
  else { m = bqm % M; q = (bqm / M) % LQ; b = bqm / (M * LQ); }                  // a warp = 8 heads of one query (product mapping)
  const int64_t pstride = HEAD_MAJOR ? D : M * D;
  float* base = dst + (HEAD_MAJOR ? ((int64_t)(b * M + m) * S) * D : ((int64_t)b * S * M + m) * D) + c4 * 4;
  const float v = 1e-6f * (float)c4;
#pragma unroll 4
  for (int pt = 0; pt < LP; ++pt) {
    int x0, y0;
    origin(q, m, pt, spread, x0, y0);
    float* p00 = base + (int64_t)(y0 * W0 + x0) * pstride;
    asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p00), "f"(v));
    asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p00 + pstride), "f"(v));
    asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p00 + W0 * pstride), "f"(v));
    asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p00 + (W0 + 1) * pstride), "f"(v));
  }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one lane per (b,q,m); bulk reductions of 64 B (row-major, 4 per point) or 128 B (head-major, 2 per point)
template <bool HEAD_MAJOR>
__global__ void __launch_bounds__(256) k_bulk_1d(float* dst, int spread) {
  __shared__ __align__(128) float stage[256 * 32];                     // 128 B per thread (contents irrelevant, constant)
  for (int i = threadIdx.x; i < 256 * 32; i += 256) stage[i] = 1e-6f;
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const int64_t bqm = (int64_t)blockIdx.x * 256 + threadIdx.x;
  int m, q, b;
  if (HEAD_MAJOR) { q = bqm % LQ; m = (bqm / LQ) % M; b = bqm / (LQ * M); }
  else { m = bqm % M; q = (bqm / M) % LQ; b = bqm / (M * LQ); }
  const int64_t pstride = HEAD_MAJOR ? D : M * D;
  float* base = dst + (HEAD_MAJOR ? ((int64_t)(b * M + m) * S) * D : ((int64_t)b * S * M + m) * D);
  const uint32_t src = smem_u32(stage + threadIdx.x * 32);
  for (int pt = 0; pt < LP; ++pt) {
    int x0, y0;
    origin(q, m, pt, spread, x0, y0);
    float* p00 = base + (int64_t)(y0 * W0 + x0) * pstride;
    if (HEAD_MAJOR) {
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 128;" ::"l"(p00), "r"(src) : "memory");
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 128;" ::"l"(p00 + W0 * pstride), "r"(src) : "memory");
    } else {
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 64;" ::"l"(p00), "r"(src) : "memory");
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 64;" ::"l"(p00 + pstride), "r"(src) : "memory");
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 64;" ::"l"(p00 + W0 * pstride), "r"(src) : "memory");
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 64;" ::"l"(p00 + (W0 + 1) * pstride), "r"(src) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// one lane per (b,q,m); ONE 5-d tensor reduction per point: box (16 d, 1 m, 2 x, 2 y, 1 b) = 256 B
__global__ void __launch_bounds__(128) k_bulk_tensor(const __grid_constant__ CUtensorMap tm, int spread) {
  __shared__ __align__(128) float stage[128 * 64];                     // 256 B per thread
  for (int i = threadIdx.x; i < 128 * 64; i += 128) stage[i] = 1e-6f;
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const int64_t bqm = (int64_t)blockIdx.x * 128 + threadIdx.x;
  const int m = bqm % M, q = (bqm / M) % LQ, b = bqm / (M * LQ);
  const uint32_t src = smem_u32(stage + threadIdx.x * 64);
  for (int pt = 0; pt < LP; ++pt) {
    int x0, y0;
    origin(q, m, pt, spread, x0, y0);
    asm volatile("cp.reduce.async.bulk.tensor.5d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"(&tm), "r"(src), "r"(0), "r"(m), "r"(x0), "r"(y0), "r"(b) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename F>
static float time_us(F launch, int reps = 10) {
  cudaEvent_t s, e;
  cudaEventCreate(&s); cudaEventCreate(&e);
  launch(); launch();
  cudaDeviceSynchronize();
  cudaEventRecord(s);
  for (int i = 0; i < reps; ++i) launch();
  cudaEventRecord(e);
  cudaEventSynchronize(e);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, s, e);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(err));
  return ms * 1e3f / reps;
}

int main() {
  float* dst = nullptr;
  const size_t n = (size_t)B * S * M * D;
  cudaMalloc(&dst, n * 4);
  cudaMemset(dst, 0, n * 4);
  const int64_t groups = (int64_t)B * LQ * M;
  const double points = (double)groups * LP;
  EncodeTiledFn enc = nullptr;
  cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qr);
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  {   // level-0 view of [B,S,M,D]: dims (d, m, x, y, b)
    cuuint64_t dims[5] = {D, M, W0, H0, B};
    cuuint64_t strides[4] = {D * 4, (cuuint64_t)M * D * 4, (cuuint64_t)W0 * M * D * 4, (cuuint64_t)S * M * D * 4};
    cuuint32_t box[5] = {D, 1, 2, 2, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, dst, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("tensor map encode failed: %d\n", (int)r);
  }
  for (int spread : {1, 4, 20}) {
    printf("spread +-%d px around the query's pixel (level 0, %dx%d), %.2f M points, 4 corners x 64 B each\n", spread, W0, H0, points / 1e6);
    float us;
    us = time_us([&] { k_red_v4<false><<<(unsigned)(groups * 4 / 256), 256>>>(dst, spread); });
    printf("  A red.v4 row-major           %8.1f us  %6.2f ns/point  %6.1f GB/s payload\n", us, us * 1e3 / points, points * 256 / us / 1e3);
    us = time_us([&] { k_red_v4<true><<<(unsigned)(groups * 4 / 256), 256>>>(dst, spread); });
    printf("  B red.v4 head-major          %8.1f us  %6.2f ns/point  %6.1f GB/s payload\n", us, us * 1e3 / points, points * 256 / us / 1e3);
    us = time_us([&] { k_bulk_1d<false><<<(unsigned)(groups / 256), 256>>>(dst, spread); });
    printf("  C bulk 64 B row-major        %8.1f us  %6.2f ns/point  %6.1f GB/s payload\n", us, us * 1e3 / points, points * 256 / us / 1e3);
    us = time_us([&] { k_bulk_1d<true><<<(unsigned)(groups / 256), 256>>>(dst, spread); });
    printf("  D bulk 128 B head-major      %8.1f us  %6.2f ns/point  %6.1f GB/s payload\n", us, us * 1e3 / points, points * 256 / us / 1e3);
    us = time_us([&] { k_bulk_tensor<<<(unsigned)(groups / 128), 128>>>(tm, spread); });
    printf("  E bulk tensor 5-d 2x2x64 B   %8.1f us  %6.2f ns/point  %6.1f GB/s payload\n", us, us * 1e3 / points, points * 256 / us / 1e3);
  }
  cudaFree(dst);
  return 0;
}
