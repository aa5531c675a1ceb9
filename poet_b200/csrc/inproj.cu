// input_proj of PoET: per level a 1x1 conv (or 3x3 / stride 2 for the extra level) + GroupNorm(32),
// reference models/pose_estimation_transformer.py:100-135 (definition) and :313-335 (use).
// The convolutions are GEMMs on the tensor-core path (poet_gemm); this file holds what surrounds them:
//   poet_im2col_3x3s2        NCHW -> [B*Ho*Wo, C*9] patch matrix (zero padding 1, stride 2), k = c*9 + ky*3 + kx
//   poet_groupnorm_tokens_*  GroupNorm over token-major rows [B*HW, C] (group = C/G consecutive channels of one
//                            image), writing straight into the level's slice of the transformer's token matrix
//                            [B, S_total, C]: the NCHW intermediate of the reference never exists.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) im2col_3x3s2_kernel(const float* __restrict__ x, float* __restrict__ col, int B, int C,
                                                           int H, int W, int Ho, int Wo) {
  poet_pdl_entry();
  // one thread per (row = b,oy,ox ; channel c): writes 9 consecutive floats
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * Ho * Wo * C;
  if (t >= total) return;
  const int c = (int)(t % C);
  const int64_t row = t / C;
  const int ox = (int)(row % Wo), oy = (int)((row / Wo) % Ho), b = (int)(row / ((int64_t)Wo * Ho));
  const float* xc = x + ((int64_t)b * C + c) * H * W;
  float* dst = col + row * ((int64_t)C * 9) + c * 9;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * 2 - 1 + ky;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox * 2 - 1 + kx;
      dst[ky * 3 + kx] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(xc + iy * W + ix) : 0.f;
    }
  }
}

// ---- GroupNorm on token rows -------------------------------------------------------------------------
// stats[b, g] = {sum, sumsq} (double) over the HW rows x CPG channels of image b, group g.
// grid (row chunks, B); block = C threads (C <= 1024): thread = channel, loops over the chunk's rows.
__global__ void gn_stats_kernel(const float* __restrict__ y, double* __restrict__ stats, int HW, int C, int CPG, int rows_per_block) {
  poet_pdl_entry();
  const int b = blockIdx.y, c = threadIdx.x;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(HW, r0 + rows_per_block);
  float s = 0.f, ss = 0.f;
  for (int r = r0; r < r1; ++r) {
    const float v = __ldg(y + ((int64_t)b * HW + r) * C + c);
    s += v; ss += v * v;
  }
  // fold the CPG channels of a group (CPG is a power of two <= 32, groups are lane-aligned)
  for (int o = CPG >> 1; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); ss += __shfl_xor_sync(0xffffffffu, ss, o); }
  if ((c % CPG) == 0) {
    double* st = stats + ((int64_t)b * (C / CPG) + c / CPG) * 2;
    atomicAdd(st, (double)s);
    atomicAdd(st + 1, (double)ss);
  }
}

// out[(b*S_total + row_offset + r), c] = (y - mean) * rstd * gamma[c] + beta[c]
__global__ void gn_apply_kernel(const float* __restrict__ y, const double* __restrict__ stats, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float* __restrict__ out, int HW, int C, int CPG, int S_total,
                                int row_offset, float eps, int rows_per_block) {
  poet_pdl_entry();
  const int b = blockIdx.y, c = threadIdx.x;
  const double* st = stats + ((int64_t)b * (C / CPG) + c / CPG) * 2;
  const double n = (double)HW * CPG;
  const double mean = st[0] / n;
  const float rstd = rsqrtf((float)fmax(st[1] / n - mean * mean, 0.0) + eps);
  const float g = gamma[c] * rstd, sh = beta[c] - (float)mean * g;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(HW, r0 + rows_per_block);
  for (int r = r0; r < r1; ++r)
    out[((int64_t)b * S_total + row_offset + r) * C + c] = __ldg(y + ((int64_t)b * HW + r) * C + c) * g + sh;
}

// backward statistics: per (b, group): sum of dy*gamma and of dy*gamma*xhat; per channel dgamma += sum dy*xhat, dbeta += sum dy
__global__ void gn_bwd_stats_kernel(const float* __restrict__ dy, const float* __restrict__ y, const double* __restrict__ stats,
                                    const float* __restrict__ gamma, double* __restrict__ bstats, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta, int HW, int C, int CPG, int S_total, int row_offset, float eps,
                                    int rows_per_block) {
  poet_pdl_entry();
  const int b = blockIdx.y, c = threadIdx.x;
  const double* st = stats + ((int64_t)b * (C / CPG) + c / CPG) * 2;
  const double n = (double)HW * CPG;
  const double mean = st[0] / n;
  const float rstd = rsqrtf((float)fmax(st[1] / n - mean * mean, 0.0) + eps), mu = (float)mean;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(HW, r0 + rows_per_block);
  float sdy = 0.f, sdyx = 0.f;
  for (int r = r0; r < r1; ++r) {
    const float d = __ldg(dy + ((int64_t)b * S_total + row_offset + r) * C + c);
    const float xh = (__ldg(y + ((int64_t)b * HW + r) * C + c) - mu) * rstd;
    sdy += d; sdyx += d * xh;
  }
  atomicAdd(dgamma + c, sdyx);
  atomicAdd(dbeta + c, sdy);
  float g1 = sdy * gamma[c], g2 = sdyx * gamma[c];
  for (int o = CPG >> 1; o > 0; o >>= 1) { g1 += __shfl_xor_sync(0xffffffffu, g1, o); g2 += __shfl_xor_sync(0xffffffffu, g2, o); }
  if ((c % CPG) == 0) {
    double* bs = bstats + ((int64_t)b * (C / CPG) + c / CPG) * 2;
    atomicAdd(bs, (double)g1);
    atomicAdd(bs + 1, (double)g2);
  }
}

// dx = rstd * (dy*gamma - mean(dy*gamma) - xhat * mean(dy*gamma*xhat))
__global__ void gn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ y, const double* __restrict__ stats,
                                    const double* __restrict__ bstats, const float* __restrict__ gamma, float* __restrict__ dx,
                                    int HW, int C, int CPG, int S_total, int row_offset, float eps, int rows_per_block) {
  poet_pdl_entry();
  const int b = blockIdx.y, c = threadIdx.x;
  const int64_t gi = ((int64_t)b * (C / CPG) + c / CPG) * 2;
  const double n = (double)HW * CPG;
  const double mean = stats[gi] / n;
  const float rstd = rsqrtf((float)fmax(stats[gi + 1] / n - mean * mean, 0.0) + eps), mu = (float)mean;
  const float m1 = (float)(bstats[gi] / n), m2 = (float)(bstats[gi + 1] / n), gm = gamma[c];
  const int r0 = blockIdx.x * rows_per_block, r1 = min(HW, r0 + rows_per_block);
  for (int r = r0; r < r1; ++r) {
    const float d = __ldg(dy + ((int64_t)b * S_total + row_offset + r) * C + c);
    const float xh = (__ldg(y + ((int64_t)b * HW + r) * C + c) - mu) * rstd;
    dx[((int64_t)b * HW + r) * C + c] = rstd * (d * gm - m1 - xh * m2);
  }
}

int gn_check(int B, int HW, int C, int G) {
  POET_REQUIRE(B > 0 && HW > 0 && C > 0 && C <= 1024 && G > 0 && C % G == 0, POET_ERR_BAD_SHAPE);
  const int cpg = C / G;
  POET_REQUIRE(cpg <= 32 && (cpg & (cpg - 1)) == 0 && C % 32 == 0, POET_ERR_UNSUPPORTED);     // lane-aligned power-of-two groups
  return POET_OK;
}
constexpr int kGnRows = 16;

}  // namespace

extern "C" int poet_im2col_3x3s2(const float* x, float* col, int B, int C, int H, int W, poet_stream_t stream) {
  POET_REQUIRE(x && col, POET_ERR_NULL_POINTER);
  POET_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, POET_ERR_BAD_SHAPE);
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const int64_t total = (int64_t)B * Ho * Wo * C;
  poet_launch(im2col_3x3s2_kernel, dim3(poet_ceil_div(total, 256)), dim3(256), 0, (cudaStream_t)stream, x, col, B, C, H, W, Ho, Wo);
  return poet_launch_status();
}

extern "C" int poet_groupnorm_tokens_fwd(const float* y, const float* gamma, const float* beta, float* tokens, double* stats,
                                         int B, int HW, int C, int G, int S_total, int row_offset, float eps,
                                         poet_stream_t stream) {
  POET_REQUIRE(y && gamma && beta && tokens && stats, POET_ERR_NULL_POINTER);
  int rc = gn_check(B, HW, C, G);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(double) * 2 * B * G, s);
  if (e != cudaSuccess) return (int)e;
  const dim3 grid(poet_ceil_div(HW, kGnRows), B);
  poet_launch(gn_stats_kernel, grid, dim3(C), 0, s, y, stats, HW, C, C / G, kGnRows);
  poet_launch(gn_apply_kernel, grid, dim3(C), 0, s, y, (const double*)stats, gamma, beta, tokens, HW, C, C / G, S_total, row_offset,
              eps, kGnRows);
  return poet_launch_status();
}

extern "C" int poet_groupnorm_tokens_bwd(const float* grad_tokens, const float* y, const double* stats, const float* gamma,
                                         float* grad_y, float* dgamma, float* dbeta, double* workspace, int B, int HW, int C,
                                         int G, int S_total, int row_offset, float eps, poet_stream_t stream) {
  POET_REQUIRE(grad_tokens && y && stats && gamma && grad_y && dgamma && dbeta && workspace, POET_ERR_NULL_POINTER);
  int rc = gn_check(B, HW, C, G);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(workspace, 0, sizeof(double) * 2 * B * G, s);
  if (e != cudaSuccess) return (int)e;
  const dim3 grid(poet_ceil_div(HW, kGnRows), B);
  poet_launch(gn_bwd_stats_kernel, grid, dim3(C), 0, s, grad_tokens, y, stats, gamma, workspace, dgamma, dbeta, HW, C, C / G, S_total,
              row_offset, eps, kGnRows);
  poet_launch(gn_bwd_apply_kernel, grid, dim3(C), 0, s, grad_tokens, y, stats, (const double*)workspace, gamma, grad_y, HW, C, C / G,
              S_total, row_offset, eps, kGnRows);
  return poet_launch_status();
}
