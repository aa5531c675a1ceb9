// Position encodings and token-layout plumbing.
//   poet_posenc_sine          models/position_encoding.py:40-60
//   poet_bbox_embed_pad       models/position_encoding.py:71-84 + pose_estimation_transformer.py:217-236
//   poet_nchw_to_tokens/back  models/deformable_transformer.py:124-140
//   poet_enc_reference_points models/deformable_transformer.py:217-230
// Accurate sinf/cosf are required (bbox arguments reach 2^31): this file must never be compiled
// with -use_fast_math.  Arithmetic is ordered exactly like the reference's fp32 tensor ops.
#include "common.cuh"

namespace {

constexpr int kPix = 32;   // pixels per block

// e = cumulative count of valid pixels along one axis, normalised like position_encoding.py:47-50
__global__ void __launch_bounds__(256) posenc_kernel(const uint8_t* __restrict__ mask, const float* __restrict__ dim_t,
                                                     const float* __restrict__ level_embed, float* __restrict__ out,
                                                     int B, int H, int W, int F, float scale, int normalize, int layout,
                                                     int S_total, int row_offset) {
  poet_pdl_entry();
  __shared__ float s_ey[kPix], s_ex[kPix];
  const int HW = H * W;
  const int64_t pix0 = (int64_t)blockIdx.x * kPix;           // over B*HW
  const int64_t npix = (int64_t)B * HW;
  if (threadIdx.x < kPix) {
    const int64_t pix = pix0 + threadIdx.x;
    float ey = 0.f, ex = 0.f;
    if (pix < npix) {
      const int b = (int)(pix / HW), yx = (int)(pix % HW), y = yx / W, x = yx % W;
      const uint8_t* mb = mask + (int64_t)b * HW;
      int cy = 0, ty = 0, cx = 0, tx = 0;
      for (int i = 0; i < H; ++i) { int v = mb[i * W + x] == 0; ty += v; if (i <= y) cy += v; }
      for (int j = 0; j < W; ++j) { int v = mb[y * W + j] == 0; tx += v; if (j <= x) cx += v; }
      ey = (float)cy; ex = (float)cx;
      if (normalize) {
        ey = __fmul_rn(__fdiv_rn(__fsub_rn(ey, 0.5f), __fadd_rn((float)ty, 1e-6f)), scale);
        ex = __fmul_rn(__fdiv_rn(__fsub_rn(ex, 0.5f), __fadd_rn((float)tx, 1e-6f)), scale);
      }
    }
    s_ey[threadIdx.x] = ey; s_ex[threadIdx.x] = ex;
  }
  __syncthreads();
  const int C = 2 * F;
  if (layout == 1) {
    // token-major: consecutive threads write consecutive channels of one pixel (float4 each)
    const int q = C / 4;
    for (int i = threadIdx.x; i < kPix * q; i += blockDim.x) {
      const int pl = i / q, c = (i % q) * 4;
      const int64_t pix = pix0 + pl;
      if (pix >= npix) break;
      const int b = (int)(pix / HW), yx = (int)(pix % HW);
      const float e = c < F ? s_ey[pl] : s_ex[pl];
      const int k = c < F ? c : c - F;
      float4 v;
      v.x = sinf(__fdiv_rn(e, __ldg(dim_t + k)));
      v.y = cosf(__fdiv_rn(e, __ldg(dim_t + k + 1)));
      v.z = sinf(__fdiv_rn(e, __ldg(dim_t + k + 2)));
      v.w = cosf(__fdiv_rn(e, __ldg(dim_t + k + 3)));
      if (level_embed) { float4 le = ldg4(level_embed + c); v.x += le.x; v.y += le.y; v.z += le.z; v.w += le.w; }
      st4(out + ((int64_t)b * S_total + row_offset + yx) * C + c, v);
    }
  } else {
    // NCHW: consecutive threads write consecutive pixels of one channel
    for (int i = threadIdx.x; i < kPix * C; i += blockDim.x) {
      const int pl = i % kPix, c = i / kPix;
      const int64_t pix = pix0 + pl;
      if (pix >= npix) continue;
      const int b = (int)(pix / HW), yx = (int)(pix % HW);
      const float e = c < F ? s_ey[pl] : s_ex[pl];
      const int k = c < F ? c : c - F;
      const float arg = __fdiv_rn(e, __ldg(dim_t + k));
      out[((int64_t)b * C + c) * HW + yx] = (k & 1) ? cosf(arg) : sinf(arg);
    }
  }
}

__global__ void __launch_bounds__(256) bbox_embed_kernel(const float* __restrict__ boxes, const int32_t* __restrict__ n_boxes,
                                                         float* __restrict__ out, int B, int Q, int F) {
  poet_pdl_entry();
  const int C = 8 * F;
  const int64_t total = (int64_t)B * Q * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t bq = i / C;
    const int q = (int)(bq % Q), b = (int)(bq / Q);
    float v = -10.f;
    if (q < __ldg(n_boxes + b)) {
      const int coord = c / (2 * F), within = c % (2 * F), k = within % F;
      const float arg = __fmul_rn(__ldg(boxes + bq * 4 + coord), scalbnf(1.f, k));
      v = within < F ? sinf(arg) : cosf(arg);
    }
    out[bq * 2 * C + c] = v;            // query_pos half
    out[bq * 2 * C + C + c] = v;        // tgt half (query_embed.repeat(1, 2))
  }
}

// 32x32 smem transpose tiles: src [B][C][HW]  <->  tokens [B][S_total][C] rows row_offset..row_offset+HW
__global__ void __launch_bounds__(256) nchw_to_tokens_kernel(const float* __restrict__ src, const float* __restrict__ add_vec,
                                                             float* __restrict__ tokens, int C, int HW, int S_total, int row_offset) {
  poet_pdl_entry();
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, pp = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && pp < HW) ? __ldg(src + ((int64_t)b * C + c) * HW + pp) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int pp = p0 + i, c = c0 + threadIdx.x;
    if (pp < HW && c < C) {
      float v = tile[threadIdx.x][i];
      if (add_vec) v += __ldg(add_vec + c);
      tokens[((int64_t)b * S_total + row_offset + pp) * C + c] = v;
    }
  }
}

__global__ void __launch_bounds__(256) tokens_to_nchw_kernel(const float* __restrict__ gtok, float* __restrict__ gsrc,
                                                             float* __restrict__ gvec, int C, int HW, int S_total, int row_offset) {
  poet_pdl_entry();
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int pp = p0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (pp < HW && c < C) ? __ldg(gtok + ((int64_t)b * S_total + row_offset + pp) * C + c) : 0.f;
  }
  __syncthreads();
  if (gsrc)
    for (int i = threadIdx.y; i < 32; i += 8) {
      const int c = c0 + i, pp = p0 + threadIdx.x;
      if (c < C && pp < HW) gsrc[((int64_t)b * C + c) * HW + pp] = tile[threadIdx.x][i];
    }
  if (gvec && threadIdx.y == 0 && c0 + threadIdx.x < C) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += tile[i][threadIdx.x];
    atomicAdd(gvec + c0 + threadIdx.x, s);
  }
}

struct RefLevels { int H[4]; int W[4]; int start[4]; int L; int S; };

__global__ void __launch_bounds__(256) enc_ref_kernel(const float* __restrict__ vr, float* __restrict__ out, RefLevels lv, int B) {
  poet_pdl_entry();
  const int64_t total = (int64_t)B * lv.S;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / lv.S), s = (int)(i % lv.S);
    int l = 0;
    while (l + 1 < lv.L && s >= lv.start[l + 1]) ++l;
    const int local = s - lv.start[l], y = local / lv.W[l], x = local % lv.W[l];
    const float* v = vr + (int64_t)b * lv.L * 2;
    const float rx = __fdiv_rn((float)x + 0.5f, __fmul_rn(v[l * 2 + 0], (float)lv.W[l]));
    const float ry = __fdiv_rn((float)y + 0.5f, __fmul_rn(v[l * 2 + 1], (float)lv.H[l]));
    for (int k = 0; k < lv.L; ++k) {
      out[(i * lv.L + k) * 2 + 0] = __fmul_rn(rx, v[k * 2 + 0]);
      out[(i * lv.L + k) * 2 + 1] = __fmul_rn(ry, v[k * 2 + 1]);
    }
  }
}

}  // namespace

extern "C" int poet_posenc_sine(const uint8_t* mask, const float* dim_t, const float* level_embed, float* out, int B,
                                int H, int W, int F, float scale, int normalize, int layout, int S_total,
                                int row_offset, poet_stream_t stream) {
  POET_REQUIRE(mask && dim_t && out, POET_ERR_NULL_POINTER);
  POET_REQUIRE(B > 0 && H > 0 && W > 0 && F > 0 && F % 4 == 0, POET_ERR_BAD_SHAPE);
  POET_REQUIRE(layout == 0 || layout == 1, POET_ERR_UNSUPPORTED);
  if (layout == 1) {
    POET_REQUIRE(row_offset >= 0 && row_offset + H * W <= S_total, POET_ERR_BAD_SHAPE);
    POET_REQUIRE(poet_aligned16(out) && (!level_embed || poet_aligned16(level_embed)), POET_ERR_BAD_ALIGNMENT);
  }
  const int grid = poet_ceil_div((int64_t)B * H * W, kPix);
  poet_launch(posenc_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, mask, dim_t, level_embed, out, B, H, W, F, scale, normalize,
                                                        layout, S_total, row_offset);
  return poet_launch_status();
}

extern "C" int poet_bbox_embed_pad(const float* boxes, const int32_t* n_boxes, float* query_embeds, int B, int Q, int F,
                                   poet_stream_t stream) {
  POET_REQUIRE(boxes && n_boxes && query_embeds, POET_ERR_NULL_POINTER);
  POET_REQUIRE(B > 0 && Q > 0 && F > 0 && F <= 32, POET_ERR_BAD_SHAPE);
  const int64_t total = (int64_t)B * Q * 8 * F;
  int grid = poet_ceil_div(total, 256);
  if (grid > POET_NUM_SMS * 8) grid = POET_NUM_SMS * 8;
  poet_launch(bbox_embed_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, boxes, n_boxes, query_embeds, B, Q, F);
  return poet_launch_status();
}

extern "C" int poet_nchw_to_tokens(const float* src, const float* add_vec, float* tokens, int B, int C, int HW,
                                   int S_total, int row_offset, poet_stream_t stream) {
  POET_REQUIRE(src && tokens, POET_ERR_NULL_POINTER);
  POET_REQUIRE(B > 0 && C > 0 && HW > 0 && row_offset >= 0 && row_offset + HW <= S_total, POET_ERR_BAD_SHAPE);
  dim3 grid(poet_ceil_div(HW, 32), poet_ceil_div(C, 32), B);
  poet_launch(nchw_to_tokens_kernel, dim3(grid), dim3(32, 8), 0, (cudaStream_t)stream, src, add_vec, tokens, C, HW, S_total, row_offset);
  return poet_launch_status();
}

extern "C" int poet_tokens_to_nchw(const float* grad_tokens, float* grad_src, float* grad_vec, int B, int C, int HW,
                                   int S_total, int row_offset, poet_stream_t stream) {
  POET_REQUIRE(grad_tokens && (grad_src || grad_vec), POET_ERR_NULL_POINTER);
  POET_REQUIRE(B > 0 && C > 0 && HW > 0 && row_offset >= 0 && row_offset + HW <= S_total, POET_ERR_BAD_SHAPE);
  dim3 grid(poet_ceil_div(HW, 32), poet_ceil_div(C, 32), B);
  poet_launch(tokens_to_nchw_kernel, dim3(grid), dim3(32, 8), 0, (cudaStream_t)stream, grad_tokens, grad_src, grad_vec, C, HW, S_total,
                                                                        row_offset);
  return poet_launch_status();
}

// ---- A6: padding mask tokens + valid ratios in one launch ------------------------------------------------
// Replaces torch.cat([m.flatten(1) ...]) and the per-level get_valid_ratio of deformable_transformer.py:111-118,
// 126-141 (about 45 tiny ATen launches per forward).  One block per (image, level).
struct MaskPrepArgs { const uint8_t* mask[4]; int H[4], W[4], start[4]; };
__global__ void __launch_bounds__(256) mask_prep_kernel(const MaskPrepArgs a, uint8_t* __restrict__ pad,
                                                        float* __restrict__ valid_ratios, int L, int S) {
  poet_pdl_entry();
  const int b = blockIdx.x / L, l = blockIdx.x % L;
  const int H = a.H[l], W = a.W[l];
  const uint8_t* m = a.mask[l] + (int64_t)b * H * W;
  uint8_t* dst = pad + (int64_t)b * S + a.start[l];
  int cnt_h = 0, cnt_w = 0;                                   // unpadded rows of column 0 / unpadded columns of row 0
  for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
    const uint8_t v = m[i];
    dst[i] = v;
    if (v == 0) { cnt_h += (i % W) == 0; cnt_w += i < W; }
  }
  __shared__ int s_h, s_w;
  if (threadIdx.x == 0) { s_h = 0; s_w = 0; }
  __syncthreads();
  if (cnt_h) atomicAdd(&s_h, cnt_h);
  if (cnt_w) atomicAdd(&s_w, cnt_w);
  __syncthreads();
  if (threadIdx.x == 0) {
    float* vr = valid_ratios + ((int64_t)b * L + l) * 2;
    vr[0] = (float)s_w / (float)W;                             // (w, h) order: deformable_transformer.py:117
    vr[1] = (float)s_h / (float)H;
  }
}

extern "C" int poet_mask_prep(const uint8_t* const* masks_host, const int32_t* shapes_host, uint8_t* pad,
                              float* valid_ratios, int B, int L, poet_stream_t stream) {
  POET_REQUIRE(masks_host && shapes_host && pad && valid_ratios, POET_ERR_NULL_POINTER);
  POET_REQUIRE(B > 0 && L >= 1 && L <= 4, POET_ERR_BAD_SHAPE);
  MaskPrepArgs a;
  int S = 0;
  for (int l = 0; l < L; ++l) {
    POET_REQUIRE(masks_host[l] != nullptr, POET_ERR_NULL_POINTER);
    a.mask[l] = masks_host[l]; a.H[l] = shapes_host[2 * l]; a.W[l] = shapes_host[2 * l + 1]; a.start[l] = S;
    POET_REQUIRE(a.H[l] > 0 && a.W[l] > 0, POET_ERR_BAD_SHAPE);
    S += a.H[l] * a.W[l];
  }
  poet_launch(mask_prep_kernel, dim3(B * L), dim3(256), 0, (cudaStream_t)stream, a, pad, valid_ratios, L, S);
  return poet_launch_status();
}

extern "C" int poet_enc_reference_points(const float* valid_ratios, float* out, const int32_t* shapes_host, int B, int L,
                                         poet_stream_t stream) {
  POET_REQUIRE(valid_ratios && out && shapes_host, POET_ERR_NULL_POINTER);
  POET_REQUIRE(B > 0 && L >= 1 && L <= 4, POET_ERR_BAD_SHAPE);
  RefLevels lv{};
  lv.L = L;
  int start = 0;
  for (int l = 0; l < L; ++l) {
    lv.H[l] = shapes_host[2 * l]; lv.W[l] = shapes_host[2 * l + 1]; lv.start[l] = start;
    POET_REQUIRE(lv.H[l] > 0 && lv.W[l] > 0, POET_ERR_BAD_SHAPE);
    start += lv.H[l] * lv.W[l];
  }
  lv.S = start;
  int grid = poet_ceil_div((int64_t)B * start, 256);
  if (grid > POET_NUM_SMS * 8) grid = POET_NUM_SMS * 8;
  poet_launch(enc_ref_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, valid_ratios, out, lv, B);
  return poet_launch_status();
}
