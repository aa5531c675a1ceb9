// Pose heads tail: class-specific slot select + 6D -> SO(3) Gram-Schmidt, forward and backward.
// Replaces the Python list-comprehension gather over B*Q rows and rotation_6d_to_matrix of the
// reference (models/pose_estimation_transformer.py:354, 365-374, 434-451).  The MLPs in front are
// poet_gemm calls.  F.normalize semantics: v / max(||v||, 1e-12).
#include "common.cuh"

namespace {

__device__ __forceinline__ void cross3(const float* a, const float* b, float* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

constexpr float kNormEps = 1e-12f;

__global__ void __launch_bounds__(128) heads_fwd_kernel(const float* __restrict__ rot_all, const float* __restrict__ trans_all,
                                                        const int64_t* __restrict__ classes, float* __restrict__ trans,
                                                        float* __restrict__ rot6d, float* __restrict__ rotmat, int R, int n_slots) {
  poet_pdl_entry();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  int slot = 0;
  if (n_slots > 1) { const int64_t c = classes[r]; slot = c > 0 ? (int)c : 0; if (slot >= n_slots) slot = n_slots - 1; }
  const float* tp = trans_all + ((int64_t)r * n_slots + slot) * 3;
  const float* rp = rot_all + ((int64_t)r * n_slots + slot) * 6;
  float a[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) { a[i] = rp[i]; rot6d[(int64_t)r * 6 + i] = a[i]; }
#pragma unroll
  for (int i = 0; i < 3; ++i) trans[(int64_t)r * 3 + i] = tp[i];
  float x[3], zu[3], z[3], y[3];
  const float n1 = fmaxf(sqrtf(dot3(a, a)), kNormEps);
#pragma unroll
  for (int i = 0; i < 3; ++i) x[i] = a[i] / n1;
  cross3(x, a + 3, zu);
  const float n2 = fmaxf(sqrtf(dot3(zu, zu)), kNormEps);
#pragma unroll
  for (int i = 0; i < 3; ++i) z[i] = zu[i] / n2;
  cross3(z, x, y);
  float* o = rotmat + (int64_t)r * 9;
#pragma unroll
  for (int i = 0; i < 3; ++i) { o[i * 3 + 0] = x[i]; o[i * 3 + 1] = y[i]; o[i * 3 + 2] = z[i]; }
}

// one block per row: thread 0 differentiates Gram-Schmidt, the block writes the (mostly zero) rows.
__global__ void __launch_bounds__(128) heads_bwd_kernel(const float* __restrict__ rot6d, const int64_t* __restrict__ classes,
                                                        const float* __restrict__ g_trans, const float* __restrict__ g_rotmat,
                                                        float* __restrict__ g_rot_all, float* __restrict__ g_trans_all,
                                                        int R, int n_slots) {
  poet_pdl_entry();
  __shared__ float s_g[9];
  __shared__ int s_slot;
  const int r = blockIdx.x;
  if (threadIdx.x == 0) {
    int slot = 0;
    if (n_slots > 1) { const int64_t c = classes[r]; slot = c > 0 ? (int)c : 0; if (slot >= n_slots) slot = n_slots - 1; }
    s_slot = slot;
    float a[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) a[i] = rot6d[(int64_t)r * 6 + i];
    float x[3], zu[3], z[3];
    const float l1 = sqrtf(dot3(a, a)), n1 = fmaxf(l1, kNormEps);
#pragma unroll
    for (int i = 0; i < 3; ++i) x[i] = a[i] / n1;
    cross3(x, a + 3, zu);
    const float l2 = sqrtf(dot3(zu, zu)), n2 = fmaxf(l2, kNormEps);
#pragma unroll
    for (int i = 0; i < 3; ++i) z[i] = zu[i] / n2;
    const float* g = g_rotmat + (int64_t)r * 9;
    float gx[3], gy[3], gz[3], t[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { gx[i] = g[i * 3 + 0]; gy[i] = g[i * 3 + 1]; gz[i] = g[i * 3 + 2]; }
    // y = z x x
    cross3(x, gy, t);
#pragma unroll
    for (int i = 0; i < 3; ++i) gz[i] += t[i];
    cross3(gy, z, t);
#pragma unroll
    for (int i = 0; i < 3; ++i) gx[i] += t[i];
    // z = zu / max(|zu|, eps)
    float gzu[3];
    if (l2 > kNormEps) { const float d = dot3(z, gz);
#pragma unroll
      for (int i = 0; i < 3; ++i) gzu[i] = (gz[i] - z[i] * d) / n2;
    } else {
#pragma unroll
      for (int i = 0; i < 3; ++i) gzu[i] = gz[i] / n2;
    }
    // zu = x x a2
    float ga2[3];
    cross3(a + 3, gzu, t);
#pragma unroll
    for (int i = 0; i < 3; ++i) gx[i] += t[i];
    cross3(gzu, x, ga2);
    // x = a1 / max(|a1|, eps)
    float ga1[3];
    if (l1 > kNormEps) { const float d = dot3(x, gx);
#pragma unroll
      for (int i = 0; i < 3; ++i) ga1[i] = (gx[i] - x[i] * d) / n1;
    } else {
#pragma unroll
      for (int i = 0; i < 3; ++i) ga1[i] = gx[i] / n1;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) { s_g[i] = ga1[i]; s_g[3 + i] = ga2[i]; s_g[6 + i] = g_trans[(int64_t)r * 3 + i]; }
  }
  __syncthreads();
  const int slot = s_slot;
  for (int i = threadIdx.x; i < n_slots * 6; i += blockDim.x)
    g_rot_all[(int64_t)r * n_slots * 6 + i] = (i / 6 == slot) ? s_g[i % 6] : 0.f;
  for (int i = threadIdx.x; i < n_slots * 3; i += blockDim.x)
    g_trans_all[(int64_t)r * n_slots * 3 + i] = (i / 3 == slot) ? s_g[6 + i % 3] : 0.f;
}

}  // namespace

extern "C" int poet_heads_select_rot6d_fwd(const float* rot_all, const float* trans_all, const int64_t* classes,
                                           float* trans, float* rot6d, float* rotmat, int R, int n_slots,
                                           poet_stream_t stream) {
  POET_REQUIRE(rot_all && trans_all && trans && rot6d && rotmat, POET_ERR_NULL_POINTER);
  POET_REQUIRE(n_slots == 1 || classes != nullptr, POET_ERR_NULL_POINTER);
  POET_REQUIRE(R > 0 && n_slots >= 1, POET_ERR_BAD_SHAPE);
  poet_launch(heads_fwd_kernel, dim3(poet_ceil_div(R, 128)), dim3(128), 0, (cudaStream_t)stream, rot_all, trans_all, classes, trans, rot6d,
                                                                            rotmat, R, n_slots);
  return poet_launch_status();
}

extern "C" int poet_heads_select_rot6d_bwd(const float* rot6d, const int64_t* classes, const float* grad_trans,
                                           const float* grad_rotmat, float* grad_rot_all, float* grad_trans_all, int R,
                                           int n_slots, poet_stream_t stream) {
  POET_REQUIRE(rot6d && grad_trans && grad_rotmat && grad_rot_all && grad_trans_all, POET_ERR_NULL_POINTER);
  POET_REQUIRE(n_slots == 1 || classes != nullptr, POET_ERR_NULL_POINTER);
  POET_REQUIRE(R > 0 && n_slots >= 1, POET_ERR_BAD_SHAPE);
  poet_launch(heads_bwd_kernel, dim3(R), dim3(128), 0, (cudaStream_t)stream, rot6d, classes, grad_trans, grad_rotmat, grad_rot_all,
                                                        grad_trans_all, R, n_slots);
  return poet_launch_status();
}
