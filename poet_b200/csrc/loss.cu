// Pose loss of PoET on the device: translation L2 + rotation geodesic distance for every decoder
// layer, forward value and gradient in one launch.
// Replaces SetCriterion.loss_translation / loss_rotation (reference models/pose_estimation_transformer.py:478-494,
// 519-537) applied to the final and the auxiliary outputs (:635-662) under a given query->target assignment
// (PoseMatcher, models/matcher.py:158-229; in bbox_mode 'gt' the assignment is query j <-> target j, :169-173),
// without the per-layer `C.cpu()` + scipy round trips.
#include "common.cuh"

namespace {

// one thread per (layer, image, query)
__global__ void __launch_bounds__(128) pose_loss_kernel(const float* __restrict__ pred_t, const float* __restrict__ pred_R,
                                                        const float* __restrict__ tgt_t, const float* __restrict__ tgt_R,
                                                        const int32_t* __restrict__ assign, const int32_t* __restrict__ n_obj,
                                                        float* __restrict__ losses, float* __restrict__ grad_t,
                                                        float* __restrict__ grad_R, int L, int B, int Q, int T, float w_trans,
                                                        float w_rot) {
  poet_pdl_entry();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float lt = 0.f, lr = 0.f;
  const int l = i / (B * Q);
  if (i < L * B * Q) {
    const int bq = i % (B * Q), b = bq / Q;
    const int j = assign[bq];
    const float inv_n = 1.f / (float)max(*n_obj, 1);
    float gt[3] = {0.f, 0.f, 0.f}, gR[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (j >= 0) {
      const float* pt = pred_t + (int64_t)i * 3;
      const float* pR = pred_R + (int64_t)i * 9;
      const float* tt = tgt_t + ((int64_t)b * T + j) * 3;
      const float* tR = tgt_R + ((int64_t)b * T + j) * 9;
      const float d0 = pt[0] - tt[0], d1 = pt[1] - tt[1], d2 = pt[2] - tt[2];
      const float nrm = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);                 // sqrt(sum(mse)) :486-488
      lt = nrm * inv_n;
      if (nrm > 0.f) { const float s = w_trans * inv_n / nrm; gt[0] = s * d0; gt[1] = s * d1; gt[2] = s * d2; }
      float trace = 0.f;                                                    // trace(R_pred R_gt^T) = <R_pred, R_gt>
#pragma unroll
      for (int k = 0; k < 9; ++k) trace += pR[k] * tR[k];
      const float raw = 0.5f * (trace - 1.f);
      const float theta = fminf(fmaxf(raw, -1.f + 1e-6f), 1.f - 1e-6f);    // :533
      lr = acosf(theta) * inv_n;
      if (raw > -1.f + 1e-6f && raw < 1.f - 1e-6f) {                        // clamp passes no gradient outside
        const float s = -w_rot * inv_n * 0.5f * rsqrtf(1.f - theta * theta);
#pragma unroll
        for (int k = 0; k < 9; ++k) gR[k] = s * tR[k];
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) grad_t[(int64_t)i * 3 + k] = gt[k];
#pragma unroll
    for (int k = 0; k < 9; ++k) grad_R[(int64_t)i * 9 + k] = gR[k];
  }
  // a block never straddles two layers when B*Q % blockDim == 0; in general reduce per warp and let lanes of the
  // same layer combine
  const unsigned full = 0xffffffffu;
  const int l0 = __shfl_sync(full, l, 0);
  const bool uniform = __all_sync(full, l == l0);
  if (uniform) {
    lt = warp_sum(lt); lr = warp_sum(lr);
    if ((threadIdx.x & 31) == 0 && l0 < L) { atomicAdd(losses + l0 * 2, lt); atomicAdd(losses + l0 * 2 + 1, lr); }
  } else if (i < L * B * Q) {
    atomicAdd(losses + l * 2, lt); atomicAdd(losses + l * 2 + 1, lr);
  }
}

}  // namespace

extern "C" int poet_pose_loss(const float* pred_t, const float* pred_R, const float* tgt_t, const float* tgt_R,
                              const int32_t* assign, const int32_t* n_obj, float* losses, float* grad_t, float* grad_R,
                              int L, int B, int Q, int T, float w_trans, float w_rot, poet_stream_t stream) {
  POET_REQUIRE(pred_t && pred_R && tgt_t && tgt_R && assign && n_obj && losses && grad_t && grad_R, POET_ERR_NULL_POINTER);
  POET_REQUIRE(L > 0 && B > 0 && Q > 0 && T > 0, POET_ERR_BAD_SHAPE);
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(losses, 0, sizeof(float) * 2 * L, s);
  if (e != cudaSuccess) return (int)e;
  poet_launch(pose_loss_kernel, dim3(poet_ceil_div((int64_t)L * B * Q, 128)), dim3(128), 0, s, pred_t, pred_R, tgt_t, tgt_R,
              assign, n_obj, losses, grad_t, grad_R, L, B, Q, T, w_trans, w_rot);
  return poet_launch_status();
}
