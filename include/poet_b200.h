/* poet_b200.h — C ABI of libpoet_b200.so: the B200 (sm_100a) kernels of PoET's deformable
 * encoder/decoder hot path.
 *
 * Boundary contract (SURVEY.md §8 B-c):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in
 *     `_host`.  The caller (PyTorch, or any other host) allocates all inputs, outputs and
 *     workspaces; the library never allocates, frees or retains device memory, never
 *     synchronises the device and launches only on `stream`.
 *   - re-entrant, no global mutable state: the reference runs backward on the autograd worker
 *     thread and fires NCCL from hooks (reference main.py:282), so calls may come from any thread.
 *   - return value: 0 = ok, <0 = poet_status argument error, >0 = cudaError_t of the launch.
 *   - layouts are row-major contiguous, channels-last ([B,S,C]); float tensors are fp32 and
 *     16-byte aligned (vector loads / TMA).
 *
 * What each entry point replaces in the reference (file:line in aau-cns/poet):
 *   poet_posenc_sine            models/position_encoding.py:40-60   PositionEmbeddingSine.forward
 *   poet_bbox_embed_pad         models/position_encoding.py:71-84 + pose_estimation_transformer.py:203-239
 *   poet_nchw_to_tokens(_bwd)   models/deformable_transformer.py:124-140 (flatten/transpose/cat, + level_embed)
 *   poet_enc_reference_points   models/deformable_transformer.py:217-230 get_reference_points
 *   poet_msda_fwd / _bwd        deformable_attention.MSDeformAttn core (third-party ms_deform_attn_forward/backward,
 *                               called at deformable_transformer.py:201,283-285), optionally fused with the
 *                               softmax and loc = ref + off/(W,H) of the module
 *   poet_gemm                   every nn.Linear on the path (deformable_transformer.py:182-185,258-261,
 *                               pose_estimation_transformer.py:684) and their dgrad/wgrad
 *   poet_add_layernorm_fwd/bwd  residual + nn.LayerNorm (deformable_transformer.py:196-197,202-203,270-271,279-280,286-287)
 *   poet_mha_smallq_fwd/bwd     nn.MultiheadAttention core on Q<=32 rows (deformable_transformer.py:277-278)
 *   poet_heads_select_rot6d_*   class-specific select + rotation_6d_to_matrix (pose_estimation_transformer.py:354,365-374,434-451)
 */
#ifndef POET_B200_H_
#define POET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* poet_stream_t; /* cudaStream_t */

enum poet_status {
  POET_OK = 0,
  POET_ERR_BAD_SHAPE = -1,
  POET_ERR_BAD_ALIGNMENT = -2,
  POET_ERR_UNSUPPORTED = -3,
  POET_ERR_NULL_POINTER = -4,
  POET_ERR_WORKSPACE = -5,
  POET_ERR_WRONG_DEVICE = -6
};

/* ---- library ------------------------------------------------------------------------- */
int poet_version(void);                      /* 100*major + minor */
int poet_sm(void);                           /* 100: the only target this library is built for */
int poet_check_device(int device);           /* POET_OK iff `device` has compute capability 10.0 */
const char* poet_error_string(int code);     /* static string for any return value */

/* ---- A0: sinusoidal image position encoding ------------------------------------------- */
/* mask [B,H,W] uint8 (1 = padded).  dim_t [F] = temperature^(2*floor(i/2)/F).
 * layout 0: out is [B,2F,H,W] (reference layout).
 * layout 1: out is token-major: row (b*S_total + row_offset + y*W + x), 2F channels, plus
 *           level_embed[2F] if non-NULL (== lvl_pos_embed_flatten, deformable_transformer.py:133-135). */
int poet_posenc_sine(const uint8_t* mask, const float* dim_t, const float* level_embed, float* out,
                     int B, int H, int W, int F, float scale, int normalize, int layout,
                     int S_total, int row_offset, poet_stream_t stream);

/* ---- A1: bounding-box embedding with dummy padding -------------------------------------- */
/* boxes [B,Q,4] already padded with -1; n_boxes [B] int32.  query_embeds [B,Q,2C], C = 8*F:
 * real rows = [emb|emb], emb = per coord [sin(c*2^k) k<F | cos(c*2^k) k<F]; dummy rows = -10. */
int poet_bbox_embed_pad(const float* boxes, const int32_t* n_boxes, float* query_embeds,
                        int B, int Q, int F, poet_stream_t stream);

/* ---- A6: flatten one NCHW level into the token matrix (and back) ------------------------ */
/* tokens[(b*S_total + row_offset + hw), c] = src[b,c,hw] (+ add_vec[c] if non-NULL). */
int poet_nchw_to_tokens(const float* src, const float* add_vec, float* tokens, int B, int C, int HW,
                        int S_total, int row_offset, poet_stream_t stream);
/* grad_src[b,c,hw] = grad_tokens[(b*S_total+row_offset+hw), c]; if grad_vec != NULL also
 * grad_vec[c] += sum_{b,hw} grad_tokens[...] (grad_vec must be pre-initialised). */
int poet_tokens_to_nchw(const float* grad_tokens, float* grad_src, float* grad_vec, int B, int C, int HW,
                        int S_total, int row_offset, poet_stream_t stream);

/* ---- A6: padding-mask tokens and valid ratios ------------------------------------------------------- */
/* masks_host: L device pointers (on the HOST) to uint8/bool masks [B,H_l,W_l] (1 = padded).  pad [B,S] = the levels'
 * masks flattened and concatenated; valid_ratios [B,L,2] = (unpadded columns of row 0 / W, unpadded rows of
 * column 0 / H): deformable_transformer.py:111-118, 126-141 in one launch. */
int poet_mask_prep(const uint8_t* const* masks_host, const int32_t* shapes_host, uint8_t* pad, float* valid_ratios,
                   int B, int L, poet_stream_t stream);

/* ---- A5: encoder reference points -------------------------------------------------------- */
/* valid_ratios [B,L,2] (w,h).  out [B,S,L,2].  shapes_host [L*2] = (H_l, W_l). */
int poet_enc_reference_points(const float* valid_ratios, float* out, const int32_t* shapes_host,
                              int B, int L, poet_stream_t stream);

/* ---- A2/A3: multi-scale deformable attention -------------------------------------------- */
/* value [B,S,M,D] fp32.  shapes_host [L*2] = (H_l,W_l) on the HOST (no device sync).
 * mode 0 ("core", == upstream ms_deform_attn_forward): a = sampling locations [B,Lq,M,L,P,2] in [0,1],
 *         w = attention weights [B,Lq,M,L,P] (already soft-maxed), ref ignored.
 * mode 1 ("block"): a = raw sampling offsets, w = raw attention logits, ref [B,Lq,L,2]; the kernel
 *         applies softmax over L*P and loc = ref + off/(W_l,H_l) itself.
 * lda / ldw: row strides (floats) between consecutive (b,q) rows of a / w, so both may live
 *         inside one fused projection output row.  out [B,Lq,M*D]. */
int poet_msda_fwd(const float* value, const float* a, int64_t lda, const float* w, int64_t ldw,
                  const float* ref, float* out, const int32_t* shapes_host,
                  int B, int S, int Lq, int M, int D, int L, int P, int mode, poet_stream_t stream);
/* grad_value [B,S,M,D] is ACCUMULATED into (caller zero-fills); grad_a / grad_w are overwritten
 * (same strides as a / w).  mode 1 returns gradients w.r.t. the raw offsets / logits. */
int poet_msda_bwd(const float* value, const float* a, int64_t lda, const float* w, int64_t ldw,
                  const float* ref, const float* grad_out, float* grad_value, float* grad_a, float* grad_w,
                  const int32_t* shapes_host, int B, int S, int Lq, int M, int D, int L, int P, int mode,
                  poet_stream_t stream);

/* ---- dense contractions ----------------------------------------------------------------- */
/* C[M,N] = epilogue( alpha * op(A)[M,K] . op(B)[K,N] ).
 *   a_kcontig = 1: A stored [M,K] (lda = row stride), 0: A stored [K,M].
 *   b_kcontig = 1: B stored [N,K] (ldb = row stride) i.e. an nn.Linear weight, 0: B stored [K,N].
 * epilogue, in order: + bias[N] (nullable); relu if flags&POET_GEMM_RELU;
 *   *(gate>0) if gate != NULL (gate [M,N], ldc stride: relu backward); rows with row_mask[m]!=0 set to 0
 *   (MSDeformAttn value masked_fill); + C_old if flags&POET_GEMM_ACCUMULATE.
 * precision: POET_GEMM_FP32 (SIMT fp32 FFMA), POET_GEMM_BF16X3 / POET_GEMM_BF16 (tcgen05).
 * workspace: poet_gemm_workspace_bytes(); may be NULL when that is 0. */
/* POET_GEMM_B_STABLE: a promise that B (or its bf16 planes) is not written by the kernels that precede this call on the
 * stream -- it is a parameter, last written by poet_split_bf16* / poet_adamw_clip_multi (which never let a dependent
 * kernel start early) or by a non-library kernel.  The query-row GEMM kernel then requests its B tiles BEFORE it waits for
 * the preceding kernel (programmatic dependent launch), hiding one L2 round trip of the dependent chain. */
/* POET_GEMM_BACKGROUND: the launch shares the GPU with a chain of latency-bound kernels on another stream; the persistent
 * tensor-core kernel then occupies fewer SMs (POET_GEMM_BG_SMS, default 116 of 148) so that those kernels are not kept off the
 * machine for the duration of the GEMM.  Results are unaffected. */
enum { POET_GEMM_RELU = 1, POET_GEMM_ACCUMULATE = 2, POET_GEMM_B_STABLE = 4, POET_GEMM_BACKGROUND = 8 };
enum { POET_GEMM_FP32 = 0, POET_GEMM_BF16X3 = 1, POET_GEMM_BF16 = 2 };
size_t poet_gemm_workspace_bytes(int M, int N, int K, int a_kcontig, int b_kcontig, int precision);
int poet_gemm(const float* A, int64_t lda, int a_kcontig, const float* Bm, int64_t ldb, int b_kcontig,
              float* C, int64_t ldc, int M, int N, int K, float alpha, const float* bias, const float* gate,
              const uint8_t* row_mask, int flags, int precision, void* workspace, size_t workspace_bytes,
              poet_stream_t stream);
/* Weights are operands of every forward and dgrad GEMM of a step: split them ONCE into bf16 planes
 * (hi = bf16(x), lo = bf16(x - hi); lo may be NULL for POET_GEMM_BF16) and hand the planes to
 * poet_gemm_bsplit, whose B operand is then fetched by TMA straight into the swizzled smem stage.
 * B_hi/B_lo have the same logical layout and ldb (in elements) as the fp32 B; Bm (fp32) is still
 * required when the shape is not tensor-core eligible (poet_gemm_tc_eligible() == 0). */
int poet_split_bf16(const float* src, void* hi, void* lo, int64_t n, poet_stream_t stream);
/* The same split for every weight matrix of a model in ONE launch (a step re-derives ~85 planes).
 * table (device): n_tensors entries of 5 x 8 bytes {const float* src, void* hi, void* lo (may be 0),
 * int64 n/4 (float4 count), int64 first_chunk}; tensor t owns chunks [first_chunk_t, first_chunk_{t+1}) of
 * 1024 float4 each, entries sorted by first_chunk; total_chunks = sum_t ceil(n4_t / 1024). */
int poet_split_bf16_multi(const void* table, int n_tensors, int64_t total_chunks, poet_stream_t stream);
int poet_gemm_tc_eligible(int M, int N, int K, int64_t lda, int64_t ldb, int64_t ldc);
int poet_gemm_bsplit(const float* A, int64_t lda, int a_kcontig, const float* Bm, const void* B_hi, const void* B_lo,
                     int64_t ldb, int b_kcontig, float* C, int64_t ldc, int M, int N, int K, float alpha,
                     const float* bias, const float* gate, const uint8_t* row_mask, int flags, int precision,
                     poet_stream_t stream);
/* poet_gemm_bsplit on the tensor-core path only, with the ReLU of an FFN carried as a sign bitmask instead of
 * an fp32 re-read (words [M, N/32] row-major, bit c%32 of word [m, c/32]):
 *   relu_bits_out (with POET_GEMM_RELU): bit = (pre-activation > 0), written by the forward GEMM's epilogue;
 *   gate_bits: the dgrad GEMM keeps an element iff its bit is set (deformable_transformer.py:193-197 backward).
 *   a_colsum (weight-gradient shape only: a_kcontig = b_kcontig = 0, no pre-split B): a_colsum[m] += sum_k A[k,m],
 *   i.e. the bias gradient colsum(dY) accumulated while the dY tiles stream through the producers (caller
 *   zero-fills or passes an existing gradient).
 *   a_row_mask (nullable): uint8 per STORED row of A (a token in every GEMM of this path, whether A is [M,K] or
 *   [K,M]); rows with a non-zero entry are read as zeros: the backward of MSDeformAttn's value masked_fill applied
 *   while the gradient streams through the producers instead of by poet_mask_rows.
 *   drop_* (drop_p > 0): nn.Dropout on the epilogue's output, after the ReLU (deformable_transformer.py:194,268:
 *   dropout2 / dropout3 on relu(linear1(x))), pair scheme of poet_dropout; the keep mask is ANDed into relu_bits_out,
 *   so the backward is gate_bits + alpha = poet_dropout_scale(p, 1) on the dgrad and needs no second mask.
 * Returns POET_ERR_UNSUPPORTED when the shape is not tensor-core eligible (ask poet_gemm_relu_bits_supported()). */
int poet_gemm_relu_bits_supported(int M, int N, int K, int precision);
int poet_gemm_ex(const float* A, int64_t lda, int a_kcontig, const float* Bm, const void* B_hi, const void* B_lo,
                 int64_t ldb, int b_kcontig, float* C, int64_t ldc, int M, int N, int K, float alpha,
                 const float* bias, const uint8_t* row_mask, uint32_t* relu_bits_out, const uint32_t* gate_bits,
                 float* a_colsum, const uint8_t* a_row_mask, int flags, int precision, const void* drop_seed,
                 uint32_t drop_site, float drop_p, poet_stream_t stream);
/* out[N] (+)= sum_m X[m,n]  (bias gradients).  accumulate=0 overwrites. */
int poet_colsum(const float* X, int64_t ldx, float* out, int M, int N, int accumulate, poet_stream_t stream);
/* the same, skipping rows with row_mask[m] != 0 (row_mask nullable) */
int poet_colsum_masked(const float* X, int64_t ldx, const uint8_t* row_mask, float* out, int M, int N, int accumulate,
                       poet_stream_t stream);

/* ---- residual + LayerNorm ---------------------------------------------------------------- */
/* Train-mode dropout (reference nn.Dropout sites deformable_transformer.py:178-286, default p 0.1 main.py:94) is
 * counter-based everywhere in this library: element idx of site `drop_site` is kept iff hash(*drop_seed, drop_site,
 * idx) >= drop_p * 2^32 and scaled by 1/(1 - drop_p).  drop_seed: DEVICE pointer to one uint64, read when the kernel
 * runs (a replayed CUDA graph draws a new mask whenever the value changed); backward entry points regenerate the
 * forward's mask from the same triple.  drop_p = 0 (eval / parity path): no dropout, drop_seed may be NULL. */

/* z = x + dropout(r) (r nullable); y = LN(z)*gamma + beta; y2 = y + pos (if y2 != NULL);
 * xhat [R,C] and rstd [R] are saved for backward when non-NULL. C % 128 == 0, C <= 1024. */
int poet_add_layernorm_fwd(const float* x, const float* r, const float* gamma, const float* beta,
                           const float* pos, float* y, float* y2, float* xhat, float* rstd,
                           int R, int C, float eps, const void* drop_seed, uint32_t drop_site, float drop_p,
                           poet_stream_t stream);
/* dz = LN backward of (dy [+ dy2] [+ dy3] [+ dy4]) = gradient of x (dy2..dy4 nullable: the gradients of the other readers
 * of y -- y + pos, the value input, the next residual -- summed in registers instead of by accumulation kernels); dgamma/dbeta [C] are ACCUMULATED into (caller zero-fills).
 * With drop_p > 0 the gradient of the dropped branch r is written to dr [R,C] (= dz * mask / (1-p)); without
 * dropout it equals dz and dr may be NULL.  dr_colsum [C] (nullable) is ACCUMULATED with the column sums of the
 * residual branch's gradient: the bias gradient of the nn.Linear that produced r (reference linear2 / output_proj /
 * out_proj in front of norm2 / norm1, deformable_transformer.py:196-197,203-204,280-281,286-287). */
int poet_layernorm_bwd(const float* dy, const float* dy2, const float* dy3, const float* dy4, const float* xhat, const float* rstd,
                       const float* gamma, float* dz, float* dgamma, float* dbeta,
                       int R, int C, float* dr, float* dr_colsum, const void* drop_seed, uint32_t drop_site, float drop_p,
                       poet_stream_t stream);
/* x <- dropout(x) in place over n floats (n % 4 == 0) with the PAIR scheme the GEMM epilogue uses for the FFN hidden
 * activation (elements 2j, 2j+1 <- low / high half of hash(j), p quantised to 1/65536): fallback for a hidden
 * activation whose producing GEMM is not tensor-core eligible; same mask as poet_gemm_ex for the same triple. */
int poet_dropout(float* x, int64_t n, const void* drop_seed, uint32_t drop_site, float drop_p, poet_stream_t stream);
/* 1 / (1 - p) as the kernels apply it (pair_scheme != 0: with p quantised to 1/65536): the alpha of the dgrad GEMM. */
float poet_dropout_scale(float drop_p, int pair_scheme);
/* x[r,:] = 0 where mask[r] != 0 (in place; MSDeformAttn value masked_fill and its backward). */
int poet_mask_rows(float* x, const uint8_t* mask, int R, int C, poet_stream_t stream);
/* out = a + b (nullable b -> copy); elementwise over n floats, n % 4 == 0. */
int poet_add(const float* a, const float* b, float* out, int64_t n, poet_stream_t stream);

/* ---- block-level entry points: one call per reference sub-block (inference semantics: no dropout, nothing saved) ---- */
/* nn.Linear followed by the layer's residual + LayerNorm (reference deformable_transformer.py:201-204 output_proj +
 * norm1, :277-281 out_proj + norm2, :283-287):   y = LN(residual + x W^T + b) * gamma + beta     (gamma != NULL)
 * or just the Linear with its epilogue:            y = act(x W^T + b)                              (gamma == NULL; residual
 * must be NULL too), act = ReLU if flags & POET_GEMM_RELU.  x [R,K] (row stride ldx), W [N,K] fp32 (nn.Linear layout) with
 * optional bf16 planes W_hi / W_lo (poet_split_bf16; may be NULL), y [R,N].  With LayerNorm the pre-norm sum needs
 * poet_linear_epilogue_workspace_bytes() bytes of caller-owned workspace (N % 128 == 0, N <= 1024). */
size_t poet_linear_epilogue_workspace_bytes(int R, int N, int K, int with_layernorm);
int poet_linear_epilogue(const float* x, int64_t ldx, const float* W, const void* W_hi, const void* W_lo, const float* b,
                         const float* residual, const float* gamma, const float* beta, float* y, int R, int N, int K,
                         int flags, float eps, int precision, void* workspace, size_t workspace_bytes, poet_stream_t stream);
/* The FFN half of an encoder / decoder layer (reference deformable_transformer.py:193-197 + 205-206, 267-271 + 289-290):
 *     y = LN(x + linear2(relu(linear1(x)))) * gamma + beta,   x, y [R,C], W1 [F,C], W2 [C,F] (+ optional bf16 planes).
 * Three launches inside one call (GEMM + bias + ReLU, GEMM + bias, residual + LayerNorm); the hidden activation [R,F] and
 * the pre-norm sum [R,C] live in the caller's workspace (poet_ffn_fused_workspace_bytes()).  A single-kernel version
 * (hidden tiles kept in TMEM) was measured against this and rejected, see DESIGN.md section 4. */
size_t poet_ffn_fused_workspace_bytes(int R, int C, int F);
int poet_ffn_fused(const float* x, const float* W1, const void* W1_hi, const void* W1_lo, const float* b1,
                   const float* W2, const void* W2_hi, const void* W2_lo, const float* b2, const float* gamma,
                   const float* beta, float* y, int R, int C, int F, float eps, int precision, void* workspace,
                   size_t workspace_bytes, poet_stream_t stream);

/* ---- decoder self-attention core (Q <= 32) ---------------------------------------------- */
/* q,k,v: [B,Q,*] with row strides ldq/ldk/ldv (so they may be slices of one projection output);
 * head m uses channels [m*D,(m+1)*D).  probs [B,M,Q,Q] saved (before dropout).  out [B,Q,M*D].
 * drop_*: nn.MultiheadAttention's attention-probability dropout (deformable_transformer.py:253), see above. */
int poet_mha_smallq_fwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                        float* out, float* probs, int B, int Q, int M, int D, float scale, const void* drop_seed,
                        uint32_t drop_site, float drop_p, poet_stream_t stream);
int poet_mha_smallq_bwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                        const float* probs, const float* grad_out, float* gq, int64_t ldgq, float* gk, int64_t ldgk,
                        float* gv, int64_t ldgv, int B, int Q, int M, int D, float scale, const void* drop_seed,
                        uint32_t drop_site, float drop_p, poet_stream_t stream);

/* ---- heads: class-specific select + 6D -> SO(3) ------------------------------------------ */
/* rot_all [R, n_slots*6], trans_all [R, n_slots*3], classes [R] int64 (slot = max(cls,0); n_slots = 1 => slot 0).
 * out: trans [R,3], rot6d [R,6] (selected, saved for backward), rotmat [R,9] row-major 3x3 with columns (x,y,z). */
int poet_heads_select_rot6d_fwd(const float* rot_all, const float* trans_all, const int64_t* classes,
                                float* trans, float* rot6d, float* rotmat, int R, int n_slots,
                                poet_stream_t stream);
/* grad_rot_all / grad_trans_all are fully overwritten (zeros outside the selected slot). */
int poet_heads_select_rot6d_bwd(const float* rot6d, const int64_t* classes, const float* grad_trans,
                                const float* grad_rotmat, float* grad_rot_all, float* grad_trans_all,
                                int R, int n_slots, poet_stream_t stream);

/* ---- input_proj (next: SURVEY.md section 8f N1; reference pose_estimation_transformer.py:100-135, 313-335) ------ */
/* The 1x1 / 3x3-stride-2 convolutions are poet_gemm calls on token rows; these are the pieces around them.
 * col[(b*Ho + oy)*Wo + ox, c*9 + ky*3 + kx] = x[b, c, 2*oy-1+ky, 2*ox-1+kx] (0 outside), Ho = (H-1)/2+1. */
int poet_im2col_3x3s2(const float* x, float* col, int B, int C, int H, int W, poet_stream_t stream);
/* GroupNorm(G, C) of the conv output y [B*HW, C] (token rows of one level), written into rows
 * (b*S_total + row_offset + hw) of the transformer's token matrix.  stats [B,G,2] doubles (sum, sum of squares)
 * is written here and read by the backward; C % 32 == 0, C/G a power of two <= 32. */
int poet_groupnorm_tokens_fwd(const float* y, const float* gamma, const float* beta, float* tokens, double* stats,
                              int B, int HW, int C, int G, int S_total, int row_offset, float eps, poet_stream_t stream);
/* grad_y [B*HW, C] overwritten; dgamma / dbeta [C] ACCUMULATED into; workspace [B,G,2] doubles. */
int poet_groupnorm_tokens_bwd(const float* grad_tokens, const float* y, const double* stats, const float* gamma,
                              float* grad_y, float* dgamma, float* dbeta, double* workspace, int B, int HW, int C, int G,
                              int S_total, int row_offset, float eps, poet_stream_t stream);

/* ---- pose loss (next: SURVEY.md section 8f N2; reference pose_estimation_transformer.py:478-494, 519-537, 635-662) -- */
/* pred_t [L,B,Q,3], pred_R [L,B,Q,9] (row-major 3x3) for the L decoder layers; tgt_t [B,T,3], tgt_R [B,T,9] padded
 * targets; assign [B,Q] int32 = target index of each query or -1 (PoseMatcher result; 'gt' mode: j for j < n_i);
 * n_obj [1] int32 on the device = number of matched pairs (clamped to >= 1).
 * losses [L,2] = (sum ||t - t*||_2 / n_obj, sum acos(clamp((tr(R R*^T) - 1)/2, -1+1e-6, 1-1e-6)) / n_obj) per layer;
 * grad_t / grad_R = gradient of  sum_l (w_trans * losses[l,0] + w_rot * losses[l,1])  w.r.t. the predictions. */
int poet_pose_loss(const float* pred_t, const float* pred_R, const float* tgt_t, const float* tgt_R,
                   const int32_t* assign, const int32_t* n_obj, float* losses, float* grad_t, float* grad_R,
                   int L, int B, int Q, int T, float w_trans, float w_rot, poet_stream_t stream);

/* ---- optimizer step (next: SURVEY.md section 8f N3; reference engine.py:77-81, main.py:253-277) ------------ */
/* out[0] = sum_i x[i]^2 (double, on the device; overwritten).  n % 4 == 0. */
int poet_sumsq(const float* x, int64_t n, double* out, poet_stream_t stream);
/* Per-tensor squared gradient norms over the same pointer table as poet_adamw_clip_multi: tensor_sumsq[t] (device,
 * n_tensors floats, overwritten) = sum of squares of tensor t's slice of the gradient arena.  Their sum is the global
 * norm of torch.nn.utils.clip_grad_norm_ (engine.py:77-78); a tensor whose norm is exactly 0 received no gradient in
 * this step (torch leaves its .grad at None and torch.optim.AdamW does not touch it). */
int poet_grad_sumsq_multi(const void* table, int n_tensors, int64_t total_chunks, const float* grad,
                          float* tensor_sumsq, poet_stream_t stream);
/* clip_grad_norm_(max_norm) + AdamW.step() over every parameter tensor in one launch.
 * table (device): n_tensors entries of 7 x 8 bytes {float* param, int64 arena offset / 4, void* hi, void* lo,
 * int64 numel, int64 first_chunk, int32 lr group, int32 0}; tensor t owns ceil(ceil(numel/4)/1024) chunks of
 * 1024 float4, entries sorted by first_chunk (as in poet_split_bf16_multi); planes need numel % 8 == 0.  grad / m / v:
 * flat fp32 arenas sharing the offsets.  Gradient norm, one of: tensor_sumsq (poet_grad_sumsq_multi; tensors with a
 * zero norm are skipped like torch skips grad-None parameters, the total is written to sumsq if non-null, and
 * touched[t] (nullable, device floats) accumulates the norms so the host can tell which tensors own optimizer state),
 * or sumsq alone (poet_sumsq of the arena, read on the device; every tensor is updated).  Ignored when max_norm <= 0
 * and tensor_sumsq is null.  lr_host[n_groups]: learning rate per group (host).  step: 1-based step count (bias
 * correction).  hi / lo (nullable per tensor): bf16 planes of the UPDATED weights. */
int poet_adamw_clip_multi(const void* table, int n_tensors, int64_t total_chunks, const float* grad, float* m,
                          float* v, double* sumsq, const float* tensor_sumsq, float* touched, float max_norm,
                          const float* lr_host, int n_groups, float beta1, float beta2, float eps, float weight_decay,
                          int64_t step, poet_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* POET_B200_H_ */
