// Library-level entry points of the C ABI (include/poet_b200.h) and the poet_gemm dispatcher.
#include "common.cuh"

int poet_gemm_simt(const float* A, int64_t lda, int a_kcontig, const float* Bm, int64_t ldb, int b_kcontig, float* C,
                   int64_t ldc, int M, int N, int K, float alpha, const float* bias, const float* gate,
                   const uint8_t* row_mask, int flags, cudaStream_t s);
bool poet_gemm_small_supported(int M, int N, int K, int a_kcontig, int b_kcontig, int64_t lda, int64_t ldb, int64_t ldc,
                               const void* A, const void* B, const void* Bhi, const void* Blo, bool relu_or_gate);
int poet_gemm_small(const float* A, int64_t lda, int a_kcontig, const float* Bm, const void* b_hi, const void* b_lo,
                    int64_t ldb, int b_kcontig, float* C, int64_t ldc, int M, int N, int K, float alpha, const float* bias,
                    const float* gate, float* a_colsum, int flags, cudaStream_t s);
#ifdef POET_HAVE_TC_GEMM
size_t poet_gemm_tc_workspace_bytes(int M, int N, int K, int a_kcontig, int b_kcontig, int precision);
int poet_gemm_tc(const float* A, int64_t lda, int a_kcontig, const float* Bm, const void* b_hi, const void* b_lo,
                 int64_t ldb, int b_kcontig, float* C, int64_t ldc, int M, int N, int K, float alpha, const float* bias,
                 const float* gate, const uint8_t* row_mask, uint32_t* relu_bits, const uint32_t* gate_bits, float* a_colsum,
                 const uint8_t* a_row_mask, int flags, int precision, cudaStream_t s, const PoetDropout* drop = nullptr);
int poet_gemm_tc_bits_supported();
bool poet_gemm_tc_supported(int M, int N, int K, int a_kcontig, int b_kcontig, int64_t lda, int64_t ldb, int64_t ldc);
int poet_split_bf16_impl(const float* src, void* hi, void* lo, int64_t n, cudaStream_t s);
int poet_split_bf16_multi_impl(const void* table_dev, int n_tensors, int64_t total_chunks, cudaStream_t s);
#endif

extern "C" int poet_version(void) { return 1; }
extern "C" int poet_sm(void) { return 100; }

extern "C" int poet_check_device(int device) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return (int)e;
  return (prop.major == 10 && prop.minor == 0) ? POET_OK : POET_ERR_WRONG_DEVICE;
}

extern "C" const char* poet_error_string(int code) {
  switch (code) {
    case POET_OK: return "ok";
    case POET_ERR_BAD_SHAPE: return "poet_b200: bad shape";
    case POET_ERR_BAD_ALIGNMENT: return "poet_b200: pointer or stride not 16-byte aligned";
    case POET_ERR_UNSUPPORTED: return "poet_b200: unsupported configuration";
    case POET_ERR_NULL_POINTER: return "poet_b200: required pointer is NULL";
    case POET_ERR_WORKSPACE: return "poet_b200: workspace missing or too small";
    case POET_ERR_WRONG_DEVICE: return "poet_b200: device is not compute capability 10.0 (B200)";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "poet_b200: unknown error";
}

extern "C" size_t poet_gemm_workspace_bytes(int M, int N, int K, int a_kcontig, int b_kcontig, int precision) {
#ifdef POET_HAVE_TC_GEMM
  if (precision != POET_GEMM_FP32) return poet_gemm_tc_workspace_bytes(M, N, K, a_kcontig, b_kcontig, precision);
#endif
  (void)M; (void)N; (void)K; (void)a_kcontig; (void)b_kcontig; (void)precision;
  return 0;
}

extern "C" int poet_gemm(const float* A, int64_t lda, int a_kcontig, const float* Bm, int64_t ldb, int b_kcontig,
                         float* C, int64_t ldc, int M, int N, int K, float alpha, const float* bias, const float* gate,
                         const uint8_t* row_mask, int flags, int precision, void* workspace, size_t workspace_bytes,
                         poet_stream_t stream) {
  POET_REQUIRE(A && Bm && C, POET_ERR_NULL_POINTER);
  POET_REQUIRE(M > 0 && N > 0 && K > 0, POET_ERR_BAD_SHAPE);
  POET_REQUIRE(lda >= (a_kcontig ? K : M) && ldb >= (b_kcontig ? K : N) && ldc >= N, POET_ERR_BAD_SHAPE);
  POET_REQUIRE(precision >= POET_GEMM_FP32 && precision <= POET_GEMM_BF16, POET_ERR_UNSUPPORTED);
  cudaStream_t s = (cudaStream_t)stream;
  // query-row problems (decoder chain, pose heads): the latency-optimised exact-fp32 kernel, whatever the precision mode
  if (row_mask == nullptr && poet_gemm_small_supported(M, N, K, a_kcontig, b_kcontig, lda, ldb, ldc, A, Bm, nullptr, nullptr,
                                                       gate != nullptr || (flags & POET_GEMM_RELU)))
    return poet_gemm_small(A, lda, a_kcontig, Bm, nullptr, nullptr, ldb, b_kcontig, C, ldc, M, N, K, alpha, bias, gate, nullptr,
                           flags, s);
#ifdef POET_HAVE_TC_GEMM
  if (precision != POET_GEMM_FP32 && poet_gemm_tc_supported(M, N, K, a_kcontig, b_kcontig, lda, ldb, ldc))
    return poet_gemm_tc(A, lda, a_kcontig, Bm, nullptr, nullptr, ldb, b_kcontig, C, ldc, M, N, K, alpha, bias, gate,
                        row_mask, nullptr, nullptr, nullptr, nullptr, flags, precision, s);
#endif
  (void)workspace; (void)workspace_bytes;
  // small / ragged shapes (decoder rows, head outputs) always take the exact fp32 SIMT path
  return poet_gemm_simt(A, lda, a_kcontig, Bm, ldb, b_kcontig, C, ldc, M, N, K, alpha, bias, gate, row_mask, flags, s);
}

extern "C" int poet_gemm_tc_eligible(int M, int N, int K, int64_t lda, int64_t ldb, int64_t ldc) {
#ifdef POET_HAVE_TC_GEMM
  return poet_gemm_tc_supported(M, N, K, 1, 1, lda, ldb, ldc) ? 1 : 0;
#else
  (void)M; (void)N; (void)K; (void)lda; (void)ldb; (void)ldc;
  return 0;
#endif
}

extern "C" int poet_split_bf16(const float* src, void* hi, void* lo, int64_t n, poet_stream_t stream) {
#ifdef POET_HAVE_TC_GEMM
  return poet_split_bf16_impl(src, hi, lo, n, (cudaStream_t)stream);
#else
  (void)src; (void)hi; (void)lo; (void)n; (void)stream;
  return POET_ERR_UNSUPPORTED;
#endif
}

extern "C" int poet_split_bf16_multi(const void* table, int n_tensors, int64_t total_chunks, poet_stream_t stream) {
#ifdef POET_HAVE_TC_GEMM
  return poet_split_bf16_multi_impl(table, n_tensors, total_chunks, (cudaStream_t)stream);
#else
  (void)table; (void)n_tensors; (void)total_chunks; (void)stream;
  return POET_ERR_UNSUPPORTED;
#endif
}

extern "C" int poet_gemm_relu_bits_supported(int M, int N, int K, int precision) {
#ifdef POET_HAVE_TC_GEMM
  return (precision != POET_GEMM_FP32 && poet_gemm_tc_supported(M, N, K, 1, 1, K, K, N) && poet_gemm_tc_bits_supported()) ? 1 : 0;
#else
  (void)M; (void)N; (void)K; (void)precision;
  return 0;
#endif
}

extern "C" int poet_gemm_ex(const float* A, int64_t lda, int a_kcontig, const float* Bm, const void* B_hi,
                            const void* B_lo, int64_t ldb, int b_kcontig, float* C, int64_t ldc, int M, int N, int K,
                            float alpha, const float* bias, const uint8_t* row_mask, uint32_t* relu_bits_out,
                            const uint32_t* gate_bits, float* a_colsum, const uint8_t* a_row_mask, int flags, int precision,
                            const void* drop_seed, uint32_t drop_site, float drop_p, poet_stream_t stream) {
  POET_REQUIRE(A && C && (Bm || B_hi), POET_ERR_NULL_POINTER);
  POET_REQUIRE(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || drop_seed != nullptr), POET_ERR_BAD_SHAPE);
  POET_REQUIRE(M > 0 && N > 0 && K > 0, POET_ERR_BAD_SHAPE);
  POET_REQUIRE(lda >= (a_kcontig ? K : M) && ldb >= (b_kcontig ? K : N) && ldc >= N, POET_ERR_BAD_SHAPE);
  POET_REQUIRE(!relu_bits_out || (flags & POET_GEMM_RELU), POET_ERR_UNSUPPORTED);
  if (!relu_bits_out && !gate_bits && !row_mask && !a_row_mask && drop_p == 0.f &&
      poet_gemm_small_supported(M, N, K, a_kcontig, b_kcontig, lda, ldb, ldc, A, Bm, B_hi, B_lo, flags & POET_GEMM_RELU))
    return poet_gemm_small(A, lda, a_kcontig, Bm, B_hi, B_lo, ldb, b_kcontig, C, ldc, M, N, K, alpha, bias, nullptr, a_colsum,
                           flags, (cudaStream_t)stream);
#ifdef POET_HAVE_TC_GEMM
  POET_REQUIRE(precision == POET_GEMM_BF16X3 || precision == POET_GEMM_BF16, POET_ERR_UNSUPPORTED);
  POET_REQUIRE(poet_gemm_tc_supported(M, N, K, a_kcontig, b_kcontig, lda, ldb, ldc), POET_ERR_UNSUPPORTED);
  POET_REQUIRE(!a_colsum || (!a_kcontig && !b_kcontig && !B_hi), POET_ERR_UNSUPPORTED);
  const PoetDropout drop = poet_make_dropout(drop_seed, drop_site, drop_p);
  return poet_gemm_tc(A, lda, a_kcontig, Bm, B_hi, B_lo, ldb, b_kcontig, C, ldc, M, N, K, alpha, bias, nullptr, row_mask,
                      relu_bits_out, gate_bits, a_colsum, a_row_mask, flags, precision, (cudaStream_t)stream, &drop);
#else
  (void)alpha; (void)bias; (void)row_mask; (void)gate_bits; (void)a_colsum; (void)a_row_mask; (void)precision; (void)stream;
  (void)drop_seed; (void)drop_site;
  return POET_ERR_UNSUPPORTED;
#endif
}

extern "C" int poet_gemm_bsplit(const float* A, int64_t lda, int a_kcontig, const float* Bm, const void* B_hi,
                                const void* B_lo, int64_t ldb, int b_kcontig, float* C, int64_t ldc, int M, int N, int K,
                                float alpha, const float* bias, const float* gate, const uint8_t* row_mask, int flags,
                                int precision, poet_stream_t stream) {
  POET_REQUIRE(A && C && (Bm || B_hi), POET_ERR_NULL_POINTER);
  POET_REQUIRE(M > 0 && N > 0 && K > 0, POET_ERR_BAD_SHAPE);
  POET_REQUIRE(lda >= (a_kcontig ? K : M) && ldb >= (b_kcontig ? K : N) && ldc >= N, POET_ERR_BAD_SHAPE);
  POET_REQUIRE(precision >= POET_GEMM_FP32 && precision <= POET_GEMM_BF16, POET_ERR_UNSUPPORTED);
  cudaStream_t s = (cudaStream_t)stream;
  if (row_mask == nullptr && poet_gemm_small_supported(M, N, K, a_kcontig, b_kcontig, lda, ldb, ldc, A, Bm, B_hi, B_lo,
                                                       gate != nullptr || (flags & POET_GEMM_RELU)))
    return poet_gemm_small(A, lda, a_kcontig, Bm, B_hi, B_lo, ldb, b_kcontig, C, ldc, M, N, K, alpha, bias, gate, nullptr,
                           flags, s);
#ifdef POET_HAVE_TC_GEMM
  if (precision != POET_GEMM_FP32 && poet_gemm_tc_supported(M, N, K, a_kcontig, b_kcontig, lda, ldb, ldc))
    return poet_gemm_tc(A, lda, a_kcontig, Bm, B_hi, B_lo, ldb, b_kcontig, C, ldc, M, N, K, alpha, bias, gate, row_mask,
                        nullptr, nullptr, nullptr, nullptr, flags, precision, s);
#endif
  POET_REQUIRE(Bm != nullptr, POET_ERR_NULL_POINTER);
  return poet_gemm_simt(A, lda, a_kcontig, Bm, ldb, b_kcontig, C, ldc, M, N, K, alpha, bias, gate, row_mask, flags, s);
}

// ------------------------------------------------------------------------------------------
// block-level entry points (inference semantics): compositions of the kernels above behind one call
// ------------------------------------------------------------------------------------------
static inline size_t poet_align256(size_t n) { return (n + 255) & ~(size_t)255; }

extern "C" size_t poet_linear_epilogue_workspace_bytes(int R, int N, int K, int with_layernorm) {
  (void)K;
  return with_layernorm ? poet_align256((size_t)R * N * sizeof(float)) : 0;
}

extern "C" int poet_linear_epilogue(const float* x, int64_t ldx, const float* W, const void* W_hi, const void* W_lo,
                                    const float* b, const float* residual, const float* gamma, const float* beta, float* y,
                                    int R, int N, int K, int flags, float eps, int precision, void* workspace,
                                    size_t workspace_bytes, poet_stream_t stream) {
  POET_REQUIRE(x && W && y, POET_ERR_NULL_POINTER);
  POET_REQUIRE(R > 0 && N > 0 && K > 0 && ldx >= K, POET_ERR_BAD_SHAPE);
  POET_REQUIRE((flags & ~POET_GEMM_RELU) == 0, POET_ERR_UNSUPPORTED);
  const bool ln = gamma != nullptr;
  if (!ln) {
    POET_REQUIRE(residual == nullptr && beta == nullptr, POET_ERR_UNSUPPORTED);
    return poet_gemm_bsplit(x, ldx, 1, W, W_hi, W_lo, K, 1, y, N, R, N, K, 1.f, b, nullptr, nullptr, flags, precision, stream);
  }
  POET_REQUIRE(beta != nullptr, POET_ERR_NULL_POINTER);
  POET_REQUIRE(!(flags & POET_GEMM_RELU), POET_ERR_UNSUPPORTED);      // no reference layer normalises an activated Linear
  POET_REQUIRE(workspace && workspace_bytes >= poet_linear_epilogue_workspace_bytes(R, N, K, 1), POET_ERR_WORKSPACE);
  float* lin = reinterpret_cast<float*>(workspace);
  int rc = poet_gemm_bsplit(x, ldx, 1, W, W_hi, W_lo, K, 1, lin, N, R, N, K, 1.f, b, nullptr, nullptr, 0, precision, stream);
  if (rc) return rc;
  // z = residual + lin (residual == NULL: LN(lin)); nothing saved for a backward pass
  return poet_add_layernorm_fwd(residual ? residual : lin, residual ? lin : nullptr, gamma, beta, nullptr, y, nullptr, nullptr,
                                nullptr, R, N, eps, nullptr, 0, 0.f, stream);
}

extern "C" size_t poet_ffn_fused_workspace_bytes(int R, int C, int F) {
  return poet_align256((size_t)R * F * sizeof(float)) + poet_align256((size_t)R * C * sizeof(float));
}

extern "C" int poet_ffn_fused(const float* x, const float* W1, const void* W1_hi, const void* W1_lo, const float* b1,
                              const float* W2, const void* W2_hi, const void* W2_lo, const float* b2, const float* gamma,
                              const float* beta, float* y, int R, int C, int F, float eps, int precision, void* workspace,
                              size_t workspace_bytes, poet_stream_t stream) {
  POET_REQUIRE(x && W1 && W2 && gamma && beta && y, POET_ERR_NULL_POINTER);
  POET_REQUIRE(R > 0 && C > 0 && F > 0, POET_ERR_BAD_SHAPE);
  POET_REQUIRE(workspace && workspace_bytes >= poet_ffn_fused_workspace_bytes(R, C, F), POET_ERR_WORKSPACE);
  float* hidden = reinterpret_cast<float*>(workspace);
  float* lin = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + poet_align256((size_t)R * F * sizeof(float)));
  int rc = poet_gemm_bsplit(x, C, 1, W1, W1_hi, W1_lo, C, 1, hidden, F, R, F, C, 1.f, b1, nullptr, nullptr, POET_GEMM_RELU,
                            precision, stream);
  if (rc) return rc;
  rc = poet_gemm_bsplit(hidden, F, 1, W2, W2_hi, W2_lo, F, 1, lin, C, R, C, F, 1.f, b2, nullptr, nullptr, 0, precision, stream);
  if (rc) return rc;
  return poet_add_layernorm_fwd(x, lin, gamma, beta, nullptr, y, nullptr, nullptr, nullptr, R, C, eps, nullptr, 0, 0.f, stream);
}
