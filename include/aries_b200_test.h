/* aries_b200_test.h — kernel-level hooks used by tests/ to check each CUDA kernel on its own against the oracle.
 * Not part of the drop-in surface (the reference has no counterpart); exported from the same library so that the
 * tests exercise exactly the code the product path runs. All pointers are device pointers; stream-ordered. */
#ifndef ARIES_B200_TEST_H_
#define ARIES_B200_TEST_H_

#include "aries_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* C[M,N] = A[M,K] (bf16, row-major) x B[N,K]^T (bf16, row-major) with one of the fused epilogues:
 *   epi 0: out bf16 = acc + bias              epi 1: out bf16 = gelu(acc + bias)
 *   epi 2: out f16  = acc + bias + resid (f16) epi 3: out f16  = gelu(acc + bias) + pos[row % pos_rows]
 *   epi 4: QKV split (rows are (b, t) with t_rows per item): columns < n_split -> out bf16 [M, n_split],
 *          columns >= n_split -> out2 bf16 [M / t_rows, (N - n_split) / 64, 64, t_pad]
 * N % 128 == 0, K % 64 == 0. */
ARIES_API int aries_test_gemm(aries_ctx* ctx, int epi, int M, int N, int K, const void* a, const void* b, const float* bias,
                    const void* resid, const float* pos, int pos_rows, void* out, void* out2, int n_split,
                    int t_rows, int t_pad, void* stream);

/* y bf16 [rows, d] = LayerNorm(x f32 [rows, d]) * gamma + beta, eps 1e-5. */
ARIES_API int aries_test_layernorm(aries_ctx* ctx, const void* x /* f16 */, const float* gamma, const float* beta, void* y,
                         int64_t rows, int d, void* stream);

/* out bf16 [batch*T, d] = softmax(Q K^T / 8) V per head; qk bf16 [batch*T, 2d]; vt bf16 [batch, heads, 64, t_pad]. */
ARIES_API int aries_test_attention(aries_ctx* ctx, const void* qk, const void* vt, int batch, int T, int n_heads, int t_pad,
                         void* out, void* stream);

/* Timeline of the ARIES_ATTN_TRACE=1 attention variant: [16 CTAs][11 warps][160 clock64 stamps] (0 = unused). */
ARIES_API int aries_test_attention_trace(aries_ctx* ctx, unsigned long long* host, size_t count);

#ifdef __cplusplus
}
#endif
#endif /* ARIES_B200_TEST_H_ */
