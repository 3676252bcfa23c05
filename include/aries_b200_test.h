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

/* The LayerNorm-folded GEMM epilogues and the statistics hand-over (csrc/gemm.h):
 *   epi 2 (bias + residual -> f16) with stats_out != NULL also writes (sum, sum of squares) per row and per slice of
 *     aries_test_gemm_stats_parts(N) column slices: stats_out f32 [M, parts, 2];
 *   epi 5: out bf16 = gelu(rstd (a b^T - mean c1) + bias), epi 6: the same without GELU, split / transposed like epi 4;
 *     a = f16 [M, K] (the residual stream), b = f16 [N, K] (gamma-scaled weights), mean / rstd from stats_in
 *     f32 [M, stats_parts, 2] over ln_dim columns, eps 1e-5. */
ARIES_API int aries_test_gemm_ln(aries_ctx* ctx, int epi, int M, int N, int K, const void* a, const void* b, const float* bias,
                       const float* c1, const void* stats_in, int stats_parts, int ln_dim, const void* resid, void* out,
                       void* out2, int n_split, int t_rows, int t_pad, void* stats_out, void* stream);
ARIES_API int aries_test_gemm_stats_parts(int N);

/* y bf16 [rows, d] = LayerNorm(x f32 [rows, d]) * gamma + beta, eps 1e-5. */
ARIES_API int aries_test_layernorm(aries_ctx* ctx, const void* x /* f16 */, const float* gamma, const float* beta, void* y,
                         int64_t rows, int d, void* stream);

/* out bf16 [batch*T, d] = softmax(Q K^T / 8) V per head; qk bf16 [batch*T, 2d]; vt bf16 [batch, heads, 64, t_pad]. */
ARIES_API int aries_test_attention(aries_ctx* ctx, const void* qk, const void* vt, int batch, int T, int n_heads, int t_pad,
                         void* out, void* stream);

/* Timeline of the ARIES_ATTN_TRACE=1 attention variant: [16 CTAs][11 warps][160 clock64 stamps] (0 = unused). */
ARIES_API int aries_test_attention_trace(aries_ctx* ctx, unsigned long long* host, size_t count);

/* ---- decode-step kernels (row f1) ---- */
/* out[b, n] = epilogue(sum_k x[b, k] w[n, k]); x bf16 [round_up(B, 16), K] (pad rows zero), w bf16 [N, K];
 *   epi 0: out bf16 = acc + bias     epi 1: out bf16 = gelu(acc + bias)
 *   epi 2: out f16 += acc + bias     epi 3: out f32 = acc (N need not be a multiple of 128)
 * K % 64 == 0, B <= 128; splits = 0 picks the production split, else 1..8 (cluster size). */
ARIES_API int aries_test_skinny_gemm(aries_ctx* ctx, int epi, int B, int N, int K, const void* x, const void* w,
                           const float* bias, void* out, int ldo, int splits, void* stream);

/* Same GEMM with the LayerNorm of the decode step fused into the operand load (B <= 8, K <= 1280, epi 0 | 1 | 3):
 * out[b, n] = epilogue(sum_k LN(x_f16[b, :])[k] * gamma[k] + beta[k]) w[n, k]); x_f16 f16 [B, K]. */
ARIES_API int aries_test_skinny_gemm_ln(aries_ctx* ctx, int epi, int B, int N, int K, const void* x_f16, const float* gamma,
                              const float* beta, const void* w, const float* bias, void* out, int ldo, int splits,
                              void* stream);

/* The decode step's folded LayerNorm, both sides in one call (csrc/skinny.h SK_LNF_*):
 *   (1) resid_x_f16 [NB, K] f16 += ctx_in [NB, K] bf16 . w_o [K, K]^T + bias_o   (SK_BIAS_RESID_F16, in place) and the
 *       partial (sum, sum of squares) of the stored rows -> stats [B, ceil(K / 128)] float2;
 *   (2) out_bf16 [B, N] = [gelu](rstd (x . w_folded_f16^T - mean c1) + c2) with w_folded_f16 [N, K] = f16(gamma * W),
 *       c1[n] = sum_k w_folded[n, k], c2[n] = sum_k beta_k W[n, k] + b[n].
 * NB = B rounded up to 16 (rows >= B must be zero); K % 64 == 0, K <= 2048. */
ARIES_API int aries_test_skinny_gemm_folded(aries_ctx* ctx, int B, int N, int K, const void* resid_x_f16, const void* ctx_in,
                                  const void* w_o, const float* bias_o, const void* w_folded_f16, const float* c1,
                                  const float* c2, int gelu, void* out_bf16, float* stats, void* stream);

/* Single-query attention over a bf16 key/value cache (see csrc/skinny.h DecAttnParams). n_keys_fixed == 0: self-attention
 * (appends new_k / new_v at position *step_dev, attends to 0 .. *step_dev); otherwise cross-attention over n_keys_fixed
 * keys with `splits` CTAs per (sequence, head). */
ARIES_API int aries_test_decode_attention(aries_ctx* ctx, const void* q, int q_ld, void* k, void* v, int64_t kv_rows, int kv_ld,
                                const void* new_k, const void* new_v, int new_ld, const int* step_dev, int n_keys_fixed,
                                int batch, int heads, void* out, int out_ld, int splits, void* stream);

/* aries_decoder_generate with teacher forcing: the first n_forced sampled positions take forced[b, i] (host int32
 * [batch, n_forced]) instead of the argmax; argmax_out (host int32 [batch, max_length], -1 where nothing was sampled)
 * receives what the argmax was; logits_out (host f32 [max_length - 1, batch, vocab], may be NULL) the raw logits of
 * every step (disables the CUDA graph). */
ARIES_API int aries_test_decoder_generate(aries_decoder* dec, const void* enc_out_dev, int batch, const int32_t* prompts,
                                int prompt_len, const aries_generate_opts* opts, const int32_t* forced, int n_forced,
                                int32_t* tokens_out, int32_t* argmax_out, float* logits_out, float* scores,
                                float* no_speech_prob, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ARIES_B200_TEST_H_ */
