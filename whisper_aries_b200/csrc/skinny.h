// Host-side interface of the decode-step kernels (row f1): the weight-streaming "skinny" GEMM (skinny_gemm_sm100.cu),
// single-query attention over a key/value cache (decode_attention.cu) and the token embedding / sampling kernels
// (decode_misc.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace aries {

// ------------------------------------------------------------------------------------------------ skinny GEMM
//   out[b, n] = epilogue( sum_k X[b, k] * W[n, k] ),  b < B <= 128 sequences, W [N, K] bf16 K-major (a Linear weight).
// One decode step multiplies a handful of rows by every weight matrix of the decoder: the op is bound by streaming W
// from HBM once, so W is the 128-row M operand of tcgen05.mma, the sequences are the N operand (padded to NB = a
// multiple of 16), K is split over the CTAs of a thread-block cluster so that about one CTA per SM pulls bytes, and the
// partial sums are reduced through distributed shared memory in a fixed order (bit-reproducible).
enum SkinnyEpilogue {
    SK_BIAS_BF16 = 0,        // out bf16 [B, ldo] = acc + bias
    SK_BIAS_GELU_BF16 = 1,   // out bf16 = gelu_erf(acc + bias)
    SK_BIAS_RESID_F16 = 2,   // out f16 = acc + bias + out (in place on the f16 residual stream)
    SK_LOGITS_F32 = 3,       // out f32 [B, ldo] = acc (no bias; N need not be a multiple of 128)
    // LayerNorm FOLDED into the GEMM (any batch size; the algebra of the encoder's EPI_LN_*): the token-row operand is the
    // f16 residual stream itself, W holds f16(gamma * W), and with the row statistics (mean, rstd) from the partial sums
    // the producing SK_BIAS_RESID_F16 GEMM left in stats_in:  out = rstd * (acc - mean * c1[n]) + bias[n],
    // c1[n] = sum_k f16(gamma_k W[n, k]), bias[n] = sum_k beta_k W[n, k] + b[n].  Both operands are f16 (kind::f16 MMA).
    SK_LNF_BF16 = 4,         // out bf16
    SK_LNF_GELU_BF16 = 5,    // out bf16 = gelu_erf(...)
    SK_COUNT = 6,
};

struct SkinnyParams {
    int B;              // sequences (rows of X that are real)
    int NB;             // rows of the X tensor map's box = UMMA N: multiple of 16 in [16, 128], >= B
    int N, K;           // weight rows (output features), depth (K % 64 == 0)
    int splits;         // K splits (>= 1; every split owns at least one 64-deep block)
    const float* bias;  // [N] (unused by SK_LOGITS_F32)
    void* out;
    int ldo;            // elements between consecutive sequences in out
    int pdl;            // launched with programmatic stream serialisation (griddepcontrol in the kernel)
    // Fused LayerNorm (small batch): when ln_x != NULL the token-row operand is not loaded by TMA but computed in the
    // kernel as LayerNorm(ln_x[b, :]) * gamma + beta (f16 residual stream in, bf16 operand written straight into the
    // swizzled shared-memory tile) -- saves the separate LayerNorm launch of the latency-bound decode step.
    // Needs B <= 8, NB == 16, K <= 1280, K % 64 == 0 and at most 8 K-blocks per split.
    const void* ln_x;       // f16 [B, K]
    const float* ln_gamma;  // [K]
    const float* ln_beta;   // [K]
    // Folded LayerNorm (SK_LNF_*): per-row partial sums [B, stats_parts] of (x, x^2) over ln_dim features, and c1 [N].
    const float* c1;
    const float2* stats_in;
    int stats_parts, ln_dim;
    // SK_BIAS_RESID_F16 only, optional: partial (sum, sum of squares) of the f16 values this launch stores, one slot
    // per 128-feature tile: stats_out[b * gridDim.x + tile].  Needs splits > 1 (the cluster reduction path).
    float2* stats_out;
};

int skinny_pick_splits(int N, int K, int sm_count);
// same, for the fused-LayerNorm variant (every split must hold all of its K-blocks at once: <= 8); 0 if impossible
int skinny_pick_splits_ln(int N, int K, int sm_count);
cudaError_t skinny_init_device();
// tmap_w: [N, K] bf16, box 64 x 128; tmap_x: [>= NB rows, K] bf16, box 64 x NB.
cudaError_t skinny_launch(int epi, const CUtensorMap& tmap_w, const CUtensorMap& tmap_x, const SkinnyParams& p,
                          cudaStream_t stream);

// ------------------------------------------------------------------------------------------------ decode attention
// One query per (sequence, head), head_dim 64, keys / values bf16 rows of 64 contiguous elements:
//   out[b, h*64 .. +64] = softmax(q . K^T / 8) V
// Self-attention (n_keys_fixed == 0): first appends this step's key / value (new_k / new_v rows of the QKV buffer) to
// the cache at position *step, then attends to positions 0 .. *step.  Cross-attention: n_keys_fixed keys, optionally
// split over a cluster of `splits` <= 8 CTAs whose partial (max, sum, weighted values) rank 0 merges through
// distributed shared memory.
struct DecAttnParams {
    int batch, heads;
    const void* q;          // bf16, q of (b, h) at q + b * q_ld + h * 64
    int q_ld;
    void* k;                // bf16, key j of (b, h) at k + (b * kv_rows + j) * kv_ld + h * 64
    void* v;
    long long kv_rows;
    int kv_ld;
    const void* new_k;      // self-attention: this step's key / value of (b, h) at new_k + b * new_ld + h * 64
    const void* new_v;
    int new_ld;
    const int* step;        // device scalar (self-attention)
    int n_keys_fixed;       // cross-attention: 1500
    void* out;              // bf16, out + b * out_ld + h * 64
    int out_ld;
    int splits;
    int pdl;
    const int* done;        // optional [batch]: sequences that have emitted EOT skip the kernel (their cache is not read)
};
cudaError_t decode_attention_launch(const DecAttnParams& p, cudaStream_t stream);

// ------------------------------------------------------------------------------------------------ embedding / sampling
// x f16 [B, d] = emb[tokens[b, *step]] (bf16 table, tied with the output projection) + pos[*step] (f32 table)
// stats != NULL (folded LayerNorm): also stats[b, 0] = (sum, sum of squares) of the stored row, stats[b, 1 ..] = 0
cudaError_t decode_embed_launch(const int* tokens, int tokens_ld, const int* step, const void* emb_bf16, const float* pos,
                                void* x_f16, int batch, int d, int pdl, cudaStream_t stream, float2* stats = nullptr,
                                int stats_parts = 0);

// y bf16 [rows, d] = LayerNorm(x f16 [rows, d]) * gamma + beta (eps 1e-5), griddepcontrol-aware (rows = sequences).
cudaError_t decode_layernorm_launch(const void* x_f16, const float* gamma, const float* beta, void* y_bf16, int rows, int d,
                                    int pdl, cudaStream_t stream);

// Device-side mirror of the greedy step of ctranslate2 Whisper.generate (SuppressTokensBegin, SuppressTokens,
// ApplyTimestampRules, argmax, cumulative log-prob); see oracle/whisper_decoder.py apply_rules for the rules.
struct SampleParams {
    int batch, vocab;
    const float* logits;        // [batch, logits_ld]
    int logits_ld;
    int* tokens;                // [batch, tokens_ld]; positions < prompt_len hold the prompt
    int tokens_ld;
    int prompt_len;
    int max_length;             // total positions
    int* step;                  // device scalar: index of the token the decoder just consumed; incremented here
    const unsigned* suppress_bits;   // bitmask over the vocabulary (1 = never sampled)
    int suppress_blank, blank_id;
    int eot, no_speech, no_timestamps, timestamp_begin;
    int max_initial_timestamp_index;
    const int* sot_index;       // [batch] position of <|startoftranscript|> in the prompt
    const int* use_timestamps;  // [batch] 1 = apply the timestamp rules (prompt has no <|notimestamps|>)
    int* done;                  // [batch]
    int* last_timestamp;        // [batch] most recent timestamp token sampled (-1 none)
    float* score;               // [batch] sum of log-probs of the sampled tokens
    float* no_speech_prob;      // [batch]
    int* n_done;                // device scalar: finished sequences
    unsigned* ticket;           // device scalar, zero before the first launch
    const int* forced;          // tests: [batch, forced_ld] continuation to force (NULL in production)
    int forced_ld, n_forced;
    int* argmax_out;            // tests: [batch, tokens_ld] what the argmax was at each sampled position (or NULL)
    int pdl;
    unsigned* stack_bar;        // grid-barrier counter of decode_stack_kernel, zeroed here for the next step (or NULL)
};
cudaError_t decode_sample_launch(const SampleParams& p, cudaStream_t stream);

// ------------------------------------------------------------------------------------------------ persistent stack
// All decoder layers of one step in ONE cooperative kernel (decode_stack.cu), <= 8 sequences: one CTA per SM, weights
// prefetched through a shared-memory ring by a producer warp, grid-wide barriers between the phases of a layer.
struct StackLayerW {
    const void *wqkv, *wo, *wq2, *wo2, *w1, *w2;      // bf16 [N, K] row-major
    const float *ln1_g, *ln1_b, *bqkv, *bo, *ln2_g, *ln2_b, *bq2, *bo2, *ln3_g, *ln3_b, *b1, *b2;
};
constexpr int kPartStride = 72;                       // floats per attention partial: 64 sums, weight sum, max, pad
struct DecStackParams {
    const StackLayerW* layers;                        // device array [n_layers]
    int n_layers, d, f, heads, batch, max_batch;
    int C, A;                                         // n_text_ctx (rows per sequence of the caches), n_audio_ctx
    const int* tokens;                                // [batch, tokens_ld]
    int tokens_ld;
    const int* step;                                  // device scalar: position of the token being consumed
    const int* done;                                  // [batch] or NULL
    const void* emb;                                  // bf16 [vocab, d]
    const float* pos;                                 // f32 [C, d]
    void* x;                                          // f16 [batch, d] residual stream
    void *q, *q2, *h, *y, *ctx;                       // bf16 [batch, d | d | f | d | d]; y = final LayerNorm output
    void *kc, *vc;                                    // bf16 [layers][max_batch][C][d]
    const void* xkv;                                  // bf16 [layers][batch * A][2 d]
    float* part;                                      // decode_stack_part_floats(heads) floats, zero at allocation
    unsigned* cnt;                                    // arrival counters per (sequence, head): tail of `part` (set by the launch)
    unsigned* bar;                                    // zero at launch (decode_sample_kernel resets it)
    const float *lnf_g, *lnf_b;
    int splits_self, splits_cross, n_stages;          // filled by decode_stack_launch
    long long* trace;                                 // diagnostics (ARIES_STACK_TRACE): clock64 stamps of CTA trace_cta, or NULL
    int trace_cta;
};
bool decode_stack_supported(int d, int f, int heads, int batch, int sm_count);
size_t decode_stack_part_floats(int heads);
cudaError_t decode_stack_launch(const DecStackParams& p, int sm_count, cudaStream_t stream);

// griddepcontrol-aware launch helper shared by the decode kernels
cudaError_t launch_maybe_pdl(const void* func, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, void** args,
                             bool pdl);

}  // namespace aries
