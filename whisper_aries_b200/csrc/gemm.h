// Host-side interface of the tcgen05 GEMM (gemm_sm100.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace aries {

enum GemmEpilogue {
    EPI_BIAS_BF16 = 0,          // out bf16 = acc + bias                       (QKV projection)
    EPI_BIAS_GELU_BF16 = 1,     // out bf16 = gelu_erf(acc + bias)             (MLP fc1, conv1)
    EPI_BIAS_RESID_F16 = 2,     // out f16  = acc + bias + resid (f16)         (attention out-proj, MLP fc2)
    EPI_BIAS_GELU_POS_F16 = 3,  // out f16  = gelu_erf(acc + bias) + pos[t]    (conv2 + positional table)
    EPI_QKV_SPLIT_BF16 = 4,     // acc + bias: columns < n_split -> out (row-major, ld = ldo, queries | keys);
                                // columns >= n_split -> out2[b][head][c][t] (values, transposed for attention)
    // LayerNorm folded into the GEMM (north_star (2)): the A operand is the f16 residual stream x itself, the weights
    // carry gamma (W' = gamma (.) W, stored f16), and with c1[n] = sum_k W'[n,k], c2[n] = sum_k beta[k] W[n,k] + bias[n]
    //   LN(x) W^T + bias = rstd * (x W'^T - mean * c1) + c2,
    // mean / rstd per row from the (sum, sum of squares) partials the producing epilogue left in `stats_in`.
    EPI_LN_GELU_BF16 = 5,       // out bf16 = gelu_erf(rstd (acc - mean c1) + c2)                        (MLP fc1)
    EPI_LN_QKV_SPLIT_BF16 = 6,  // rstd (acc - mean c1) + c2, split / transposed like EPI_QKV_SPLIT_BF16 (QKV projection)
    EPI_COUNT = 7,
};

struct GemmParams {
    int M, N, K;        // GEMM rows (incl. rows the remap drops), columns, depth (K % 64 == 0)
    int a_cols;         // inner extent of A's tensor map: K index kk reads (row + kk / a_cols, col kk % a_cols)
    int p_in;           // row r -> b = r / p_in, t = r % p_in
    int t_valid;        // row is written only when t < t_valid
    int p_out;          // output row = b * p_out + t + row_off
    int row_off;
    int ldo;            // leading dimension of out / resid (elements)
    const float* bias;  // [N]
    const void* resid;  // f16, indexed like out (EPI_BIAS_RESID_F16; may alias out): the residual stream
    const float* pos;   // f32 [t_valid, N] (EPI_BIAS_GELU_POS_F16)
    void* out;
    void* out2;         // EPI_QKV_SPLIT_BF16: bf16 [batch, (N - n_split) / 64, 64, t_pad]
    int n_split;        // EPI_QKV_SPLIT_BF16: first column of the transposed part (2 * d_model)
    int t_pad;          // EPI_QKV_SPLIT_BF16: row pitch of out2 (elements)
    // LayerNorm statistics hand-over between GEMMs.  A producing epilogue (EPI_BIAS_RESID_F16 / EPI_BIAS_GELU_POS_F16
    // with stats_out != nullptr) writes, per OUTPUT row and per `BN / 2`-column slice it owns, (sum, sum of squares) of
    // the f32 values it stores: stats_out[row * (N / (BN / 2)) + slice].  A consuming epilogue (EPI_LN_*) sums the
    // `stats_parts` partials of GEMM row r (its rows are the producer's output rows) into mean / rstd over ln_dim columns.
    float2* stats_out;
    const float2* stats_in;
    int stats_parts;
    int ln_dim;         // normalised width (d_model)
    float ln_eps;
    const float* c1;    // [N] (EPI_LN_*); c2 is passed in `bias`
    // test-only timeline (tests/gemm_trace.py): clock64 stamps of CTA 0's MMA warp and first epilogue warp,
    // [tile][8]: mma {before accumulator wait, after, cycles spent waiting for operands, main loop end} |
    //            epilogue {before accumulator-full wait, after, end of tile, unused}
    unsigned long long* trace;
    int trace_tiles;
};

int gemm_block_n(int N);
int gemm_b_box_rows();      // rows of the B (weight) tensor map's box: 128 for every kernel variant
cudaError_t gemm_init_device();
cudaError_t gemm_launch(int epi, const CUtensorMap& tmap_a, const CUtensorMap& tmap_b, const GemmParams& p,
                        int sm_count, cudaStream_t stream);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency).
// bf16 tensor, dims/strides innermost first; strides[0] is implied (2 bytes); 128-byte swizzle.
// number of (sum, sum of squares) slices a stats-producing GEMM with N output columns writes per row
inline int gemm_stats_parts(int N) { return N / (gemm_block_n(N) / 2); }

cudaError_t make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const unsigned long long* dims,
                           const unsigned long long* strides_bytes, const unsigned* box);

}  // namespace aries
