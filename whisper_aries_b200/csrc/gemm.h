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
    EPI_COUNT = 5,
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
};

int gemm_block_n(int N);
int gemm_b_box_rows();      // rows of the B (weight) tensor map's box: 128 for every kernel variant
cudaError_t gemm_init_device();
cudaError_t gemm_launch(int epi, const CUtensorMap& tmap_a, const CUtensorMap& tmap_b, const GemmParams& p,
                        int sm_count, cudaStream_t stream);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency).
// bf16 tensor, dims/strides innermost first; strides[0] is implied (2 bytes); 128-byte swizzle.
cudaError_t make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const unsigned long long* dims,
                           const unsigned long long* strides_bytes, const unsigned* box);

}  // namespace aries
