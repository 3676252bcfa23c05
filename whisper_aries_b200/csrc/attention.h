// Host-side interface of the fused attention kernel (attention_sm100.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "gemm.h"

namespace aries {

struct AttnParams {
    int batch;
    int T;          // sequence length (1500)
    int d_model;    // n_heads * 64
    int n_heads;
    void* out;      // bf16 [batch * T, d_model]
};

cudaError_t attention_init_device();
// qk: bf16 [batch * T, 2 * d_model] (queries then keys, as the QKV GEMM writes them);
// vt: bf16 [batch, n_heads, 64, t_pad] (values, transposed by the QKV GEMM's epilogue).
struct AttnMaps {
    CUtensorMap q;     // [batch, T, 2d] view, box 64 columns x 128 rows (one query tile)
    CUtensorMap k;     // same tensor, box 64 columns x 64 rows (one key tile)
    CUtensorMap vt;    // [batch, heads*64, T] view of the transposed values, box 64 keys x 64 rows
};
cudaError_t attention_make_maps(const void* qk, const void* vt, int batch, int T, int d_model, int n_heads, int t_pad,
                                AttnMaps* maps);
cudaError_t attention_launch(const AttnMaps& maps, const AttnParams& p, cudaStream_t stream);
// Test-only: clock64 timeline written by the ARIES_ATTN_TRACE=1 variant ([16 CTAs][11 warps][160 stamps]).
cudaError_t attention_read_trace(unsigned long long* host, size_t count);

}  // namespace aries
