// Host-side interface of the LayerNorm kernel (layernorm.cu).
#pragma once
#include <cuda_runtime.h>

namespace aries {

// y[r, :] = (x[r, :] - mean) * rsqrt(var + eps) * gamma + beta;  x f16 [rows, d] (the residual stream), y bf16 [rows, d]; d % 128 == 0,
// d <= 2048.  (CT2 ops::LayerNorm, eps 1e-5)
cudaError_t layernorm_launch(const void* x_f16, const float* gamma, const float* beta, void* y_bf16, long long rows,
                             int d, float eps, cudaStream_t stream);

// Widths the register-resident kernel is instantiated for (d = 128 * {1..6, 8, 10, 12, 16}).
bool layernorm_supports(int d);

// Same kernel launched with programmatic stream serialisation (the decode step's dependent-launch chain).
cudaError_t layernorm_launch_pdl(const void* x_f16, const float* gamma, const float* beta, void* y_bf16, long long rows,
                                 int d, float eps, cudaStream_t stream, bool pdl);

}  // namespace aries
