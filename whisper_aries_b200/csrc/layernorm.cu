// K4: LayerNorm over the f16 residual stream, bf16 out (the A operand of the next GEMM, or the encoder output).
// One warp per row; the row stays in registers (as f32) between the mean, variance and normalise passes, so HBM sees
// one f16 read and one bf16 write per element; statistics are f32.  (CT2 ops::LayerNorm; eps 1e-5.)
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "layernorm.h"

namespace aries {

namespace {

constexpr int kWarpsPerBlock = 8;

template <int NV>   // NV float4 per lane: d = 128 * NV
__global__ void __launch_bounds__(kWarpsPerBlock * 32) layernorm_kernel(const __half* __restrict__ x,
                                                                        const float* __restrict__ gamma,
                                                                        const float* __restrict__ beta,
                                                                        __nv_bfloat16* __restrict__ y, long long rows,
                                                                        float eps) {
    // no-ops unless launched with programmatic stream serialisation (the decode step, skinny.h)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (row >= rows) return;
    constexpr int d = NV * 128;
    const uint2* xr = reinterpret_cast<const uint2*>(x + row * d);
    float4 v[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const uint2 h4 = __ldcs(xr + lane + 32 * j);
        const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&h4.x));
        const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&h4.y));
        v[j] = make_float4(lo.x, lo.y, hi.x, hi.y);
    }
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < NV; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / d);
    float q = 0.0f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, e = v[j].w - mean;
        q += (a * a + b * b) + (c * c + e * e);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.0f / d) + eps);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
    uint2* yr = reinterpret_cast<uint2*>(y + row * d);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const float4 g = __ldg(g4 + lane + 32 * j);
        const float4 b = __ldg(b4 + lane + 32 * j);
        const __nv_bfloat162 lo = __floats2bfloat162_rn((v[j].x - mean) * rstd * g.x + b.x,
                                                        (v[j].y - mean) * rstd * g.y + b.y);
        const __nv_bfloat162 hi = __floats2bfloat162_rn((v[j].z - mean) * rstd * g.z + b.z,
                                                        (v[j].w - mean) * rstd * g.w + b.w);
        uint2 o;
        o.x = *reinterpret_cast<const unsigned*>(&lo);
        o.y = *reinterpret_cast<const unsigned*>(&hi);
        yr[lane + 32 * j] = o;
    }
}

template <int NV>
cudaError_t launch(const __half* x, const float* g, const float* b, void* y, long long rows, float eps,
                   cudaStream_t stream, bool pdl = false) {
    const long long blocks = (rows + kWarpsPerBlock - 1) / kWarpsPerBlock;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)blocks);
    cfg.blockDim = dim3(kWarpsPerBlock * 32);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, layernorm_kernel<NV>, x, g, b, reinterpret_cast<__nv_bfloat16*>(y), rows, eps);
}

}  // namespace

bool layernorm_supports(int d) {
    if (d <= 0 || d % 128 != 0) return false;
    const int nv = d / 128;
    return (nv >= 1 && nv <= 6) || nv == 8 || nv == 10 || nv == 12 || nv == 16;
}

cudaError_t layernorm_launch(const void* x_f16, const float* gamma, const float* beta, void* y, long long rows, int d,
                             float eps, cudaStream_t stream) {
    return layernorm_launch_pdl(x_f16, gamma, beta, y, rows, d, eps, stream, false);
}

cudaError_t layernorm_launch_pdl(const void* x_f16, const float* gamma, const float* beta, void* y, long long rows, int d,
                                 float eps, cudaStream_t stream, bool pdl) {
    if (rows <= 0) return cudaSuccess;
    if (d % 128 != 0) return cudaErrorInvalidValue;
    const __half* x = reinterpret_cast<const __half*>(x_f16);
    switch (d / 128) {
        case 1: return launch<1>(x, gamma, beta, y, rows, eps, stream, pdl);
        case 2: return launch<2>(x, gamma, beta, y, rows, eps, stream, pdl);
        case 3: return launch<3>(x, gamma, beta, y, rows, eps, stream, pdl);
        case 4: return launch<4>(x, gamma, beta, y, rows, eps, stream, pdl);
        case 5: return launch<5>(x, gamma, beta, y, rows, eps, stream, pdl);
        case 6: return launch<6>(x, gamma, beta, y, rows, eps, stream, pdl);
        case 8: return launch<8>(x, gamma, beta, y, rows, eps, stream, pdl);
        case 10: return launch<10>(x, gamma, beta, y, rows, eps, stream, pdl);
        case 12: return launch<12>(x, gamma, beta, y, rows, eps, stream, pdl);
        case 16: return launch<16>(x, gamma, beta, y, rows, eps, stream, pdl);
    }
    return cudaErrorInvalidValue;
}

}  // namespace aries
