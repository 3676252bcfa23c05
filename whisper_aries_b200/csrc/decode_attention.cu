// Single-query attention of the decode step (row f1): out[b, h] = softmax(q . K^T / 8) V over a bf16 key/value cache.
//
// Replaces, per generated token, the self-attention over the decoder's own cache and the cross-attention over the
// encoder keys / values inside CT2 layers::MultiHeadAttention (cached mode) as driven by ctranslate2 Whisper.generate
// (SURVEY.md row f1).  HBM-bound: every key and value row of the (sequence, head) is read exactly once per step --
// for 64 sequences the cross-attention cache of large-v3 is 245 MB per sequence, 15.7 GB per step.
//   grid = (heads, batch, splits), 128 threads.  Eight lanes share one 128-byte key / value row (16 bytes each, full
//   sectors), a warp covers 4 rows per load instruction and four loads are kept in flight per lane.
//   pass 1: scores (log2 domain) -> shared memory + running max; pass 2: p = exp2(s - max), o += p V.
//   splits > 1 (cross-attention at small batch): partial (max, sum, o) per split go to a small f32 buffer and the
//   last CTA of a (b, h) to finish merges them (ticket counter, self-resetting).
// Self-attention also appends the step's new key / value row to the cache before attending (same CTA, so no race).
#include <cuda_bf16.h>

#include "skinny.h"

namespace aries {

namespace {

constexpr int kThreads = 128;
constexpr int kMaxKeys = 1536;            // shared score buffer (cross-attention without splitting: 1500 keys)

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}

__global__ void __launch_bounds__(kThreads) decode_attention_kernel(const DecAttnParams p) {
    __shared__ float s_score[kMaxKeys];
    __shared__ float s_red[4][66];
    __shared__ float s_max[4];
    __shared__ unsigned s_ticket;

    const int h = blockIdx.x, b = blockIdx.y, split = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = lane >> 3, sub = lane & 7;       // 8 lanes per key row; lane `sub` owns dims 8 sub .. 8 sub + 7

    pdl_wait();
    pdl_trigger();

    int n_keys = p.n_keys_fixed;
    const __nv_bfloat16* kbase = reinterpret_cast<const __nv_bfloat16*>(p.k) + (size_t)b * p.kv_rows * p.kv_ld + h * 64;
    const __nv_bfloat16* vbase = reinterpret_cast<const __nv_bfloat16*>(p.v) + (size_t)b * p.kv_rows * p.kv_ld + h * 64;
    if (n_keys == 0) {
        // self-attention: append this step's key / value, then attend to positions 0 .. step
        const int step = *p.step;
        n_keys = step + 1;
        if (tid < 16) {
            const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(tid < 8 ? p.new_k : p.new_v) +
                                       (size_t)b * p.new_ld + h * 64 + (tid & 7) * 8;
            __nv_bfloat16* dst = const_cast<__nv_bfloat16*>(tid < 8 ? kbase : vbase) + (size_t)step * p.kv_ld + (tid & 7) * 8;
            *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
        }
        __syncthreads();
    }
    const int per = (n_keys + p.splits - 1) / p.splits;
    const int j0 = split * per;
    const int j1 = (j0 + per < n_keys) ? j0 + per : n_keys;

    // query, pre-scaled by head_dim^-0.5 * log2(e)
    float q[8];
    {
        const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.q) + (size_t)b * p.q_ld +
                                                        h * 64 + sub * 8);
        unpack8(u, q);
        const float sc = 0.125f * 1.4426950408889634f;
#pragma unroll
        for (int i = 0; i < 8; ++i) q[i] *= sc;
    }

    // ---------------------------------------------------------------- pass 1: scores
    float m = -INFINITY;
    for (int jw = j0 + warp * 4; jw < j1; jw += 64) {          // warp-uniform trip count (shuffles inside)
        const int jb = jw + grp;
        uint4 kk[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = jb + 16 * u;
            kk[u] = (j < j1) ? *reinterpret_cast<const uint4*>(kbase + (size_t)j * p.kv_ld + sub * 8) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = jb + 16 * u;
            float f[8];
            unpack8(kk[u], f);
            float dot = 0.0f;
#pragma unroll
            for (int i = 0; i < 8; ++i) dot = fmaf(q[i], f[i], dot);
            dot += __shfl_xor_sync(0xffffffffu, dot, 1);
            dot += __shfl_xor_sync(0xffffffffu, dot, 2);
            dot += __shfl_xor_sync(0xffffffffu, dot, 4);
            if (j < j1) {
                if (sub == 0) s_score[j - j0] = dot;
                m = fmaxf(m, dot);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) s_max[warp] = m;
    __syncthreads();
    m = fmaxf(fmaxf(s_max[0], s_max[1]), fmaxf(s_max[2], s_max[3]));      // -inf only for an empty split

    // ---------------------------------------------------------------- pass 2: weights and weighted values
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float l = 0.0f;
    for (int jw = j0 + warp * 4; jw < j1; jw += 64) {
        const int jb = jw + grp;
        uint4 vv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = jb + 16 * u;
            vv[u] = (j < j1) ? *reinterpret_cast<const uint4*>(vbase + (size_t)j * p.kv_ld + sub * 8) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = jb + 16 * u;
            if (j < j1) {
                const float pj = exp2f(s_score[j - j0] - m);
                float f[8];
                unpack8(vv[u], f);
                l += pj;
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = fmaf(pj, f[i], acc[i]);
            }
        }
    }
    // the 4 row groups of a warp, then the 4 warps
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
        acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
    }
    l += __shfl_xor_sync(0xffffffffu, l, 8);
    l += __shfl_xor_sync(0xffffffffu, l, 16);
    if (grp == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) s_red[warp][sub * 8 + i] = acc[i];
        if (sub == 0) s_red[warp][64] = l;
    }
    __syncthreads();

    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)b * p.out_ld + h * 64;
    if (p.splits == 1) {
        if (tid < 64) {
            const float o = s_red[0][tid] + s_red[1][tid] + s_red[2][tid] + s_red[3][tid];
            const float ls = s_red[0][64] + s_red[1][64] + s_red[2][64] + s_red[3][64];
            out[tid] = __float2bfloat16_rn(o / ls);
        }
        return;
    }
    // ---------------------------------------------------------------- split merge
    float* part = p.partial + ((size_t)(b * p.heads + h) * p.splits) * 66;
    if (tid < 66) {
        float v;
        if (tid < 65) v = s_red[0][tid] + s_red[1][tid] + s_red[2][tid] + s_red[3][tid];
        else v = m;
        part[split * 66 + tid] = v;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(&p.tickets[b * p.heads + h], 1u);
    __syncthreads();
    if (s_ticket != (unsigned)(p.splits - 1)) return;
    __threadfence();
    if (tid < 64) {
        float M = -INFINITY;
        for (int s = 0; s < p.splits; ++s) M = fmaxf(M, __ldcg(part + s * 66 + 65));
        float o = 0.0f, ls = 0.0f;
        for (int s = 0; s < p.splits; ++s) {
            const float ms = __ldcg(part + s * 66 + 65);
            const float w = (ms == -INFINITY) ? 0.0f : exp2f(ms - M);
            o = fmaf(w, __ldcg(part + s * 66 + tid), o);
            ls = fmaf(w, __ldcg(part + s * 66 + 64), ls);
        }
        out[tid] = __float2bfloat16_rn(o / ls);
    }
    if (tid == 0) p.tickets[b * p.heads + h] = 0u;
}

}  // namespace

cudaError_t launch_maybe_pdl(const void* func, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, void** args,
                             bool pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelExC(&cfg, func, args);
}

cudaError_t decode_attention_launch(const DecAttnParams& p, cudaStream_t stream) {
    if (p.batch <= 0 || p.heads <= 0 || p.splits < 1) return cudaErrorInvalidValue;
    const int n_max = p.n_keys_fixed ? p.n_keys_fixed : (int)p.kv_rows;
    if ((n_max + p.splits - 1) / p.splits > kMaxKeys) return cudaErrorInvalidValue;
    if (p.n_keys_fixed == 0 && p.splits != 1) return cudaErrorInvalidValue;
    DecAttnParams q = p;
    void* args[] = {&q};
    return launch_maybe_pdl(reinterpret_cast<const void*>(decode_attention_kernel), dim3(p.heads, p.batch, p.splits),
                            dim3(kThreads), 0, stream, args, p.pdl != 0);
}

}  // namespace aries
