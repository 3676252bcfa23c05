// Single-query attention of the decode step (row f1): out[b, h] = softmax(q . K^T / 8) V over a bf16 key/value cache.
//
// Replaces, per generated token, the self-attention over the decoder's own cache and the cross-attention over the
// encoder keys / values inside CT2 layers::MultiHeadAttention (cached mode) as driven by ctranslate2 Whisper.generate
// (SURVEY.md row f1).  HBM-bound: every key and value row of the (sequence, head) is read exactly once per step --
// for 64 sequences the cross-attention cache of large-v3 is 245 MB per sequence, 15.7 GB per step.
//   grid = (heads, batch, splits), 128 threads.  Eight lanes share one 128-byte key / value row (16 bytes each, full
//   sectors), a warp covers 4 rows per load instruction and four loads are kept in flight per lane.
//   pass 1: scores (log2 domain) -> shared memory + running max; pass 2: p = exp2(s - max), o += p V.
//   splits > 1 (cross-attention at small batch): the CTAs of a (b, h) form a thread-block cluster (grid.z = cluster
//   size <= 8); each parks its partial (max, sum, o) in its own shared memory and rank 0 merges them through
//   distributed shared memory -- no HBM round trip, no atomics.  A share of <= 192 keys also fetches its value rows
//   together with the keys (one HBM round trip instead of two).
// Self-attention also appends the step's new key / value row to the cache before attending (same CTA, so no race).
#include <cuda_bf16.h>

#include "skinny.h"

namespace aries {

namespace {

constexpr int kThreads = 128;
constexpr int kMaxKeys = 1536;            // shared score buffer (cross-attention without splitting: 1500 keys)

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_peer_f32(const float* local, uint32_t rank) {
    uint32_t remote;
    float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"((uint32_t)__cvta_generic_to_shared(local)), "r"(rank));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote));
    return v;
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}

template <bool kPrefetchV>
__global__ void __launch_bounds__(kThreads) decode_attention_kernel(const DecAttnParams p) {
    __shared__ float s_score[kMaxKeys];
    __shared__ float s_red[4][66];
    __shared__ float s_max[4];

    const int h = blockIdx.x, b = blockIdx.y, split = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = lane >> 3, sub = lane & 7;       // 8 lanes per key row; lane `sub` owns dims 8 sub .. 8 sub + 7

    pdl_wait();
    // A window that has emitted EOT only repeats EOT (decode_sample_kernel): it stops reading its caches -- 245.8 MB per
    // step for large-v3's cross-attention.  All CTAs of a cluster share b, so they leave together.
    if (p.done != nullptr && p.done[b] != 0) {
        pdl_trigger();
        return;
    }

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
    // A share of <= 64 keys (cross-attention split finely at small batch, early self-attention steps) is ONE iteration:
    // its value rows are fetched together with the keys, so the CTA pays one HBM round trip instead of two.
    constexpr int kPre = kPrefetchV ? 3 : 1;                 // iterations whose value rows are fetched with the keys
    const bool pre = kPrefetchV && (j1 - j0) <= 64 * kPre;   // (a template flag: the prefetch costs registers)
    uint4 vpre[kPre][4];
    float m = -INFINITY;
    int it = 0;
    for (int jw = j0 + warp * 4; jw < j1; jw += 64, ++it) {  // warp-uniform trip count (shuffles inside)
        const int jb = jw + grp;
        uint4 kk[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = jb + 16 * u;
            kk[u] = (j < j1) ? *reinterpret_cast<const uint4*>(kbase + (size_t)j * p.kv_ld + sub * 8) : make_uint4(0, 0, 0, 0);
        }
        if (pre) {
#pragma unroll
            for (int i = 0; i < kPre; ++i)
                if (i == it) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int j = jb + 16 * u;
                        vpre[i][u] = (j < j1) ? *reinterpret_cast<const uint4*>(vbase + (size_t)j * p.kv_ld + sub * 8)
                                              : make_uint4(0, 0, 0, 0);
                    }
                }
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
    it = 0;
    for (int jw = j0 + warp * 4; jw < j1; jw += 64, ++it) {
        const int jb = jw + grp;
        uint4 vv[4];
        if (pre) {
#pragma unroll
            for (int i = 0; i < kPre; ++i)
                if (i == it) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) vv[u] = vpre[i][u];
                }
        } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = jb + 16 * u;
                vv[u] = (j < j1) ? *reinterpret_cast<const uint4*>(vbase + (size_t)j * p.kv_ld + sub * 8) : make_uint4(0, 0, 0, 0);
            }
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
    pdl_trigger();       // after the streaming part: dependents launched earlier would only occupy SM slots
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
    // ---------------------------------------------------------------- split merge (cluster, distributed shared memory)
    __shared__ float s_part[66];
    if (tid < 65) s_part[tid] = s_red[0][tid] + s_red[1][tid] + s_red[2][tid] + s_red[3][tid];
    if (tid == 65) s_part[65] = m;
    cluster_sync_all();                                   // every CTA's partial is in its shared memory
    if (split == 0 && tid < 64) {
        float ms[8], ls[8], os[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const uint32_t r = (uint32_t)(s < p.splits ? s : 0);
            ms[s] = ld_peer_f32(&s_part[65], r);
            ls[s] = ld_peer_f32(&s_part[64], r);
            os[s] = ld_peer_f32(&s_part[tid], r);
        }
        float M = -INFINITY;
#pragma unroll
        for (int s = 0; s < 8; ++s)
            if (s < p.splits) M = fmaxf(M, ms[s]);
        float o = 0.0f, lsum = 0.0f;
#pragma unroll
        for (int s = 0; s < 8; ++s)
            if (s < p.splits && ms[s] != -INFINITY) {
                const float w = exp2f(ms[s] - M);
                o = fmaf(w, os[s], o);
                lsum = fmaf(w, ls[s], lsum);
            }
        out[tid] = __float2bfloat16_rn(o / lsum);
    }
    cluster_sync_all();                                   // rank 0 is done reading its peers
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
    if (p.batch <= 0 || p.heads <= 0 || p.splits < 1 || p.splits > 8) return cudaErrorInvalidValue;
    const int n_max = p.n_keys_fixed ? p.n_keys_fixed : (int)p.kv_rows;
    if ((n_max + p.splits - 1) / p.splits > kMaxKeys) return cudaErrorInvalidValue;
    if (p.n_keys_fixed == 0 && p.splits != 1) return cudaErrorInvalidValue;
    const bool prefetch = p.n_keys_fixed > 0 && (p.n_keys_fixed + p.splits - 1) / p.splits <= 192;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.heads, p.batch, p.splits);
    cfg.blockDim = dim3(kThreads);
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (p.splits > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 1;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = p.splits;
        ++na;
    }
    if (p.pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    return prefetch ? cudaLaunchKernelEx(&cfg, decode_attention_kernel<true>, p)
                    : cudaLaunchKernelEx(&cfg, decode_attention_kernel<false>, p);
}

}  // namespace aries
