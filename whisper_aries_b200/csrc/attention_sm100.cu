// K6: fused non-causal self-attention for sm_100a (head_dim 64): softmax(Q K^T / 8) V without ever writing the
// [T, T] score matrix to HBM (CT2 runs this as batched GEMM -> softmax kernel -> batched GEMM).
//
// One CTA per (128-query tile, head, batch item), two CTAs resident per SM (83 KB of shared memory and 256 TMEM
// columns each).  Per 128-key tile j:
//   warp 0   TMA: K_j [128 x 64] and V^T_j [64 x 128] (two 64-wide boxes), single-buffered, refilled as soon as the
//            MMA that read them retires
//   warp 1   tcgen05.mma  S = Q K_j^T (M128 N128 K64) into TMEM, then O += P_j V_j (M128 N64 K128) into TMEM
//   warps 2-5 one query row per thread: tcgen05.ld S (pipelined), row max, p = exp2(s*c - m_ref*c) in f32, bf16 P_j
//            into shared memory in the 128B-swizzled K-major layout the second MMA reads.  O accumulates in TMEM;
//            it is rescaled (tcgen05.ld / st) only when a row's maximum grew by more than 2^8 since the last
//            rescale ("lazy rescale": p stays <= 256, exact after the final division by the row sum).
// Within a CTA the steps of a tile are serial; the second resident CTA fills the bubbles (the kernel is bound by
// the 16 exp2/clk/SM of the MUFU pipe and by instruction issue, not by the tensor pipe).
// Q and K are read straight out of the QKV GEMM's row-major [B*T, 2d] output through a 3-D tensor map; V arrives
// pre-transposed ([B, h, 64, t_pad]) from that GEMM's epilogue so that both MMAs use K-major operands.
// Keys >= T are zero-filled by TMA and masked to -inf here; query rows >= T are computed and dropped.
#include "attention.h"
#include "ptx.cuh"

namespace aries {

namespace {

constexpr int kBlockQ = 128;
constexpr int kBlockKV = 128;
constexpr int kHeadDim = 64;
constexpr int kThreads = 192;
constexpr int kSoftmaxThreads = 128;

constexpr int kQBytes = kBlockQ * kHeadDim * 2;          // 16 KB
constexpr int kKBytes = kBlockKV * kHeadDim * 2;         // 16 KB
constexpr int kVBytes = kHeadDim * kBlockKV * 2;         // 16 KB (two 8 KB boxes)
constexpr int kPBytes = kBlockQ * kBlockKV * 2;          // 32 KB (two 16 KB K-major sub-tiles)
constexpr int kSmemBytes = kQBytes + kKBytes + kVBytes + kPBytes + 256 + 1024;    // 83,200 B: two CTAs per SM
constexpr uint32_t kTmemCols = 256;                      // S: [0,128)  O: [128,192); two CTAs share the 512 columns
constexpr float kScale = 0.18033688011112042f;           // log2(e) / sqrt(64)
constexpr float kRescaleThreshold = 8.0f;                // lazy rescale: only when the row max grew by > 2^8

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// Row maximum of one 128-key score tile (TMEM -> registers, loads software-pipelined against the max).
template <bool kMasked>
__device__ __forceinline__ float row_max(uint32_t taddr, int kv_valid) {
    uint32_t a[32], b[32];
    float m = -INFINITY;
    tmem_ld_32x32b_x32(taddr, a);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t(&cur)[32] = (c & 1) ? b : a;
        uint32_t(&nxt)[32] = (c & 1) ? a : b;
        tmem_ld_wait_on(cur);
        if (c < 3) tmem_ld_32x32b_x32(taddr + (c + 1) * 32, nxt);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            float v = __uint_as_float(cur[i]);
            if (kMasked && c * 32 + i >= kv_valid) v = -INFINITY;
            m = fmaxf(m, v);
        }
    }
    return m;
}

// p = exp2(s * scale - m * scale) -> bf16 into the swizzled K-major A tile; returns the row sum of p (f32).
template <bool kMasked>
__device__ __forceinline__ float exp_and_store(uint32_t taddr, int kv_valid, float neg_m, uint8_t* p_row, int swz) {
    uint32_t a[32], b[32];
    float l0 = 0.0f, l1 = 0.0f;
    tmem_ld_32x32b_x32(taddr, a);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t(&cur)[32] = (c & 1) ? b : a;
        uint32_t(&nxt)[32] = (c & 1) ? a : b;
        tmem_ld_wait_on(cur);
        if (c < 3) tmem_ld_32x32b_x32(taddr + (c + 1) * 32, nxt);
        float pv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            float e = fast_exp2(fmaf(__uint_as_float(cur[i]), kScale, neg_m));
            if (kMasked && c * 32 + i >= kv_valid) e = 0.0f;
            pv[i] = e;
        }
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            l0 += pv[i];
            l1 += pv[i + 1];
        }
        uint8_t* dst = p_row + (c >> 1) * (kPBytes / 2);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int chunk = (c & 1) * 4 + g;           // 16-byte chunk (8 keys) inside the 64-key sub-tile
            uint4 q;
            q.x = pack_bf16x2(pv[8 * g + 0], pv[8 * g + 1]);
            q.y = pack_bf16x2(pv[8 * g + 2], pv[8 * g + 3]);
            q.z = pack_bf16x2(pv[8 * g + 4], pv[8 * g + 5]);
            q.w = pack_bf16x2(pv[8 * g + 6], pv[8 * g + 7]);
            *reinterpret_cast<uint4*>(dst + ((chunk ^ swz) << 4)) = q;
        }
    }
    return l0 + l1;
}

__global__ void __launch_bounds__(kThreads, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmap_qk, const __grid_constant__ CUtensorMap tmap_vt,
                     const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + kQBytes;
    uint8_t* sV = sK + kKBytes;
    uint8_t* sP = sV + kVBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + kPBytes);
    uint64_t* q_full = bars;
    uint64_t* k_full = bars + 1;
    uint64_t* v_full = bars + 2;
    uint64_t* k_empty = bars + 3;
    uint64_t* v_empty = bars + 4;
    uint64_t* s_full = bars + 5;
    uint64_t* p_full = bars + 6;
    uint64_t* o_full = bars + 7;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * kBlockQ;
    const int head = blockIdx.y;
    const int b = blockIdx.z;
    const int n_kv = (p.T + kBlockKV - 1) / kBlockKV;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_qk);
        tma_prefetch_desc(&tmap_vt);
        mbar_init(q_full, 1);
        mbar_init(k_full, 1);
        mbar_init(v_full, 1);
        mbar_init(k_empty, 1);
        mbar_init(v_empty, 1);
        mbar_init(s_full, 1);
        mbar_init(p_full, kSoftmaxThreads);
        mbar_init(o_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_s = tmem_base;
    const uint32_t tmem_o = tmem_base + 128;

    if (warp == 0) {
        if (lane == 0) {
            // ---------------------------------------------------------------- TMA producer
            // K and V are single-buffered: K_{j+1} is fetched as soon as S_j = Q K_j^T has retired, V_{j+1} as soon
            // as O += P_j V_j has; both then have a whole softmax phase to arrive.
            mbar_expect_tx(q_full, kQBytes);
            tma_load_3d(sQ, &tmap_qk, q_full, head * kHeadDim, q0, b);
            for (int j = 0; j < n_kv; ++j) {
                if (j > 0) mbar_wait_relaxed(k_empty, (j - 1) & 1);
                mbar_expect_tx(k_full, kKBytes);
                tma_load_3d(sK, &tmap_qk, k_full, p.d_model + head * kHeadDim, j * kBlockKV, b);
                if (j > 0) mbar_wait_relaxed(v_empty, (j - 1) & 1);
                mbar_expect_tx(v_full, kVBytes);
                tma_load_3d(sV, &tmap_vt, v_full, j * kBlockKV, head * kHeadDim, b);
                tma_load_3d(sV + kVBytes / 2, &tmap_vt, v_full, j * kBlockKV + 64, head * kHeadDim, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---------------------------------------------------------------- MMA issuer
            constexpr uint32_t idesc_s = umma_idesc_bf16(kBlockQ, kBlockKV, false, false);
            constexpr uint32_t idesc_o = umma_idesc_bf16(kBlockQ, kHeadDim, false, false);
            constexpr uint64_t desc_hi = umma_smem_desc_hi(16, 1024);
            const uint32_t aQ = base;
            const uint32_t aK = aQ + kQBytes;
            const uint32_t aV = aK + kKBytes;
            const uint32_t aP = aV + kVBytes;
            mbar_wait(q_full, 0);
            for (int j = 0; j <= n_kv; ++j) {
                if (j > 0) {
                    // O (+)= P_{j-1} V_{j-1}: P published, S_{j-1} fully read, O rescaled if it had to be
                    mbar_wait(p_full, (j - 1) & 1);
                    mbar_wait(v_full, (j - 1) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < kBlockKV / 16; ++k) {
                        const uint32_t sub = (k >> 2), kk = (k & 3);
                        umma_bf16_ss(tmem_o, umma_smem_desc(aP + sub * (kPBytes / 2) + kk * 32, desc_hi),
                                     umma_smem_desc(aV + sub * (kVBytes / 2) + kk * 32, desc_hi), idesc_o,
                                     (j > 1) || (k != 0));
                    }
                    umma_commit(v_empty);
                    if (j == n_kv) umma_commit(o_full);
                }
                if (j < n_kv) {
                    mbar_wait(k_full, j & 1);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < kHeadDim / 16; ++k)
                        umma_bf16_ss(tmem_s, umma_smem_desc(aQ + k * 32, desc_hi), umma_smem_desc(aK + k * 32, desc_hi),
                                     idesc_s, k != 0);
                    umma_commit(k_empty);
                    umma_commit(s_full);
                }
            }
        }
    } else {
        // -------------------------------------------------------------------- softmax warps: one query row per thread
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;                 // query row inside the tile == TMEM lane
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        float m_ref = 0.0f, l_run = 0.0f;
        uint8_t* p_row = sP + (row >> 3) * 1024 + (row & 7) * 128;
        const int swz = row & 7;

        for (int j = 0; j < n_kv; ++j) {
            mbar_wait(s_full, j & 1);        // S_j ready; tensor-pipe order also guarantees O += P_{j-1} V_{j-1} is done
            tc_fence_after();
            const int kv_valid = p.T - j * kBlockKV;
            const bool masked = kv_valid < kBlockKV;         // only the last tile
            const float m_tile = masked ? row_max<true>(tmem_s + lane_addr, kv_valid)
                                        : row_max<false>(tmem_s + lane_addr, kv_valid);
            if (j == 0) {
                m_ref = m_tile;
            } else {
                // lazy rescale: keep exponentiating against a stale maximum until it is off by more than 2^8
                const bool need = (m_tile - m_ref) * kScale > kRescaleThreshold;
                if (__any_sync(0xffffffffu, need)) {
                    const float f = need ? fast_exp2((m_ref - m_tile) * kScale) : 1.0f;
                    if (need) m_ref = m_tile;
                    l_run *= f;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        uint32_t o[32];
                        tmem_ld_32x32b_x32(tmem_o + lane_addr + c * 32, o);
                        tmem_ld_wait_on(o);
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
                        tmem_st_32x32b_x32(tmem_o + lane_addr + c * 32, o);
                    }
                    tmem_st_wait();
                }
            }
            const float neg_m = -m_ref * kScale;
            l_run += masked ? exp_and_store<true>(tmem_s + lane_addr, kv_valid, neg_m, p_row, swz)
                            : exp_and_store<false>(tmem_s + lane_addr, kv_valid, neg_m, p_row, swz);
            tc_fence_before();              // our TMEM reads / writes are complete and ordered before the arrive
            fence_proxy_async_smem();       // P visible to the tensor core's (async-proxy) reads
            mbar_arrive(p_full);
        }

        mbar_wait(o_full, 0);
        tc_fence_after();
        const int t = q0 + row;
        const float inv = 1.0f / l_run;
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + ((long long)b * p.T + t) * p.d_model +
                             head * kHeadDim;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(tmem_o + lane_addr + c * 32, o);
            tmem_ld_wait_on(o);
            if (t < p.T) {
                uint4* d4 = reinterpret_cast<uint4*>(dst + c * 32);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 q;
                    q.x = pack_bf16x2(__uint_as_float(o[8 * g + 0]) * inv, __uint_as_float(o[8 * g + 1]) * inv);
                    q.y = pack_bf16x2(__uint_as_float(o[8 * g + 2]) * inv, __uint_as_float(o[8 * g + 3]) * inv);
                    q.z = pack_bf16x2(__uint_as_float(o[8 * g + 4]) * inv, __uint_as_float(o[8 * g + 5]) * inv);
                    q.w = pack_bf16x2(__uint_as_float(o[8 * g + 6]) * inv, __uint_as_float(o[8 * g + 7]) * inv);
                    d4[g] = q;
                }
            }
        }
        tc_fence_before();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<kTmemCols>(tmem_base);
    }
}

}  // namespace

cudaError_t attention_init_device() {
    return cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
}

cudaError_t attention_make_maps(const void* qk, const void* vt, int batch, int T, int d_model, int n_heads, int t_pad,
                                CUtensorMap* map_qk, CUtensorMap* map_vt) {
    {
        const unsigned long long dims[3] = {(unsigned long long)(2 * d_model), (unsigned long long)T,
                                            (unsigned long long)batch};
        const unsigned long long strides[3] = {2, (unsigned long long)(2 * d_model) * 2,
                                               (unsigned long long)T * (2 * d_model) * 2};
        const unsigned box[3] = {64, 128, 1};
        cudaError_t e = make_tmap_bf16(map_qk, qk, 3, dims, strides, box);
        if (e != cudaSuccess) return e;
    }
    {
        const unsigned long long dims[3] = {(unsigned long long)T, (unsigned long long)(n_heads * 64),
                                            (unsigned long long)batch};
        const unsigned long long strides[3] = {2, (unsigned long long)t_pad * 2,
                                               (unsigned long long)(n_heads * 64) * t_pad * 2};
        const unsigned box[3] = {64, 64, 1};
        return make_tmap_bf16(map_vt, vt, 3, dims, strides, box);
    }
}

cudaError_t attention_launch(const CUtensorMap& map_qk, const CUtensorMap& map_vt, const AttnParams& p,
                             cudaStream_t stream) {
    if (p.d_model != p.n_heads * kHeadDim || p.T <= 0 || p.batch <= 0) return cudaErrorInvalidValue;
    dim3 grid((p.T + kBlockQ - 1) / kBlockQ, p.n_heads, p.batch);
    attention_fwd_kernel<<<grid, kThreads, kSmemBytes, stream>>>(map_qk, map_vt, p);
    return cudaGetLastError();
}

}  // namespace aries
