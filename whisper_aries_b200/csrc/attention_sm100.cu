// K6: fused non-causal self-attention for sm_100a (head_dim 64): softmax(Q K^T / 8) V without ever writing the
// [T, T] score matrix to HBM (CT2 runs this as batched GEMM -> softmax kernel -> batched GEMM).
//
// One CTA per (128-query tile, head, batch item), two CTAs resident per SM.  Per 128-key tile j:
//   warp 0   TMA: K_j [128 x 64] and V^T_j [64 x 128] (two 64-wide boxes) into a 2-stage ring
//   warp 1   tcgen05.mma  S = Q K_j^T (M128 N128 K64) into TMEM, later O_j = P_j V_j (M128 N64 K128) into TMEM
//   warps 2-5 one query row per thread: tcgen05.ld S, online softmax in f32 registers (exp2 on pre-scaled scores),
//            P_j -> bf16 -> shared memory in the 128B-swizzled K-major layout the MMA reads, then
//            o = o * alpha_j + O_j from TMEM.  The running output lives in registers, so TMEM never needs a
//            read-modify-write correction pass.
// Within a CTA the three steps of a tile are serial; the second resident CTA fills the bubbles (the kernel is bound
// by the 16 exp2/clk/SM of the MUFU pipe, not by the tensor pipe).
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
constexpr int kStages = 2;
constexpr int kThreads = 192;
constexpr int kSoftmaxThreads = 128;

constexpr int kQBytes = kBlockQ * kHeadDim * 2;          // 16 KB
constexpr int kKBytes = kBlockKV * kHeadDim * 2;         // 16 KB
constexpr int kVBytes = kHeadDim * kBlockKV * 2;         // 16 KB (two 8 KB boxes)
constexpr int kPBytes = kBlockQ * kBlockKV * 2;          // 32 KB (two 16 KB K-major sub-tiles)
constexpr int kSmemBytes = kQBytes + kStages * (kKBytes + kVBytes) + kPBytes + 256 + 1024;
constexpr uint32_t kTmemCols = 256;                      // S: [0,128)  O: [128,192)

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

__global__ void __launch_bounds__(kThreads, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmap_qk, const __grid_constant__ CUtensorMap tmap_vt,
                     const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + kQBytes;                          // [stage][16 KB]
    uint8_t* sV = sK + kStages * kKBytes;                // [stage][16 KB]
    uint8_t* sP = sV + kStages * kVBytes;                // 32 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + kPBytes);
    uint64_t* q_full = bars;                // 1
    uint64_t* k_full = bars + 1;            // [2]
    uint64_t* v_full = bars + 3;            // [2]
    uint64_t* kv_empty = bars + 5;          // [2]
    uint64_t* s_full = bars + 7;            // 1
    uint64_t* p_full = bars + 8;            // 1
    uint64_t* o_full = bars + 9;            // 1
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

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
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&k_full[s], 1);
            mbar_init(&v_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
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
            mbar_expect_tx(q_full, kQBytes);
            tma_load_3d(sQ, &tmap_qk, q_full, head * kHeadDim, q0, b);
            for (int j = 0; j < n_kv; ++j) {
                const int s = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                mbar_wait(&kv_empty[s], ph ^ 1);
                mbar_expect_tx(&k_full[s], kKBytes);
                tma_load_3d(sK + s * kKBytes, &tmap_qk, &k_full[s], p.d_model + head * kHeadDim, j * kBlockKV, b);
                mbar_expect_tx(&v_full[s], kVBytes);
                tma_load_3d(sV + s * kVBytes, &tmap_vt, &v_full[s], j * kBlockKV, head * kHeadDim, b);
                tma_load_3d(sV + s * kVBytes + kVBytes / 2, &tmap_vt, &v_full[s], j * kBlockKV + 64, head * kHeadDim,
                            b);
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
            const uint32_t aV = aK + kStages * kKBytes;
            const uint32_t aP = aV + kStages * kVBytes;
            mbar_wait(q_full, 0);
            mbar_wait(&k_full[0], 0);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < kHeadDim / 16; ++k)
                umma_bf16_ss(tmem_s, umma_smem_desc(aQ + k * 32, desc_hi), umma_smem_desc(aK + k * 32, desc_hi),
                             idesc_s, k != 0);
            umma_commit(s_full);
            for (int j = 0; j < n_kv; ++j) {
                const int s = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                // O_j = P_j V_j once the softmax warps have published P_j (they have also finished reading S_j and
                // O_{j-1} by then, so both TMEM regions may be overwritten)
                mbar_wait(p_full, j & 1);
                mbar_wait(&v_full[s], ph);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < kBlockKV / 16; ++k) {
                    const uint32_t sub = (k >> 2), kk = (k & 3);
                    umma_bf16_ss(tmem_o, umma_smem_desc(aP + sub * (kPBytes / 2) + kk * 32, desc_hi),
                                 umma_smem_desc(aV + s * kVBytes + sub * (kVBytes / 2) + kk * 32, desc_hi), idesc_o,
                                 k != 0);
                }
                umma_commit(&kv_empty[s]);       // K_j / V_j slot reusable once these MMAs retire
                umma_commit(o_full);
                if (j + 1 < n_kv) {
                    const int s1 = (j + 1) & 1;
                    const uint32_t ph1 = ((j + 1) >> 1) & 1;
                    mbar_wait(&k_full[s1], ph1);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < kHeadDim / 16; ++k)
                        umma_bf16_ss(tmem_s, umma_smem_desc(aQ + k * 32, desc_hi),
                                     umma_smem_desc(aK + s1 * kKBytes + k * 32, desc_hi), idesc_s, k != 0);
                    umma_commit(s_full);
                }
            }
        }
    } else {
        // -------------------------------------------------------------------- softmax / output warps
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;                 // query row inside the tile == TMEM lane
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const float scale = 0.18033688011112042f;            // log2(e) / sqrt(64)
        float m_run = -INFINITY, l_run = 0.0f;
        float o_acc[kHeadDim];
#pragma unroll
        for (int i = 0; i < kHeadDim; ++i) o_acc[i] = 0.0f;
        uint8_t* p_row = sP + (row >> 3) * 1024 + (row & 7) * 128;
        const int swz = row & 7;

        for (int j = 0; j < n_kv; ++j) {
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            const int kv_valid = p.T - j * kBlockKV;         // >= 128 except on the last tile
            // pass 1: row maximum
            float m_tile = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t sr[32];
                tmem_ld_32x32b_x32(tmem_s + lane_addr + c * 32, sr);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float v = (c * 32 + i < kv_valid) ? __uint_as_float(sr[i]) : -INFINITY;
                    m_tile = fmaxf(m_tile, v);
                }
            }
            const float m_new = fmaxf(m_run, m_tile);
            const float alpha = fast_exp2((m_run - m_new) * scale);
            const float neg_m = -m_new * scale;
            float l_tile = 0.0f;
            // pass 2: p = exp2(s * scale - m * scale), bf16, into the swizzled A-operand tile
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t sr[32];
                tmem_ld_32x32b_x32(tmem_s + lane_addr + c * 32, sr);
                tmem_ld_wait();
                float pv[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float e = fast_exp2(fmaf(__uint_as_float(sr[i]), scale, neg_m));
                    pv[i] = (c * 32 + i < kv_valid) ? e : 0.0f;
                    l_tile += pv[i];
                }
                uint8_t* dst = p_row + (c >> 1) * (kPBytes / 2);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int chunk = (c & 1) * 4 + g;       // 16-byte chunk (8 keys) inside the 64-key sub-tile
                    uint4 q;
                    q.x = pack_bf16x2(pv[8 * g + 0], pv[8 * g + 1]);
                    q.y = pack_bf16x2(pv[8 * g + 2], pv[8 * g + 3]);
                    q.z = pack_bf16x2(pv[8 * g + 4], pv[8 * g + 5]);
                    q.w = pack_bf16x2(pv[8 * g + 6], pv[8 * g + 7]);
                    *reinterpret_cast<uint4*>(dst + ((chunk ^ swz) << 4)) = q;
                }
            }
            l_run = l_run * alpha + l_tile;
            m_run = m_new;
            tc_fence_before();              // our tcgen05.ld of S are complete (wait::ld) and ordered before the arrive
            fence_proxy_async_smem();       // P visible to the tensor core's (async-proxy) reads
            mbar_arrive(p_full);

            // o = o * alpha + P_j V_j
            mbar_wait(o_full, j & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t orr[32];
                tmem_ld_32x32b_x32(tmem_o + lane_addr + c * 32, orr);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] = fmaf(o_acc[c * 32 + i], alpha, __uint_as_float(orr[i]));
            }
            tc_fence_before();
        }

        const int t = q0 + row;
        if (t < p.T) {
            const float inv = 1.0f / l_run;
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + ((long long)b * p.T + t) * p.d_model +
                                 head * kHeadDim;
            uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                uint4 q;
                q.x = pack_bf16x2(o_acc[8 * g + 0] * inv, o_acc[8 * g + 1] * inv);
                q.y = pack_bf16x2(o_acc[8 * g + 2] * inv, o_acc[8 * g + 3] * inv);
                q.z = pack_bf16x2(o_acc[8 * g + 4] * inv, o_acc[8 * g + 5] * inv);
                q.w = pack_bf16x2(o_acc[8 * g + 6] * inv, o_acc[8 * g + 7] * inv);
                d4[g] = q;
            }
        }
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
