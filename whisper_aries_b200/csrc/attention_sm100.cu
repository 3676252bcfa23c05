// K6: fused non-causal self-attention for sm_100a (head_dim 64): softmax(Q K^T / 8) V without ever writing the
// [T, T] score matrix to HBM (CT2 runs this as batched GEMM -> softmax kernel -> batched GEMM).
//
// One CTA per (128-query tile, head, batch item), two CTAs resident per SM (97 KB of shared memory and 256 TMEM
// columns each).  Keys are walked in tiles of 64:
//   warp 0   lane 0: TMA for Q and the K ring (3 x 8 KB); lane 1: TMA for the V^T ring (3 x 8 KB)
//   warp 1   tcgen05.mma  S_j = Q K_j^T (M128 N64 K64) into one of TWO TMEM score buffers, issued two tiles ahead of
//            the softmax, and O += P_j V_j (M128 N64 K64) as soon as P_j is published
//   warps 2-5 one query row per thread: tcgen05.ld S_j (64 values, kept in registers), row max,
//            p = exp2(s*c - m_ref*c) in f32 (every 4th one as an FMA-pipe polynomial to unload the MUFU), bf16 P_j
//            into one of two 128B-swizzled K-major shared tiles.  O accumulates in TMEM and is rescaled
//            (tcgen05.ld / st) only when a row's maximum grew by more than 2^8 since the last rescale.
// Because S is double-buffered, the softmax warps never wait for the tensor pipe in steady state and the MMAs of
// tile j overlap the exponentials of tile j+1; the kernel is bound by MUFU + issue slots of the softmax warps.
// Q and K are read straight out of the QKV GEMM's row-major [B*T, 2d] output through a 3-D tensor map; V arrives
// pre-transposed ([B, h, 64, t_pad]) from that GEMM's epilogue so that both MMAs use K-major operands.
// Keys >= T are zero-filled by TMA and masked to -inf here; query rows >= T are computed and dropped.
#include <cstdlib>

#include "attention.h"
#include "ptx.cuh"

namespace aries {

namespace {

constexpr int kBlockQ = 128;
constexpr int kBlockKV = 64;
constexpr int kHeadDim = 64;
constexpr int kOCols = 80;                               // 64 output columns + the row-sum column (+ 15 of padding)
constexpr int kThreads = 192;
constexpr int kSoftmaxThreads = 128;
constexpr int kKVStages = 3;

constexpr int kQBytes = kBlockQ * kHeadDim * 2;          // 16 KB
constexpr int kKBytes = kBlockKV * kHeadDim * 2;         // 8 KB
constexpr int kVTmaBytes = kHeadDim * kBlockKV * 2;      // 8 KB written by TMA ...
constexpr int kVBytes = kOCols * kBlockKV * 2;           // ... + 2 KB constant tail: a row of ones and 15 rows of zeros
constexpr int kPBytes = kBlockQ * kBlockKV * 2;          // 16 KB (one 128B-swizzled K-major tile)
constexpr int kSmemBytes = kQBytes + kKVStages * (kKBytes + kVBytes) + 2 * kPBytes + 256 + 1024;   // 105,728 B
constexpr uint32_t kTmemCols = 256;                      // S0 [0,64) S1 [64,128) O [128,208); two CTAs per SM
constexpr float kScale = 0.18033688011112042f;           // log2(e) / sqrt(64)
constexpr float kRescaleThreshold = 8.0f;                // lazy rescale: only when the row max grew by > 2^8

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 2^x on the FMA / ALU pipes (no MUFU): round-to-nearest split x = n + f, |f| <= 0.5, degree-4 polynomial for 2^f
// (relative error < 5e-5, far below the bf16 rounding of P), exponent patched in with an integer add.
__device__ __forceinline__ float poly_exp2(float x) {
    x = fmaxf(x, -120.0f);
    const float t = x + 12582912.0f;                      // 1.5 * 2^23: the integer part lands in the low mantissa bits
    const float f = x - (t - 12582912.0f);
    float p = fmaf(9.6181291e-3f, f, 5.5504109e-2f);
    p = fmaf(p, f, 2.4022651e-1f);
    p = fmaf(p, f, 6.9314718e-1f);
    p = fmaf(p, f, 1.0f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// Row maximum of the 64 scores of one tile (tree-shaped so the compares are independent).
__device__ __forceinline__ float max64(const uint32_t (&s0)[32], const uint32_t (&s1)[32]) {
    float m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float a = fmaxf(__uint_as_float(s0[4 * k]), __uint_as_float(s0[4 * k + 1]));
        const float b = fmaxf(__uint_as_float(s0[4 * k + 2]), __uint_as_float(s0[4 * k + 3]));
        const float c = fmaxf(__uint_as_float(s1[4 * k]), __uint_as_float(s1[4 * k + 1]));
        const float d = fmaxf(__uint_as_float(s1[4 * k + 2]), __uint_as_float(s1[4 * k + 3]));
        m[k] = fmaxf(fmaxf(a, b), fmaxf(c, d));
    }
    return fmaxf(fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])), fmaxf(fmaxf(m[4], m[5]), fmaxf(m[6], m[7])));
}

// p = exp2(s * c - m * c) for 32 scores -> 16 packed bf16x2.  The scale / offset is a packed f32x2 FMA (two scores per
// instruction); with kPoly every 4th exponential runs on the FMA pipe instead of the MUFU.
// Compiler-only join point: all 32 values must exist before anything after it may be scheduled.  Without it ptxas
// interleaves every MUFU.EX2 pair with the F2FP that consumes it and, having only six scoreboard slots, keeps just a
// few exponentials in flight per warp (measured: XU pipe 52 % busy).  With it the 32 MUFUs issue back to back.
__device__ __forceinline__ void join32(float (&e)[32]) {
    asm volatile(""
                 : "+f"(e[0]), "+f"(e[1]), "+f"(e[2]), "+f"(e[3]), "+f"(e[4]), "+f"(e[5]), "+f"(e[6]), "+f"(e[7]),
                   "+f"(e[8]), "+f"(e[9]), "+f"(e[10]), "+f"(e[11]), "+f"(e[12]), "+f"(e[13]), "+f"(e[14]), "+f"(e[15]),
                   "+f"(e[16]), "+f"(e[17]), "+f"(e[18]), "+f"(e[19]), "+f"(e[20]), "+f"(e[21]), "+f"(e[22]),
                   "+f"(e[23]), "+f"(e[24]), "+f"(e[25]), "+f"(e[26]), "+f"(e[27]), "+f"(e[28]), "+f"(e[29]),
                   "+f"(e[30]), "+f"(e[31]));
}

// p = exp2(s * c - m * c) for 32 scores -> 16 packed bf16x2.  The scale / offset is a packed f32x2 FMA (two scores per
// instruction); with kPoly every 4th exponential runs on the FMA pipe instead of the MUFU.
template <bool kPoly>
__device__ __forceinline__ void exp_pack(const uint32_t (&s)[32], float neg_m, uint32_t (&out)[16]) {
    const float2 c2 = make_float2(kScale, kScale);
    const float2 m2 = make_float2(neg_m, neg_m);
    float e[32];
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
        const float2 x = __ffma2_rn(make_float2(__uint_as_float(s[i]), __uint_as_float(s[i + 1])), c2, m2);
        e[i] = x.x;
        e[i + 1] = x.y;
    }
    join32(e);
#pragma unroll
    for (int i = 0; i < 32; ++i) e[i] = (kPoly && (i & 7) == 7) ? poly_exp2(e[i]) : fast_exp2(e[i]);
    join32(e);
#pragma unroll
    for (int i = 0; i < 32; i += 2) out[i >> 1] = pack_bf16x2(e[i], e[i + 1]);
}

__device__ __forceinline__ void store_p(uint8_t* p_row, int swz, int chunk0, const uint32_t (&pk)[16]) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        uint4 q;
        q.x = pk[4 * g + 0]; q.y = pk[4 * g + 1]; q.z = pk[4 * g + 2]; q.w = pk[4 * g + 3];
        *reinterpret_cast<uint4*>(p_row + (((chunk0 + g) ^ swz) << 4)) = q;      // 16-byte chunk = 8 keys
    }
}

template <bool kPoly>
__global__ void __launch_bounds__(kThreads, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_vt, const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + kQBytes;                          // [stage][8 KB]
    uint8_t* sV = sK + kKVStages * kKBytes;              // [stage][10 KB]
    uint8_t* sP = sV + kKVStages * kVBytes;              // [2][16 KB]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kPBytes);
    uint64_t* q_full = bars;                             // 1
    uint64_t* k_full = bars + 1;                         // [3]
    uint64_t* v_full = bars + 4;                         // [3]
    uint64_t* k_empty = bars + 7;                        // [3]
    uint64_t* v_empty = bars + 10;                       // [3]
    uint64_t* s_full = bars + 13;                        // [2]
    uint64_t* p_full = bars + 15;                        // [2]
    uint64_t* pv_done = bars + 17;                       // [2]  O += P_j V_j retired (j & 1)
    uint64_t* o_full = bars + 19;                        // 1
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * kBlockQ;
    const int head = blockIdx.y;
    const int b = blockIdx.z;
    const int n_kv = (p.T + kBlockKV - 1) / kBlockKV;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_k);
        tma_prefetch_desc(&tmap_vt);
        mbar_init(q_full, 1);
        for (int s = 0; s < kKVStages; ++s) {
            mbar_init(&k_full[s], 1);
            mbar_init(&v_full[s], 1);
            mbar_init(&k_empty[s], 1);
            mbar_init(&v_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&s_full[s], 1);
            mbar_init(&p_full[s], kSoftmaxThreads);
            mbar_init(&pv_done[s], 1);
        }
        mbar_init(o_full, 1);
        fence_mbar_init();
    }
    // Constant tail of every V stage: B-operand rows 64..79 of the O MMA.  Row 64 is all ones, so column 64 of O
    // accumulates the row sums of (the bf16-rounded) P on the tensor core; rows 65..79 pad N to a legal 80.
    for (int i = threadIdx.x; i < kKVStages * 512; i += kThreads) {
        const int s = i >> 9, w = i & 511;               // 512 words per 2 KB tail
        reinterpret_cast<uint32_t*>(sV + s * kVBytes + kVTmaBytes)[w] = (w < 32) ? 0x3F803F80u : 0u;
    }
    fence_proxy_async_smem();
    if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_o = tmem_base + 128;

    if (warp == 0) {
        // ---------------------------------------------------------------- TMA producers: lane 0 = Q + K, lane 1 = V
        // (two independent rings: a K tile is consumed two softmax phases before the V tile of the same index)
        if (lane == 0) {
            mbar_expect_tx(q_full, kQBytes);
            tma_load_3d(sQ, &tmap_q, q_full, head * kHeadDim, q0, b);
            for (int j = 0; j < n_kv; ++j) {
                const int s = j % kKVStages;
                if (j >= kKVStages) mbar_wait_relaxed(&k_empty[s], (j / kKVStages - 1) & 1);
                mbar_expect_tx(&k_full[s], kKBytes);
                tma_load_3d(sK + s * kKBytes, &tmap_k, &k_full[s], p.d_model + head * kHeadDim, j * kBlockKV, b);
            }
        } else if (lane == 1) {
            for (int j = 0; j < n_kv; ++j) {
                const int s = j % kKVStages;
                if (j >= kKVStages) mbar_wait_relaxed(&v_empty[s], (j / kKVStages - 1) & 1);
                mbar_expect_tx(&v_full[s], kVTmaBytes);
                tma_load_3d(sV + s * kVBytes, &tmap_vt, &v_full[s], j * kBlockKV, head * kHeadDim, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---------------------------------------------------------------- MMA issuer
            constexpr uint32_t idesc_s = umma_idesc_bf16(kBlockQ, kBlockKV, false, false);
            constexpr uint32_t idesc_o = umma_idesc_bf16(kBlockQ, kOCols, false, false);
            constexpr uint64_t desc_hi = umma_smem_desc_hi(16, 1024);
            const uint32_t aQ = base;
            const uint32_t aK = aQ + kQBytes;
            const uint32_t aV = aK + kKVStages * kKBytes;
            const uint32_t aP = aV + kKVStages * kVBytes;
            auto issue_qk = [&](int j) {
                const int s = j % kKVStages;
                mbar_wait(&k_full[s], (j / kKVStages) & 1);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < kHeadDim / 16; ++k)
                    umma_bf16_ss(tmem_base + (j & 1) * kBlockKV, umma_smem_desc(aQ + k * 32, desc_hi),
                                 umma_smem_desc(aK + s * kKBytes + k * 32, desc_hi), idesc_s, k != 0);
                umma_commit(&k_empty[s]);
                umma_commit(&s_full[j & 1]);
            };
            mbar_wait(q_full, 0);
            issue_qk(0);
            if (n_kv > 1) issue_qk(1);
            for (int j = 0; j < n_kv; ++j) {
                const int s = j % kKVStages;
                // O (+)= P_j [V_j | 1]: P published (so S_j is fully read and O rescaled if it had to be)
                mbar_wait(&p_full[j & 1], (j >> 1) & 1);
                mbar_wait(&v_full[s], (j / kKVStages) & 1);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < kBlockKV / 16; ++k)
                    umma_bf16_ss(tmem_o, umma_smem_desc(aP + (j & 1) * kPBytes + k * 32, desc_hi),
                                 umma_smem_desc(aV + s * kVBytes + k * 32, desc_hi), idesc_o, (j > 0) || (k != 0));
                umma_commit(&v_empty[s]);
                umma_commit(&pv_done[j & 1]);
                if (j == n_kv - 1) umma_commit(o_full);
                if (j + 2 < n_kv) issue_qk(j + 2);       // S buffer j & 1 is free again
            }
        }
    } else {
        // -------------------------------------------------------------------- softmax warps: one query row per thread
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;                 // query row inside the tile == TMEM lane
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        float m_ref = 0.0f;
        const int swz = row & 7;
        uint8_t* p_row0 = sP + (row >> 3) * 1024 + (row & 7) * 128;

        for (int j = 0; j < n_kv; ++j) {
            const int buf = j & 1;
            // S_j ready.  The tensor pipe retires in issue order and S_j was issued after O += P_{j-2} V_{j-2}, so the
            // P buffer this tile overwrites is no longer being read.
            mbar_wait(&s_full[buf], (j >> 1) & 1);
            tc_fence_after();
            uint32_t s0[32], s1[32];
            tmem_ld_32x32b_x32(tmem_base + lane_addr + buf * kBlockKV, s0);
            tmem_ld_32x32b_x32(tmem_base + lane_addr + buf * kBlockKV + 32, s1);
            tmem_ld_wait_on(s0);
            tmem_ld_wait_on(s1);
            const int kv_valid = p.T - j * kBlockKV;
            if (kv_valid < kBlockKV) {                       // only the last tile: keys >= T do not exist
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (i >= kv_valid) s0[i] = 0xFF800000u;  // -inf
                    if (32 + i >= kv_valid) s1[i] = 0xFF800000u;
                }
            }
            uint32_t pk0[16], pk1[16];
            bool redo = (j == 0);
            if (j > 0) {
                // speculate that the running reference maximum still holds (it almost always does)
                exp_pack<kPoly>(s0, -m_ref * kScale, pk0);
                exp_pack<kPoly>(s1, -m_ref * kScale, pk1);
            }
            const float m_tile = max64(s0, s1);
            if (j == 0) {
                m_ref = m_tile;
            } else {
                // lazy rescale: keep exponentiating against a stale maximum until it is off by more than 2^8
                const bool need = (m_tile - m_ref) * kScale > kRescaleThreshold;
                if (__any_sync(0xffffffffu, need)) {
                    redo = true;
                    mbar_wait(&pv_done[buf ^ 1], ((j - 1) >> 1) & 1);      // O += P_{j-1} V_{j-1} has retired
                    tc_fence_after();
                    const float f = need ? fast_exp2((m_ref - m_tile) * kScale) : 1.0f;
                    if (need) m_ref = m_tile;
#pragma unroll 1
                    for (int c = 0; c < 3; ++c) {            // 96 columns cover the 80 of O (incl. the row sums)
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
            if (redo) {
                exp_pack<false>(s0, -m_ref * kScale, pk0);
                exp_pack<false>(s1, -m_ref * kScale, pk1);
            }
            uint8_t* p_row = p_row0 + buf * kPBytes;
            store_p(p_row, swz, 0, pk0);
            store_p(p_row, swz, 4, pk1);
            tc_fence_before();              // our TMEM reads / writes are complete and ordered before the arrive
            fence_proxy_async_smem();       // P visible to the tensor core's (async-proxy) reads
            mbar_arrive(&p_full[buf]);
        }

        mbar_wait(o_full, 0);
        tc_fence_after();
        const int t = q0 + row;
        uint32_t osum[32];
        tmem_ld_32x32b_x32(tmem_o + lane_addr + 64, osum);       // column 64 = row sum of P
        tmem_ld_wait_on(osum);
        const float inv = 1.0f / __uint_as_float(osum[0]);
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

bool use_poly_exp() {
    static const bool on = [] {
        const char* e = getenv("ARIES_ATTN_POLY");
        return e ? (e[0] != '0') : false;
    }();
    return on;
}

}  // namespace

cudaError_t attention_init_device() {
    cudaError_t e = cudaFuncSetAttribute(attention_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kSmemBytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(attention_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
}

cudaError_t attention_make_maps(const void* qk, const void* vt, int batch, int T, int d_model, int n_heads, int t_pad,
                                AttnMaps* maps) {
    const unsigned long long dims[3] = {(unsigned long long)(2 * d_model), (unsigned long long)T,
                                        (unsigned long long)batch};
    const unsigned long long strides[3] = {2, (unsigned long long)(2 * d_model) * 2,
                                           (unsigned long long)T * (2 * d_model) * 2};
    const unsigned box_q[3] = {64, (unsigned)kBlockQ, 1};
    const unsigned box_k[3] = {64, (unsigned)kBlockKV, 1};
    cudaError_t e = make_tmap_bf16(&maps->q, qk, 3, dims, strides, box_q);
    if (e != cudaSuccess) return e;
    if ((e = make_tmap_bf16(&maps->k, qk, 3, dims, strides, box_k)) != cudaSuccess) return e;
    const unsigned long long vdims[3] = {(unsigned long long)T, (unsigned long long)(n_heads * 64),
                                         (unsigned long long)batch};
    const unsigned long long vstrides[3] = {2, (unsigned long long)t_pad * 2,
                                            (unsigned long long)(n_heads * 64) * t_pad * 2};
    const unsigned box_v[3] = {(unsigned)kBlockKV, 64, 1};
    return make_tmap_bf16(&maps->vt, vt, 3, vdims, vstrides, box_v);
}

cudaError_t attention_launch(const AttnMaps& maps, const AttnParams& p, cudaStream_t stream) {
    if (p.d_model != p.n_heads * kHeadDim || p.T <= 0 || p.batch <= 0) return cudaErrorInvalidValue;
    dim3 grid((p.T + kBlockQ - 1) / kBlockQ, p.n_heads, p.batch);
    if (use_poly_exp())
        attention_fwd_kernel<true><<<grid, kThreads, kSmemBytes, stream>>>(maps.q, maps.k, maps.vt, p);
    else
        attention_fwd_kernel<false><<<grid, kThreads, kSmemBytes, stream>>>(maps.q, maps.k, maps.vt, p);
    return cudaGetLastError();
}

}  // namespace aries
