// K6: fused non-causal self-attention for sm_100a (head_dim 64): softmax(Q K^T / 8) V without ever writing the
// [T, T] score matrix to HBM (CT2 runs this as batched GEMM -> softmax kernel -> batched GEMM).
//
// One CTA per (256-query block, head, batch item), one CTA per SM (416 of the 512 TMEM columns, 138 KB of shared
// memory).  The 256 queries are two tiles A and B of 128 rows that share every K / V tile; keys are walked in tiles
// of 64 and every query tile owns TWO score buffers in tensor memory:
//   warps 0-3  softmax of tile A, warps 4-7 softmax of tile B: one query row per thread (row == TMEM lane):
//              tcgen05.ld S_j (64 values, kept in registers), row max, p = exp2(s*c - m_ref*c) in f32, bf16 P_j
//              written back INTO the first 32 columns of the buffer S_j occupied (tcgen05.st) -- P never touches
//              shared memory.  O accumulates in TMEM and is rescaled (tcgen05.ld / st) only when a row's maximum
//              grew by more than 2^8 since the last rescale.
//   warp 8     lane 0: TMA for both Q tiles and the K ring (4 x 8 KB); lane 1: TMA for the V^T ring (4 x 8 KB)
//   warps 9-10 tcgen05.mma, one elected thread per query tile:  S_g = Q_g K_j^T (SS, M128 N64 K64), issued TWO key tiles ahead of the
//              softmax, and O_g += P_g [V_j | 1] (TS: the A operand is P_g in tensor memory; M128 N80 K64).  Row 64
//              of the B operand is all ones, so column 64 of O accumulates the row sums of the bf16-rounded P on
//              the tensor core.
// Because S is double-buffered the softmax warps never wait for the tensor pipe in steady state.  Warp i of tile A
// and warp i of tile B share one scheduler and its MUFU -- the bound of this kernel at head_dim 64.  Measured with the
// clock64 timeline (tests/attn_trace.py): the two run in lockstep, ~950 cycles in which their 128 MUFU instructions
// saturate the unit (7.4 cycles each) and ~630 cycles of barrier / TMEM round trips (each ~170 cycles through the
// scheduler's in-order MIO queue) with the MUFU idle.  Forcing them into anti-phase with a token (named barriers or
// mbarriers, early hand-off) was tried and is slower: ONE warp cannot keep the MUFU busy (13 cycles per instruction).
// Q and K are read straight out of the QKV GEMM's row-major [B*T, 2d] output through a 3-D tensor map; V arrives
// pre-transposed ([B, h, 64, t_pad]) from that GEMM's epilogue so that both MMAs use K-major operands.
// Keys >= T are zero-filled by TMA and masked to -inf here; query rows >= T are computed and dropped.
#include <cstdlib>
#include <type_traits>

#include "attention.h"
#include "ptx.cuh"

namespace aries {

namespace {

constexpr int kBlockQ = 128;                             // rows per query tile (= TMEM lanes)
#ifndef ARIES_ATTN_QTILES
#define ARIES_ATTN_QTILES 1
#endif
constexpr int kQTiles = ARIES_ATTN_QTILES;               // query tiles per CTA (1: two CTAs per SM, 2: one)
constexpr int kCtasPerSm = 2 / kQTiles;
constexpr int kBlockKV = 64;
constexpr int kHeadDim = 64;
constexpr int kOCols = 80;                               // 64 output columns + the row-sum column (+ 15 of padding)
constexpr int kSoftmaxThreads = kBlockQ;                 // per query tile
constexpr int kThreads = kQTiles * kSoftmaxThreads + 32 + kQTiles * 32;
constexpr int kTmaWarp = kQTiles * 4;
constexpr int kMmaWarp = kTmaWarp + 1;                   // + g: one issuing warp per query tile
constexpr int kKStages = 4;
constexpr int kVStages = 4;

constexpr int kQTileBytes = kBlockQ * kHeadDim * 2;      // 16 KB
constexpr int kKBytes = kBlockKV * kHeadDim * 2;         // 8 KB (64 keys x 128 B, one 128B-swizzled K-major tile)
constexpr int kVTmaBytes = kHeadDim * kBlockKV * 2;      // 8 KB written by TMA ...
constexpr int kVBytes = kOCols * kBlockKV * 2;           // ... + 2 KB constant tail: a row of ones and 15 rows of zeros
constexpr int kSmemBytes = kQTiles * kQTileBytes + kKStages * kKBytes + kVStages * kVBytes + 256 + 1024;   // 107,776 B
constexpr uint32_t kTmemCols = 256 * kQTiles;            // per tile g: S0 [128g, +64) S1 [128g+64, +64); O_g [kTmemO+80g, +80)
constexpr uint32_t kTmemO = 128 * kQTiles;
constexpr float kScale = 0.18033688011112042f;           // log2(e) / sqrt(64)
constexpr int kDefaultPoly = 104;                        // ARIES_ATTN_POLY overrides (0, 4, 8, 102 .. 108): 1/4 of the exponentials as packed polynomials, measured best (tests/attn_ab.py)
constexpr float kRescaleThreshold = 8.0f;                // lazy rescale: only when the row max grew by > 2^8

// Test-only timeline (variant bit 2, ARIES_ATTN_TRACE=1): clock64 stamps of lane 0 of every warp of a few CTAs.
constexpr int kTraceCtas = 16, kTraceWarps = 11, kTraceEvents = 160;
__device__ unsigned long long g_attn_trace[kTraceCtas * kTraceWarps * kTraceEvents];

struct Tracer {
    unsigned long long* p;
    int n;
    __device__ __forceinline__ void stamp() {
        if (p != nullptr && n < kTraceEvents) p[n++] = clock64();
    }
};

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

// Maximum of 32 scores (tree-shaped so the compares are independent).
__device__ __forceinline__ float max32(const uint32_t (&s)[32]) {
    float m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
        m[k] = fmaxf(fmaxf(__uint_as_float(s[4 * k]), __uint_as_float(s[4 * k + 1])),
                     fmaxf(__uint_as_float(s[4 * k + 2]), __uint_as_float(s[4 * k + 3])));
    return fmaxf(fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])), fmaxf(fmaxf(m[4], m[5]), fmaxf(m[6], m[7])));
}

// keys >= kv_valid do not exist: score -> -inf (only the last key tile takes this path)
__device__ __forceinline__ void mask32(uint32_t (&s)[32], int first_col, int kv_valid) {
#pragma unroll
    for (int i = 0; i < 32; ++i)
        if (first_col + i >= kv_valid) s[i] = 0xFF800000u;
}

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

// p = exp2(s * c - m * c) for 32 scores -> 16 packed bf16x2 (low half = the lower key index).  The scale / offset is
// a packed f32x2 FMA (two scores per instruction); with kPoly every 8th exponential runs on the FMA pipe instead of
// the MUFU.
// Two exponentials at once on the FMA pipe: the range reduction and the polynomial are packed f32x2 instructions
// (5.5 issue slots per element against ~10 for the scalar form).
__device__ __forceinline__ float2 poly_exp2x2(float2 x) {
    x.x = fmaxf(x.x, -120.0f);
    x.y = fmaxf(x.y, -120.0f);
    const float2 magic = make_float2(12582912.0f, 12582912.0f);
    const float2 t = __fadd2_rn(x, magic);
    const float2 r = __fadd2_rn(t, make_float2(-12582912.0f, -12582912.0f));
    const float2 f = __ffma2_rn(r, make_float2(-1.0f, -1.0f), x);
    float2 p = __ffma2_rn(make_float2(9.6181291e-3f, 9.6181291e-3f), f, make_float2(5.5504109e-2f, 5.5504109e-2f));
    p = __ffma2_rn(p, f, make_float2(2.4022651e-1f, 2.4022651e-1f));
    p = __ffma2_rn(p, f, make_float2(6.9314718e-1f, 6.9314718e-1f));
    p = __ffma2_rn(p, f, make_float2(1.0f, 1.0f));
    return make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23)),
                       __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23)));
}

// kPoly: 0 = every exponential on the MUFU; 4 / 8 = every 4th / 8th on the FMA pipe (scalar polynomial);
// 100 + n = n PAIRS of every 16 pairs on the FMA pipe with the packed polynomial (102: 1/8, 104: 1/4, 106: 3/8, 108: 1/2)
template <int kPoly>
__device__ __forceinline__ void exp_pack(const uint32_t (&s)[32], float neg_m, uint32_t (&out)[16]) {
    if constexpr (kPoly >= 100) {
        constexpr int kPairs = kPoly - 100;              // of 16
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
        for (int k = 0; k < 16; ++k) {
            // pair k goes to the FMA pipe when the running count k * kPairs / 16 steps up (evenly spread)
            const bool poly = ((k + 1) * kPairs) / 16 != (k * kPairs) / 16;
            if (poly) {
                const float2 y = poly_exp2x2(make_float2(e[2 * k], e[2 * k + 1]));
                e[2 * k] = y.x;
                e[2 * k + 1] = y.y;
            } else {
                e[2 * k] = fast_exp2(e[2 * k]);
                e[2 * k + 1] = fast_exp2(e[2 * k + 1]);
            }
        }
        join32(e);
#pragma unroll
        for (int i = 0; i < 32; i += 2) out[i >> 1] = pack_bf16x2(e[i], e[i + 1]);
        return;
    }
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
    for (int i = 0; i < 32; ++i) e[i] = (kPoly > 0 && (i % (kPoly > 0 ? kPoly : 1)) == (kPoly - 1)) ? poly_exp2(e[i]) : fast_exp2(e[i]);
    join32(e);
#pragma unroll
    for (int i = 0; i < 32; i += 2) out[i >> 1] = pack_bf16x2(e[i], e[i + 1]);
}

template <int kPoly, bool kTrace>
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_vt, const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
    uint8_t* sQ = smem;                                  // [tile][16 KB]
    uint8_t* sK = sQ + kQTiles * kQTileBytes;            // [stage][8 KB]
    uint8_t* sV = sK + kKStages * kKBytes;               // [stage][10 KB]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kVStages * kVBytes);
    uint64_t* q_full = bars;                             // 1
    uint64_t* k_full = bars + 1;                         // [4]
    uint64_t* v_full = bars + 5;                         // [4]
    uint64_t* k_empty = bars + 9;                        // [4]
    uint64_t* v_empty = bars + 13;                       // [4]
    uint64_t* s_full = bars + 17;                        // [tile][buffer]  S_g(j) written
    uint64_t* p_full = bars + 21;                        // [tile][buffer]  P_g(j) published by the 128 softmax threads
    uint64_t* pv_done = bars + 25;                       // [tile]  O_g += P_g(n-2) V_(n-2) retired
    uint64_t* o_full = bars + 27;                        // [tile]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 29);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * (kQTiles * kBlockQ);
    const int head = blockIdx.y;
    const int b = blockIdx.z;
    const int n_kv = (p.T + kBlockKV - 1) / kBlockKV;
    Tracer tr{nullptr, 0};
    if (kTrace) {
        const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
        if (lane == 0 && (cta % 61) == 5 && cta / 61 < kTraceCtas)
            tr.p = g_attn_trace + ((cta / 61) * kTraceWarps + warp) * kTraceEvents;
        if (kTrace) tr.stamp();
    }

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_k);
        tma_prefetch_desc(&tmap_vt);
        mbar_init(q_full, 1);
        for (int s = 0; s < kKStages; ++s) {
            mbar_init(&k_full[s], 1);
            mbar_init(&k_empty[s], kQTiles);
        }
        for (int s = 0; s < kVStages; ++s) {
            mbar_init(&v_full[s], 1);
            mbar_init(&v_empty[s], kQTiles);
        }
        for (int g = 0; g < kQTiles; ++g) {
            mbar_init(&s_full[2 * g], 1);
            mbar_init(&s_full[2 * g + 1], 1);
            mbar_init(&p_full[2 * g], kSoftmaxThreads);
            mbar_init(&p_full[2 * g + 1], kSoftmaxThreads);
            mbar_init(&pv_done[g], 1);
            mbar_init(&o_full[g], 1);
        }
        fence_mbar_init();
    }
    // Constant tail of every V stage: B-operand rows 64..79 of the O MMA.  Row 64 is all ones, so column 64 of O
    // accumulates the row sums of (the bf16-rounded) P on the tensor core; rows 65..79 pad N to a legal 80.
    for (int i = threadIdx.x; i < kVStages * 512; i += kThreads) {
        const int s = i >> 9, w = i & 511;               // 512 words per 2 KB tail
        reinterpret_cast<uint32_t*>(sV + s * kVBytes + kVTmaBytes)[w] = (w < 32) ? 0x3F803F80u : 0u;
    }
    fence_proxy_async_smem();
    if (warp == kMmaWarp) tmem_alloc<kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == kTmaWarp) {
        // ---------------------------------------------------------------- TMA producers: lane 0 = Q + K, lane 1 = V
        if (lane == 0) {
            mbar_expect_tx(q_full, kQTiles * kQTileBytes);
            for (int g = 0; g < kQTiles; ++g)
                tma_load_3d(sQ + g * kQTileBytes, &tmap_q, q_full, head * kHeadDim, q0 + g * kBlockQ, b);
            for (int j = 0; j < n_kv; ++j) {
                const int s = j % kKStages;
                if (j >= kKStages) mbar_wait_relaxed(&k_empty[s], (j / kKStages - 1) & 1);
                mbar_expect_tx(&k_full[s], kKBytes);
                tma_load_3d(sK + s * kKBytes, &tmap_k, &k_full[s], p.d_model + head * kHeadDim, j * kBlockKV, b);
            }
        } else if (lane == 1) {
            for (int j = 0; j < n_kv; ++j) {
                const int s = j % kVStages;
                if (j >= kVStages) mbar_wait_relaxed(&v_empty[s], (j / kVStages - 1) & 1);
                mbar_expect_tx(&v_full[s], kVTmaBytes);
                tma_load_3d(sV + s * kVBytes, &tmap_vt, &v_full[s], j * kBlockKV, head * kHeadDim, b);
            }
        }
    } else if (warp >= kMmaWarp) {
        {
            // ---------------------------------------------------------------- MMA issuer of query tile g
            // tcgen05.mma is issued by one thread and these MMAs are small (~40 tensor-pipe cycles each), so the
            // instruction count per MMA of the issuing warp bounds the kernel: one issuing warp per query tile, the
            // whole warp runs this loop convergently (the MMA itself is predicated on elect.sync, see ptx.cuh), and
            // the key loop is unrolled by the ring depth so that every descriptor is "base + compile-time constant".
            static_assert(kKStages == 4 && kVStages == 4, "the issue loop is unrolled by the ring depth");
            const int g = warp - kMmaWarp;
            constexpr uint32_t idesc_s = umma_idesc_bf16(kBlockQ, kBlockKV, false, false);
            constexpr uint32_t idesc_o = umma_idesc_bf16(kBlockQ, kOCols, false, false);
            constexpr uint64_t desc_hi64 = umma_smem_desc_hi(16, 1024);
            constexpr uint32_t desc_hi = (uint32_t)(desc_hi64 >> 32);
            const uint32_t dQ = (uint32_t)(desc_hi64 & 0xFFFFFFFFu) | (((base + g * kQTileBytes) >> 4) & 0x3FFF);
            const uint32_t dK = (uint32_t)(desc_hi64 & 0xFFFFFFFFu) | (((base + kQTiles * kQTileBytes) >> 4) & 0x3FFF);
            const uint32_t dV = dK + ((kKStages * kKBytes) >> 4);
            const uint32_t tS = tmem_base + g * 128;     // two score buffers of 64 columns
            const uint32_t tO = tmem_base + kTmemO + g * kOCols;
            auto issue_qk = [&](int s, int buf, uint32_t k_parity) {     // S_g = Q_g K^T into score buffer buf
                mbar_wait(&k_full[s], k_parity);
                tc_fence_after();
                static_assert(kHeadDim == 64, "one x4 group per QK tile");
                umma_bf16_ss_x4_elect(tS + buf * kBlockKV, dQ, dK + ((s * kKBytes) >> 4), desc_hi, idesc_s, 0);
                umma_commit_elect(&k_empty[s]);
                umma_commit_elect(&s_full[2 * g + buf]);
            };
            mbar_wait(q_full, 0);
            issue_qk(0, 0, 0);
            if (n_kv > 1) issue_qk(1, 1, 0);
            for (int j0 = 0; j0 < n_kv; j0 += 4) {
                const uint32_t ring_parity = (j0 >> 2) & 1;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + u;
                    if (j >= n_kv) break;
                    // O_g (+)= P_g(j) [V_j | 1]: P published (so S_g(j) is fully read and O_g rescaled if needed).
                    // One barrier per score buffer: the softmax may run two tiles ahead of this thread, which a
                    // single parity bit could not tell apart.
                    if (kTrace) tr.stamp();
                    mbar_wait(&p_full[2 * g + (u & 1)], (j >> 1) & 1);
                    if (kTrace) tr.stamp();
                    mbar_wait(&v_full[u], ring_parity);
                    tc_fence_after();
                    umma_bf16_ts_x4_elect(tO, tS + (u & 1) * kBlockKV, dV + ((u * kVBytes) >> 4), desc_hi, idesc_o,
                                          j > 0);
                    umma_commit_elect(&v_empty[u]);
                    if (j == n_kv - 2) umma_commit_elect(&pv_done[g]);
                    if (j == n_kv - 1) umma_commit_elect(&o_full[g]);
                    // S_g(j+2) overwrites P_g(j): the tensor pipe executes in issue order, so PV_g(j) has read it
                    if (kTrace) tr.stamp();
                    if (j + 2 < n_kv) issue_qk((u + 2) & 3, u & 1, u < 2 ? ring_parity : ring_parity ^ 1);
                    if (kTrace) tr.stamp();
                }
            }
        }
    } else {
        // -------------------------------------------------------------------- softmax warps: one query row per thread
        const int g = warp >> 2;                             // query tile
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;                 // query row inside the tile == TMEM lane
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const uint32_t tmem_s = tmem_base + lane_addr + g * 128;               // S_g buffers, P_g on top of them
        const uint32_t tmem_o = tmem_base + lane_addr + kTmemO + g * kOCols;
        float m_ref = 0.0f;
        // Padding trim (T = 1500 = 11.7 query tiles = 23.4 key tiles): a warp whose 32 query rows all lie past T only keeps
        // the handshake going -- its rows are never stored, so what the PV MMA reads as their P is irrelevant -- and the
        // second 32 columns of a last key tile with <= 32 real keys are published as zeros without being loaded,
        // compared or exponentiated.  Both are warp-uniform; together ~4 % of the MUFU work of large-v3.
        const bool dead_rows = q0 + g * kBlockQ + quarter * 32 >= p.T;

        // One key tile.  kLast is a compile-time tag: the steady-state instance carries no masking and no branch on the
        // tile index at all (this loop is bound by the MUFU / MIO issue order and measurably sensitive to any extra
        // control flow), the last-tile instance masks the keys >= T and drops the second half when it is all padding.
        auto tile = [&](int j, auto last_tag) {
            constexpr bool kLast = decltype(last_tag)::value;
            const int buf = j & 1;
            if (kTrace) tr.stamp();
            mbar_wait(&s_full[2 * g + buf], (j >> 1) & 1);
            tc_fence_after();
            if (kTrace) tr.stamp();
            const int kv_valid = p.T - j * kBlockKV;
            const bool half_tile = kLast && kv_valid <= 32;
            uint32_t s0[32], s1[32];
            tmem_ld_32x32b_x32(tmem_s + buf * kBlockKV, s0);
            if (!half_tile) tmem_ld_32x32b_x32(tmem_s + buf * kBlockKV + 32, s1);
            tmem_ld_wait_on(s0);
            if (!half_tile) {
                tmem_ld_wait_on(s1);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) s1[i] = 0xFF800000u;
            }
            if (kLast && kv_valid < kBlockKV) {              // keys >= T do not exist
                mask32(s0, 0, kv_valid);
                if (!half_tile) mask32(s1, 32, kv_valid);
            }
            const float m_tile = half_tile ? max32(s0) : fmaxf(max32(s0), max32(s1));
            if (j == 0) {
                m_ref = m_tile;
            } else {
                // lazy rescale: keep exponentiating against a stale maximum until it is off by more than 2^8
                const bool need = (m_tile - m_ref) * kScale > kRescaleThreshold;
                if (__any_sync(0xffffffffu, need)) {
                    // O_g += P_g(j-1) V_(j-1) must have retired: S_g(j+1) was issued after it, so its commit
                    // covers it; the last tile has no successor and uses a commit of its own
                    if (!kLast) mbar_wait(&s_full[2 * g + (buf ^ 1)], ((j + 1) >> 1) & 1);
                    else mbar_wait(&pv_done[g], 0);
                    tc_fence_after();
                    const float f = need ? fast_exp2((m_ref - m_tile) * kScale) : 1.0f;
                    if (need) m_ref = m_tile;
                    {
                        uint32_t o[32];
#pragma unroll 1
                        for (int c = 0; c < 2; ++c) {
                            tmem_ld_32x32b_x32(tmem_o + c * 32, o);
                            tmem_ld_wait_on(o);
#pragma unroll
                            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
                            tmem_st_32x32b_x32(tmem_o + c * 32, o);
                        }
                    }
                    uint32_t o16[16];                        // columns 64..79 (the row sums live in column 64)
                    tmem_ld_32x32b_x16(tmem_o + 64, o16);
                    tmem_ld_wait_on(o16);
#pragma unroll
                    for (int i = 0; i < 16; ++i) o16[i] = __float_as_uint(__uint_as_float(o16[i]) * f);
                    tmem_st_32x32b_x16(tmem_o + 64, o16);
                }
            }
            const float neg_m = -m_ref * kScale;
            uint32_t pk0[16], pk1[16];
            if (kTrace) tr.stamp();
            exp_pack<kPoly>(s0, neg_m, pk0);
            if (!half_tile) {
                exp_pack<kPoly>(s1, neg_m, pk1);
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) pk1[i] = 0u;    // exp2(-inf) = 0, as the masked path would have produced
            }
            if (kTrace) tr.stamp();
            tmem_st_32x32b_x16(tmem_s + buf * kBlockKV, pk0);
            tmem_st_32x32b_x16(tmem_s + buf * kBlockKV + 16, pk1);
            tmem_st_wait();
            tc_fence_before();              // our TMEM reads / writes are complete and ordered before the arrive
            mbar_arrive(&p_full[2 * g + buf]);
        };
        if (dead_rows) {
            // all 32 query rows of this warp lie past T: keep the handshake going, compute nothing
#pragma unroll 1
            for (int j = 0; j < n_kv; ++j) {
                mbar_wait(&s_full[2 * g + (j & 1)], (j >> 1) & 1);
                tc_fence_after();
                tc_fence_before();
                mbar_arrive(&p_full[2 * g + (j & 1)]);
            }
        } else {
#pragma unroll 1
            for (int j = 0; j + 1 < n_kv; ++j) tile(j, std::false_type{});
            tile(n_kv - 1, std::true_type{});
        }

        if (kTrace) tr.stamp();
        mbar_wait(&o_full[g], 0);
        tc_fence_after();
        if (kTrace) tr.stamp();
        const int t = q0 + g * kBlockQ + row;
        uint32_t osum[16];
        tmem_ld_32x32b_x16(tmem_o + 64, osum);                   // column 64 = row sum of P
        tmem_ld_wait_on(osum);
        const float inv = 1.0f / __uint_as_float(osum[0]);
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + ((long long)b * p.T + t) * p.d_model +
                             head * kHeadDim;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(tmem_o + c * 32, o);
            tmem_ld_wait_on(o);
            if (t < p.T) {
                uint4* d4 = reinterpret_cast<uint4*>(dst + c * 32);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint4 q;
                    q.x = pack_bf16x2(__uint_as_float(o[8 * i + 0]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
                    q.y = pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
                    q.z = pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
                    q.w = pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
                    d4[i] = q;
                }
            }
        }
        tc_fence_before();
    }

    if (kTrace) tr.stamp();
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc<kTmemCols>(tmem_base);
    }
    if (kTrace) tr.stamp();
}

int attn_variant() {               // 0: all exponentials on the MUFU; 8 / 4: every 8th / 4th on the FMA pipe; -1: trace
    static const int v = [] {
        const char* e = getenv("ARIES_ATTN_POLY");
        const char* r = getenv("ARIES_ATTN_TRACE");
        if (r && r[0] != '0') return -1;
        const int n = e ? atoi(e) : kDefaultPoly;
        return (n == 4 || n == 8 || n == 102 || n == 104 || n == 106 || n == 108) ? n : 0;
    }();
    return v;
}

template <int kPoly, bool kTrace>
cudaError_t set_smem() {
    return cudaFuncSetAttribute(attention_fwd_kernel<kPoly, kTrace>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                kSmemBytes);
}

}  // namespace

cudaError_t attention_init_device() {
    cudaError_t e;
    if ((e = set_smem<0, false>()) != cudaSuccess) return e;
    if ((e = set_smem<8, false>()) != cudaSuccess) return e;
    if ((e = set_smem<4, false>()) != cudaSuccess) return e;
    if ((e = set_smem<102, false>()) != cudaSuccess) return e;
    if ((e = set_smem<104, false>()) != cudaSuccess) return e;
    if ((e = set_smem<106, false>()) != cudaSuccess) return e;
    if ((e = set_smem<108, false>()) != cudaSuccess) return e;
    return set_smem<0, true>();
}

cudaError_t attention_read_trace(unsigned long long* host, size_t count) {
    if (count > sizeof(g_attn_trace) / sizeof(unsigned long long)) count = sizeof(g_attn_trace) / sizeof(unsigned long long);
    return cudaMemcpyFromSymbol(host, g_attn_trace, count * sizeof(unsigned long long));
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
    dim3 grid((p.T + kQTiles * kBlockQ - 1) / (kQTiles * kBlockQ), p.n_heads, p.batch);
    switch (attn_variant()) {
        case 8: attention_fwd_kernel<8, false><<<grid, kThreads, kSmemBytes, stream>>>(maps.q, maps.k, maps.vt, p); break;
        case 4: attention_fwd_kernel<4, false><<<grid, kThreads, kSmemBytes, stream>>>(maps.q, maps.k, maps.vt, p); break;
        case 102: attention_fwd_kernel<102, false><<<grid, kThreads, kSmemBytes, stream>>>(maps.q, maps.k, maps.vt, p); break;
        case 104: attention_fwd_kernel<104, false><<<grid, kThreads, kSmemBytes, stream>>>(maps.q, maps.k, maps.vt, p); break;
        case 106: attention_fwd_kernel<106, false><<<grid, kThreads, kSmemBytes, stream>>>(maps.q, maps.k, maps.vt, p); break;
        case 108: attention_fwd_kernel<108, false><<<grid, kThreads, kSmemBytes, stream>>>(maps.q, maps.k, maps.vt, p); break;
        case -1: attention_fwd_kernel<0, true><<<grid, kThreads, kSmemBytes, stream>>>(maps.q, maps.k, maps.vt, p); break;
        default: attention_fwd_kernel<0, false><<<grid, kThreads, kSmemBytes, stream>>>(maps.q, maps.k, maps.vt, p); break;
    }
    return cudaGetLastError();
}

}  // namespace aries
