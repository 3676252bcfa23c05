// Weight-streaming GEMM of the decode step (row f1) for sm_100a: out[b, n] = epilogue(sum_k X[b, k] W[n, k]).
//
// A decode step multiplies B <= 128 token rows by every Linear weight of the Whisper decoder (CT2 layers::Dense inside
// layers::WhisperDecoder, reached from ctranslate2.models.Whisper.generate).  The op is bound by
// reading W from HBM exactly once, so the roles of the tcgen05 operands are swapped with respect to the encoder GEMM:
//   * W is the M operand: one CTA owns 128 weight rows (one TMA box of 64 x 128 per 64-deep block, SWIZZLE_128B),
//   * the token rows are the N operand (NB = batch padded to a multiple of 16, one TMA box of 64 x NB),
//   * the accumulator is 128 TMEM lanes (output features) x NB columns (sequences).
// K is split over a thread-block CLUSTER (grid.y = cluster size = splits <= 8) so that 70..300 CTAs stream weights at
// once; the CTAs of a cluster park their f32 partial tile in their own shared memory and every CTA then reduces and
// finishes 1/splits of the sequences by reading its peers' tiles through distributed shared memory -- fixed summation
// order (bit-reproducible), no HBM round trip, no atomics.
// Programmatic dependent launch: the producer prefetches its first WEIGHT blocks before griddepcontrol.wait (weights
// are never written during decoding), so weight streaming overlaps the tail of the previous kernel of the step.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdlib>

#include "ptx.cuh"
#include "skinny.h"

namespace aries {

namespace {

constexpr int BMW = 128;            // weight rows per CTA
constexpr int BK = 64;
constexpr int kStages = 4;
constexpr int kStagesLn = 8;        // fused-LayerNorm variant: all K-blocks of a split are resident (no ring reuse)
constexpr int kLnMaxB = 8;          // sequences the fused variant handles (two rows per epilogue warp)
constexpr int kLnMaxJ = 5;          // 16-byte chunks of a row per lane: K <= 5 * 256
constexpr int kWBytes = BMW * BK * 2;
constexpr int kThreads = 192;       // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
constexpr uint32_t kTmemCols = 128;
constexpr int kMaxSplits = 8;

__host__ __device__ inline int stage_bytes(int NB) { return kWBytes + NB * BK * 2; }
// MODE 0: token rows by TMA, 4-stage ring (two CTAs per SM: grids beyond one wave, i.e. the logits projection)
// MODE 1: fused LayerNorm, 8 resident stages    MODE 2: token rows by TMA, 8 stages (one CTA per SM: with programmatic
// dependent launch up to 8 weight blocks per CTA are in flight before the previous kernel has finished)
constexpr int kModeRing4 = 0, kModeLn = 1, kModeRing8 = 2;
inline int smem_bytes(int NB, int mode = kModeRing4) {
    const int pipe = (mode == kModeRing4 ? kStages : kStagesLn) * stage_bytes(NB);
    const int part = NB * BMW * 4;
    return (pipe > part ? pipe : part) + 256 + 1024;
}

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ uint32_t dsmem_addr(uint32_t local_addr, uint32_t rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
    return remote;
}
__device__ __forceinline__ float4 ld_dsmem_f32x4(uint32_t remote) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(remote));
    return v;
}

// The residual value is passed in: the callers fetch the residuals of ALL the elements they own before the first store
// (a load inside the store loop may alias the previous store, so the compiler keeps them in order and every element pays
// a full memory round trip: 10 us for 16 sequences).
template <int EPI>
__device__ __forceinline__ float load_resid(const SkinnyParams& p, int b, int n) {
    if (EPI != SK_BIAS_RESID_F16) return 0.0f;
    return __half2float(reinterpret_cast<const __half*>(p.out)[(size_t)b * p.ldo + n]);
}

constexpr bool is_lnf(int epi) { return epi == SK_LNF_BF16 || epi == SK_LNF_GELU_BF16; }
constexpr bool is_gelu(int epi) { return epi == SK_BIAS_GELU_BF16 || epi == SK_LNF_GELU_BF16; }

// Row statistics of the folded LayerNorm: (-mean * rstd, rstd) from the producer's partial sums.
struct RowStat {
    float nm, rstd;
};
template <int EPI>
__device__ __forceinline__ RowStat row_stat(const SkinnyParams& p, int b) {
    RowStat r{0.0f, 1.0f};
    if (!is_lnf(EPI)) return r;
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll 4
    for (int k = 0; k < p.stats_parts; ++k) {
        const float2 v = __ldcg(p.stats_in + (size_t)b * p.stats_parts + k);
        s1 += v.x;
        s2 += v.y;
    }
    const float inv_d = 1.0f / (float)p.ln_dim;
    const float mean = s1 * inv_d;
    r.rstd = rsqrtf(fmaxf(s2 * inv_d - mean * mean, 0.0f) + 1e-5f);
    r.nm = -mean * r.rstd;
    return r;
}

// Returns the value as stored (after rounding) for the residual epilogue's statistics, else the input.
template <int EPI>
__device__ __forceinline__ float store_one(const SkinnyParams& p, int b, int n, float v, float resid, RowStat rs) {
    const size_t at = (size_t)b * p.ldo + n;
    if (EPI == SK_LOGITS_F32) {
        reinterpret_cast<float*>(p.out)[at] = v;
        return v;
    }
    if (is_lnf(EPI)) v = fmaf(rs.rstd, v, fmaf(rs.nm, __ldg(p.c1 + n), __ldg(p.bias + n)));
    else v += __ldg(p.bias + n);
    if (is_gelu(EPI)) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
    if (EPI == SK_BIAS_RESID_F16) {
        __half* o = reinterpret_cast<__half*>(p.out);
        v += resid;
        const __half h = __float2half_rn(fminf(fmaxf(v, -65504.0f), 65504.0f));
        o[at] = h;
        return __half2float(h);
    } else {
        reinterpret_cast<__nv_bfloat16*>(p.out)[at] = __float2bfloat16_rn(v);
    }
    return v;
}

// Four consecutive output features of one sequence (n % 4 == 0): vector loads / stores when the row pitch allows it.
template <int EPI>
__device__ __forceinline__ float4 load_resid4(const SkinnyParams& p, int b, int n) {
    if (EPI != SK_BIAS_RESID_F16) return make_float4(0, 0, 0, 0);
    if (n + 3 < p.N && (p.ldo & 3) == 0) {
        const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(p.out) + (size_t)b * p.ldo + n);
        const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
        const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
        return make_float4(lo.x, lo.y, hi.x, hi.y);
    }
    float r[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) r[e] = (n + e < p.N) ? load_resid<EPI>(p, b, n + e) : 0.0f;
    return make_float4(r[0], r[1], r[2], r[3]);
}

// (returns the four values as stored, see store_one)
template <int EPI>
__device__ __forceinline__ float4 store_four(const SkinnyParams& p, int b, int n, float4 v, float4 resid, RowStat rs) {
    if (n + 3 < p.N && (p.ldo & 3) == 0) {
        const size_t at = (size_t)b * p.ldo + n;
        if (EPI == SK_LOGITS_F32) {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + at) = v;
            return v;
        }
        const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + n));
        if (is_lnf(EPI)) {
            const float4 cc = __ldg(reinterpret_cast<const float4*>(p.c1 + n));
            v.x = fmaf(rs.rstd, v.x, fmaf(rs.nm, cc.x, bb.x));
            v.y = fmaf(rs.rstd, v.y, fmaf(rs.nm, cc.y, bb.y));
            v.z = fmaf(rs.rstd, v.z, fmaf(rs.nm, cc.z, bb.z));
            v.w = fmaf(rs.rstd, v.w, fmaf(rs.nm, cc.w, bb.w));
        } else {
            v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
        }
        if (is_gelu(EPI)) {
            v.x = 0.5f * v.x * (1.0f + erff(v.x * 0.70710678118654752f));
            v.y = 0.5f * v.y * (1.0f + erff(v.y * 0.70710678118654752f));
            v.z = 0.5f * v.z * (1.0f + erff(v.z * 0.70710678118654752f));
            v.w = 0.5f * v.w * (1.0f + erff(v.w * 0.70710678118654752f));
        }
        uint2 q;
        if (EPI == SK_BIAS_RESID_F16) {
            const float kMax = 65504.0f;
            const __half2 lo = __floats2half2_rn(fminf(fmaxf(v.x + resid.x, -kMax), kMax), fminf(fmaxf(v.y + resid.y, -kMax), kMax));
            const __half2 hi = __floats2half2_rn(fminf(fmaxf(v.z + resid.z, -kMax), kMax), fminf(fmaxf(v.w + resid.w, -kMax), kMax));
            q.x = *reinterpret_cast<const unsigned*>(&lo);
            q.y = *reinterpret_cast<const unsigned*>(&hi);
            *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p.out) + at) = q;
            const float2 flo = __half22float2(lo), fhi = __half22float2(hi);
            return make_float4(flo.x, flo.y, fhi.x, fhi.y);
        } else {
            q.x = pack_bf16x2(v.x, v.y);
            q.y = pack_bf16x2(v.z, v.w);
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + at) = q;
        }
        return v;
    }
    const float vv[4] = {v.x, v.y, v.z, v.w}, rr[4] = {resid.x, resid.y, resid.z, resid.w};
    float st[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int e = 0; e < 4; ++e)
        if (n + e < p.N) st[e] = store_one<EPI>(p, b, n + e, vv[e], rr[e], rs);
    return make_float4(st[0], st[1], st[2], st[3]);
}

template <int EPI, int MODE>
__global__ void __launch_bounds__(kThreads)
skinny_gemm_tcgen05(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
                    const SkinnyParams p) {
    constexpr bool LN = MODE == kModeLn;
    constexpr int kStages = (MODE == kModeRing4) ? aries::kStages : aries::kStagesLn;      // shadows the namespace constant
    extern __shared__ uint8_t smem_raw[];
    __shared__ float s_gamma[LN ? kStagesLn * BK : 1], s_beta[LN ? kStagesLn * BK : 1];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
    const int sbytes = stage_bytes(p.NB);
    const int pipe = kStages * sbytes, part = p.NB * BMW * 4;
    uint8_t* tail = smem + (pipe > part ? pipe : part);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tfull_bar = empty_bar + kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BMW;
    const int split = blockIdx.y;               // == rank in the cluster (cluster = 1 x splits x 1)
    const int num_kb = p.K / BK;
    const int kps = (num_kb + p.splits - 1) / p.splits;
    const int kb0 = split * kps;
    const int kb1 = (kb0 + kps < num_kb) ? kb0 + kps : num_kb;
    const int my_kb = kb1 - kb0;                // >= 1 by construction of `splits`

    const int pre = my_kb < kStages ? my_kb : kStages;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_x);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], LN ? 2 : 1);      // LN: the weight TMA and the epilogue warps' operand tile
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tfull_bar, 1);
        fence_mbar_init();
        // The producer starts streaming WEIGHTS at once: before the TMEM allocation, the CTA-wide sync and
        // griddepcontrol.wait (weights are never written during decoding, so they do not depend on the previous
        // kernel) -- the first HBM round trip overlaps the whole prologue and the tail of the previous kernel.
        for (int i = 0; i < pre; ++i) {
            mbar_expect_tx(&full_bar[i], LN ? kWBytes : sbytes);
            tma_load_2d(smem + i * sbytes, &tmap_w, &full_bar[i], (kb0 + i) * BK, n0);
        }
    }
    if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ---------------------------------------------------------------- TMA producer (continued)
            pdl_wait();
            pdl_trigger();
            if (!LN) {
                for (int i = 0; i < pre; ++i)
                    tma_load_2d(smem + i * sbytes + kWBytes, &tmap_x, &full_bar[i], (kb0 + i) * BK, 0);
            }
            int stage = pre % kStages;
            uint32_t phase = (pre == kStages) ? 1 : 0;          // the ring wrapped once
            for (int i = pre; i < my_kb; ++i) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* s = smem + stage * sbytes;
                mbar_expect_tx(&full_bar[stage], sbytes);
                tma_load_2d(s, &tmap_w, &full_bar[stage], (kb0 + i) * BK, n0);
                tma_load_2d(s + kWBytes, &tmap_x, &full_bar[stage], (kb0 + i) * BK, 0);
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        } else {
            pdl_wait();
            pdl_trigger();
        }
    } else if (warp == 1) {
        // -------------------------------------------------------------------- MMA issuer (convergent, elect-predicated)
        pdl_wait();
        pdl_trigger();
        const uint32_t idesc = is_lnf(EPI) ? umma_idesc_f16(BMW, (uint32_t)p.NB) : umma_idesc_bf16(BMW, (uint32_t)p.NB, false, false);
        constexpr uint64_t desc_hi64 = umma_smem_desc_hi(16, 1024);
        constexpr uint32_t desc_hi = (uint32_t)(desc_hi64 >> 32);
        const uint32_t desc_lo0 = (uint32_t)(desc_hi64 & 0xFFFFFFFFu) | ((base >> 4) & 0x3FFF);
        int stage = 0;
        uint32_t phase = 0;
        for (int i = 0; i < my_kb; ++i) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t a_lo = desc_lo0 + ((uint32_t)(stage * sbytes) >> 4);
            umma_bf16_ss_x4_elect(tmem_base, a_lo, a_lo + (kWBytes >> 4), desc_hi, idesc, i != 0);
            umma_commit_elect(&empty_bar[stage]);
            if (i == my_kb - 1) umma_commit_elect(tfull_bar);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
    } else {
        // -------------------------------------------------------------------- epilogue warps: one output feature per thread
        if (LN) {
            // ---- fused LayerNorm: these 4 warps BUILD the token-row operand (they are idle until the MMAs retire).
            // Before griddepcontrol.wait (nothing here depends on the previous kernel): gamma / beta of this split's
            // K range to shared memory, and zeros into the pad rows of every operand tile.
            const int et = threadIdx.x - 64;                       // 0 .. 127
            const int k_lo = kb0 * BK, k_n = my_kb * BK;
            for (int i = et; i < k_n; i += 128) {
                s_gamma[i] = __ldg(p.ln_gamma + k_lo + i);
                s_beta[i] = __ldg(p.ln_beta + k_lo + i);
            }
            for (int i = et; i < (p.NB - p.B) * 8 * my_kb; i += 128) {
                const int st = i / ((p.NB - p.B) * 8), rem = i - st * (p.NB - p.B) * 8;
                const int r = p.B + rem / 8, cc = rem & 7;
                *reinterpret_cast<uint4*>(smem + st * sbytes + kWBytes + r * 128 + ((cc ^ (r & 7)) << 4)) = make_uint4(0, 0, 0, 0);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");         // gamma / beta visible to the four warps
            pdl_wait();
            pdl_trigger();
            // One warp per row (rows ew and ew + 4): the row lives in registers as 16-byte chunks (lane + 32 j), the
            // statistics are two-pass f32 like layernorm_kernel, and every chunk that falls into this split's K range is
            // normalised, rounded to bf16 and stored at its SWIZZLE_128B position (chunk ^ (row & 7)) of the K-major tile.
            const int ew = warp - 2;
            const int n_chunks = p.K / 8;
            const __half* xs = reinterpret_cast<const __half*>(p.ln_x);
            float v[2][kLnMaxJ][8];
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int r = ew + 4 * rr;
#pragma unroll
                for (int j = 0; j < kLnMaxJ; ++j) {
                    const int c = lane + 32 * j;
                    uint4 u = make_uint4(0, 0, 0, 0);
                    if (r < p.B && c < n_chunks) u = *reinterpret_cast<const uint4*>(xs + (size_t)r * p.K + 8 * c);
                    const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 f = __half22float2(h2[e]);
                        v[rr][j][2 * e] = f.x;
                        v[rr][j][2 * e + 1] = f.y;
                    }
                }
            }
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int r = ew + 4 * rr;
                if (r >= p.B) continue;                            // warp-uniform
                float sum = 0.0f;
#pragma unroll
                for (int j = 0; j < kLnMaxJ; ++j)
#pragma unroll
                    for (int e = 0; e < 8; ++e) sum += v[rr][j][e];        // chunks past the row are zero
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                const float mean = sum / (float)p.K;
                float sq = 0.0f;
#pragma unroll
                for (int j = 0; j < kLnMaxJ; ++j) {
                    if (lane + 32 * j < n_chunks) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) sq += (v[rr][j][e] - mean) * (v[rr][j][e] - mean);
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
                const float rstd = rsqrtf(sq / (float)p.K + 1e-5f);
#pragma unroll
                for (int j = 0; j < kLnMaxJ; ++j) {
                    const int c = lane + 32 * j;
                    const int kb = c >> 3;
                    if (c < n_chunks && kb >= kb0 && kb < kb1) {
                        const int kk = (kb - kb0) * BK + (c & 7) * 8;          // offset inside this split's K range
                        uint4 o4;
                        uint32_t* o = reinterpret_cast<uint32_t*>(&o4);
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            o[e] = pack_bf16x2((v[rr][j][2 * e] - mean) * rstd * s_gamma[kk + 2 * e] + s_beta[kk + 2 * e],
                                               (v[rr][j][2 * e + 1] - mean) * rstd * s_gamma[kk + 2 * e + 1] + s_beta[kk + 2 * e + 1]);
                        *reinterpret_cast<uint4*>(smem + (kb - kb0) * sbytes + kWBytes + r * 128 + (((c & 7) ^ (r & 7)) << 4)) = o4;
                    }
                }
            }
            fence_proxy_async_smem();                              // generic-proxy stores -> visible to tcgen05.mma
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (et == 0)
                for (int i = 0; i < my_kb; ++i) mbar_arrive(&full_bar[i]);
        } else {
            pdl_wait();
            pdl_trigger();
        }
        const int quarter = warp & 3;
        const int nl = quarter * 32 + lane;
        const int n = n0 + nl;
        float* my_part = reinterpret_cast<float*>(smem);
        mbar_wait(tfull_bar, 0);
        tc_fence_after();
        for (int c = 0; c < p.NB / 16; ++c) {
            uint32_t acc[16];
            tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(quarter * 32) << 16) + c * 16, acc);
            tmem_ld_wait_on(acc);
            if (p.splits == 1) {
                float res[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int b = c * 16 + j;
                    res[j] = (b < p.B && n < p.N) ? load_resid<EPI>(p, b, n) : 0.0f;
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int b = c * 16 + j;
                    if (b < p.B && n < p.N) store_one<EPI>(p, b, n, __uint_as_float(acc[j]), res[j], row_stat<EPI>(p, b));
                }
            } else {
                // the pipeline stages are free: every MMA of this CTA has retired (tfull)
#pragma unroll
                for (int j = 0; j < 16; ++j) my_part[(c * 16 + j) * BMW + nl] = __uint_as_float(acc[j]);
            }
        }
        tc_fence_before();
    }

    __syncwarp();
    if (p.splits > 1) {
        cluster_sync_all();                         // every CTA's partial tile is in its shared memory
        if (warp >= 2) {
            // Peers' tiles through distributed shared memory, 16 bytes per load: thread (n4, bsel) owns four consecutive
            // output features of every fourth sequence of this CTA's share (b = split + splits * i), and issues all the
            // loads of two sequences -- up to 16 -- before the first add (a chain of dependent scalar remote loads cost
            // 13 us for 64 sequences).  Summation order over the splits is fixed: bit-reproducible.
            const int et = threadIdx.x - 64;
            const int n4 = et & 31, bsel = et >> 5;
            const int n = n0 + 4 * n4;
            uint32_t peer[kMaxSplits];
#pragma unroll
            for (int s = 0; s < kMaxSplits; ++s) peer[s] = dsmem_addr(base + (uint32_t)n4 * 16u, (uint32_t)(s < p.splits ? s : 0));
            for (int i = bsel; split + p.splits * i < p.B; i += 8) {
                const int b = split + p.splits * i;
                const int b2 = split + p.splits * (i + 4);
                const bool has2 = b2 < p.B;
                float4 r1 = make_float4(0, 0, 0, 0), r2 = r1;
                if (n < p.N) {
                    r1 = load_resid4<EPI>(p, b, n);
                    if (has2) r2 = load_resid4<EPI>(p, b2, n);
                }
                const RowStat rs1 = row_stat<EPI>(p, b), rs2 = has2 ? row_stat<EPI>(p, b2) : RowStat{0.0f, 1.0f};
                float4 v[kMaxSplits], w[kMaxSplits];
#pragma unroll
                for (int s = 0; s < kMaxSplits; ++s) {
                    v[s] = (s < p.splits) ? ld_dsmem_f32x4(peer[s] + (uint32_t)(b * BMW) * 4u) : make_float4(0, 0, 0, 0);
                    w[s] = (s < p.splits && has2) ? ld_dsmem_f32x4(peer[s] + (uint32_t)(b2 * BMW) * 4u) : make_float4(0, 0, 0, 0);
                }
                float4 sum = make_float4(0, 0, 0, 0), sum2 = sum;
#pragma unroll
                for (int s = 0; s < kMaxSplits; ++s) {
                    sum.x += v[s].x; sum.y += v[s].y; sum.z += v[s].z; sum.w += v[s].w;
                    sum2.x += w[s].x; sum2.y += w[s].y; sum2.z += w[s].z; sum2.w += w[s].w;
                }
                float4 st1 = make_float4(0, 0, 0, 0), st2 = st1;
                if (n < p.N) {
                    st1 = store_four<EPI>(p, b, n, sum, r1, rs1);
                    if (has2) st2 = store_four<EPI>(p, b2, n, sum2, r2, rs2);
                }
                if (EPI == SK_BIAS_RESID_F16 && p.stats_out != nullptr) {
                    // folded LayerNorm, producing side: (sum, sum of squares) of this tile's 128 stored values of rows
                    // b and b2 -- the 32 lanes of the warp hold 4 features each (b, b2 and has2 are warp-uniform)
                    float a1 = st1.x + st1.y + st1.z + st1.w, q1 = st1.x * st1.x + st1.y * st1.y + st1.z * st1.z + st1.w * st1.w;
                    float a2 = st2.x + st2.y + st2.z + st2.w, q2 = st2.x * st2.x + st2.y * st2.y + st2.z * st2.z + st2.w * st2.w;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
                        q1 += __shfl_xor_sync(0xffffffffu, q1, o);
                        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
                        q2 += __shfl_xor_sync(0xffffffffu, q2, o);
                    }
                    if (n4 == 0) {
                        p.stats_out[(size_t)b * gridDim.x + blockIdx.x] = make_float2(a1, q1);
                        if (has2) p.stats_out[(size_t)b2 * gridDim.x + blockIdx.x] = make_float2(a2, q2);
                    }
                }
            }
        }
        cluster_sync_all();                         // peers are done reading this CTA's shared memory
    } else {
        __syncthreads();
    }
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<kTmemCols>(tmem_base);
    }
}

template <int EPI>
cudaError_t launch_epi(const CUtensorMap& tw, const CUtensorMap& tx, const SkinnyParams& p, cudaStream_t stream) {
    const int n_tiles = (p.N + BMW - 1) / BMW;
    // deep ring whenever the grid fits one CTA per SM anyway (every decode GEMM but the logits projection)
    // measured per token step with / without the deep ring: 1 window 1.567 / 1.602 ms, 8: 2.019 / 2.040, 64: 4.397 / 4.398
    static const bool ring8 = [] {                   // ARIES_SKINNY_RING8=0 switches it off (A/B runs)
        const char* e = getenv("ARIES_SKINNY_RING8");
        return !(e && e[0] == '0');
    }();
    const int mode = p.ln_x != nullptr ? kModeLn
                     : (ring8 && p.NB <= 64 && n_tiles * p.splits <= 160 ? kModeRing8 : kModeRing4);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_tiles, p.splits);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem_bytes(p.NB, mode);
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (p.splits > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 1;
        attr[na].val.clusterDim.y = p.splits;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (p.pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    if (mode == kModeLn) return cudaLaunchKernelEx(&cfg, skinny_gemm_tcgen05<EPI, kModeLn>, tw, tx, p);
    if (mode == kModeRing8) return cudaLaunchKernelEx(&cfg, skinny_gemm_tcgen05<EPI, kModeRing8>, tw, tx, p);
    return cudaLaunchKernelEx(&cfg, skinny_gemm_tcgen05<EPI, kModeRing4>, tw, tx, p);
}

template <int EPI>
cudaError_t set_smem() {
    cudaError_t e = cudaFuncSetAttribute(skinny_gemm_tcgen05<EPI, kModeRing4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         smem_bytes(128));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(skinny_gemm_tcgen05<EPI, kModeRing8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             smem_bytes(64, kModeRing8));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(skinny_gemm_tcgen05<EPI, kModeLn>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                smem_bytes(16, kModeLn));
}

}  // namespace

int skinny_pick_splits(int N, int K, int sm_count) {
    // About one CTA per SM: clusters are co-scheduled inside a GPC, so a grid much beyond one wave (measured: 280 CTAs in
    // clusters of 7 = 1.33 waves) pays a second round of launch + HBM latency, while 4 stages of 16 KB per CTA already
    // keep ~60 GB/s per SM in flight.
    const int n_tiles = (N + BMW - 1) / BMW;
    const int num_kb = K / BK;
    int want = sm_count / n_tiles;
    if (want > kMaxSplits) want = kMaxSplits;
    if (want > num_kb) want = num_kb;
    if (want < 1) want = 1;
    const int kps = (num_kb + want - 1) / want;
    return (num_kb + kps - 1) / kps;                             // every split owns >= 1 block
}

int skinny_pick_splits_ln(int N, int K, int sm_count) {
    const int num_kb = K / BK;
    int splits = skinny_pick_splits(N, K, sm_count);
    while (splits <= kMaxSplits && (num_kb + splits - 1) / splits > kStagesLn) ++splits;
    if (splits > kMaxSplits || splits > num_kb) return 0;
    const int kps = (num_kb + splits - 1) / splits;
    return (num_kb + kps - 1) / kps;
}

cudaError_t skinny_init_device() {
    cudaError_t e;
    if ((e = set_smem<SK_BIAS_BF16>()) != cudaSuccess) return e;
    if ((e = set_smem<SK_BIAS_GELU_BF16>()) != cudaSuccess) return e;
    if ((e = set_smem<SK_BIAS_RESID_F16>()) != cudaSuccess) return e;
    if ((e = set_smem<SK_LNF_BF16>()) != cudaSuccess) return e;
    if ((e = set_smem<SK_LNF_GELU_BF16>()) != cudaSuccess) return e;
    return set_smem<SK_LOGITS_F32>();
}

cudaError_t skinny_launch(int epi, const CUtensorMap& tw, const CUtensorMap& tx, const SkinnyParams& p,
                          cudaStream_t stream) {
    if (p.K % BK != 0 || p.K <= 0 || p.N <= 0 || p.B <= 0 || p.B > p.NB || p.NB % 16 != 0 || p.NB < 16 || p.NB > 128 ||
        p.splits < 1 || p.splits > kMaxSplits)
        return cudaErrorInvalidValue;
    const int num_kb = p.K / BK, kps = (num_kb + p.splits - 1) / p.splits;
    if ((p.splits - 1) * kps >= num_kb) return cudaErrorInvalidValue;      // an empty split would hang its cluster
    if (p.ln_x && (p.B > kLnMaxB || p.NB != 16 || p.K > kLnMaxJ * 256 || kps > kStagesLn || !p.ln_gamma || !p.ln_beta ||
                   epi == SK_BIAS_RESID_F16))
        return cudaErrorInvalidValue;
    const bool lnf = epi == SK_LNF_BF16 || epi == SK_LNF_GELU_BF16;
    if (lnf && (p.ln_x || !p.c1 || !p.stats_in || p.stats_parts < 1 || p.ln_dim < 1 || !p.bias)) return cudaErrorInvalidValue;
    if (p.stats_out && (epi != SK_BIAS_RESID_F16 || p.splits < 2)) return cudaErrorInvalidValue;
    switch (epi) {
        case SK_LNF_BF16: return launch_epi<SK_LNF_BF16>(tw, tx, p, stream);
        case SK_LNF_GELU_BF16: return launch_epi<SK_LNF_GELU_BF16>(tw, tx, p, stream);
        case SK_BIAS_BF16: return launch_epi<SK_BIAS_BF16>(tw, tx, p, stream);
        case SK_BIAS_GELU_BF16: return launch_epi<SK_BIAS_GELU_BF16>(tw, tx, p, stream);
        case SK_BIAS_RESID_F16: return launch_epi<SK_BIAS_RESID_F16>(tw, tx, p, stream);
        case SK_LOGITS_F32: return launch_epi<SK_LOGITS_F32>(tw, tx, p, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace aries
