// The whole decoder stack of one decode step as ONE persistent kernel (row f1, <= 8 sequences).  OPT-IN
// (ARIES_DECODE_STACK=1): correct, 3 launches per token instead of 260, but measured SLOWER than the launch-per-op step on
// B200 -- see "Measured" below.
//
// Replaces, per generated token, every layers::TransformerDecoderLayer of CT2's layers::WhisperDecoder as driven by
// ctranslate2.models.Whisper.generate (reference call site final_optimized_transcriber.py:326 with
// beam_size=1, temperature=0 at :432-441) [upstream, unverified offline].  The launch-per-op step (decoder.cu run_step)
// spends 260 kernels x ~6 us on a token whose weights stream from HBM in 0.25 ms: with <= 8 sequences the step is bound
// by launch + first-byte latency, eight times per layer.  Here one CTA per SM stays resident for the whole stack:
//
//   * weights: a producer warp walks the token's weight stream (QKV, O, Q2, O2, fc1, fc2 of every layer; the CTA owns
//     rows [cta N / G, (cta + 1) N / G) of each matrix) and keeps a ring of 16-row x 256-deep stages full with per-lane
//     16-byte cp.async copies (one instruction per 512-byte row piece, cp.async.mbarrier.arrive for completion).
//     Weights never depend on activations, so the ring runs AHEAD through grid barriers and attention phases;
//   * math: eight consumer warps.  The <= 8 sequences are the N = 8 dimension of mma.sync.m16n8k16 (bf16 x bf16 -> f32):
//     a 16-row weight tile is the A operand straight from the stage, the activations [8, K] (bf16, shared memory) the B
//     operand; each warp owns a 32-deep slice of every 256-deep stage, the eight partial tiles are summed in a fixed
//     order through shared memory (bit-reproducible) and the epilogue (bias, GELU / f16 residual / cache append) writes
//     the CTA's rows.  Tensor-core peak is irrelevant here (M = 16 per instruction) -- the op is the weight stream;
//   * attention: (sequence, head, key split) items over the CTAs, one pass with an online softmax, eight lanes per
//     128-byte key / value row; the splits of a (sequence, head) leave their partial (max, sum, weighted values) in a
//     small global buffer and the LAST one to arrive (a counter per pair) merges them in split order -- deterministic
//     -- into the bf16 context row.  The first 256 rows of the CTA's key / value share are requested between the
//     arrival at and the release of the grid barrier that precedes the phase (cache rows of earlier steps and encoder
//     keys do not depend on this step; requested BEFORE the arrival they delayed it: the release orders prior loads);
//   * small vectors: LayerNorm gamma / beta and the CTA's bias slice of every phase ride the same ring as extra stages
//     (they are evicted from L2 by the 1.6 GB weight stream between tokens: fetched on demand each was a DRAM miss on
//     the critical path of every phase -- 5 us per LayerNorm phase in the first version);
//   * phases of a layer, each closed by a grid-wide barrier (red.release + ld.acquire on one counter):
//       LN1 + QKV (+ cache append) | self-attention | O + residual | LN2 + Q | cross-attention | O + residual |
//       LN3 + fc1 + GELU | fc2 + residual
//     The token embedding is folded into the first phase, the final LayerNorm into the last.
// Numerics are those of the launch-per-op path: bf16 weights and activations, f32 accumulation, f16 residual stream,
// two-pass f32 LayerNorm statistics, exact-erf GELU (tests/test_gpu_decoder.py: both paths against the fp32 oracle).
//
// Measured (B200, large-v3 decoder, graph replay, tests/gpu_diag_decode.py stack; per-phase clock64 stamps with
// ARIES_STACK_TRACE=<cta>, tests/stack_trace.py): 1.69 ms per token at 1 window against 1.54 for the launch-per-op step,
// 3.4 against 2.0 at 8 windows.  One layer = ~97 000 cycles (51 us at 1.92 GHz), of which
//     8 grid barriers          2 300 - 5 500 each   (arrival fence + one L2 round trip + poll round trip + skew)
//     3 LayerNorm stagings     3 600 - 7 500 each   (every CTA re-reads the residual rows other SMs have just written)
//     2 attention phases       6 000 / 9 500        (q round trip, key/value round trip, split merge through L2)
//     6 weight phases          ~500 per 8 KB stage  (the same with cp.async.bulk from 16 lanes, from one lane, and
//                                                    with the ring demonstrably full: a consumer-side cost per stage)
// i.e. every dependent global round trip of this design costs 1 - 1.3 us on B200, a layer needs ~3 per phase, and the
// launch-per-op step with programmatic dependent launch already overlaps most of its launch latency with the previous
// kernel's tail.  Tried and measured without gain: L2 prefetch of the next layer's slabs (cp.async.bulk.prefetch.L2;
// the bursts delayed the latency-critical loads), single-lane bulk copies, spin vs back-off producer waits.  What
// would be needed next is fewer GLOBAL barriers per layer (clusters per head with DSMEM for QKV -> attention and
// Q -> cross-attention: 8 -> 6 barriers) and 40 KB stages; not built this round.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "ptx.cuh"
#include "skinny.h"

namespace aries {

namespace {

constexpr int kConsumerThreads = 256;                // 8 warps
constexpr int kStackThreads = 288;                   // + the producer warp
constexpr int kTileRows = 16;
constexpr int kKBlock = 256;
constexpr int kRowPitch = kKBlock * 2 + 64;          // bytes; the 64 keeps 16-byte reads of 8 rows on distinct banks
constexpr int kStageBytes = kTileRows * kRowPitch;   // 9216
constexpr int kMaxStages = 24;
constexpr int kBarBytes = 1024;
constexpr int kRedBytes = 2 * 8 * 128 * 4;           // two buffers of [8 warps][16 rows][8 sequences] f32
constexpr int kMaxB = 8;
constexpr int kLnJ = 5;                              // 16-byte chunks of a row per lane: d <= 1280
constexpr int kSmemLimit = 232448;
constexpr int kMaxHeads = 32;                        // head_dim 64: d <= 2048 (the LayerNorm staging limits d to 1280 anyway)

struct Ctx {
    uint64_t *full, *empty;
    uint8_t *ring, *act;
    float* red;
    float* s_bias;                                   // the phase's bias slice (rows n0 & ~3 ..)
    unsigned* s_flag;                                // "this CTA merges the splits" (attention phases)
    int act_pitch;                                   // bytes between sequences in the activation operand
    int ns, stage;
    uint32_t ph;
    int red_buf;
    int G, cta, tid, warp, lane;
    unsigned bar_target;
    int ti;                                          // next trace slot (diagnostics)
};
#define STACK_TRACE(p, c)                                                                     \
    do {                                                                                      \
        if ((p).trace != nullptr && (c).cta == (p).trace_cta && (c).tid == 0) (p).trace[(c).ti++] = clock64(); \
    } while (0)

__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// 16-byte asynchronous copy (LDGSTS, L1 bypassed) and "arrive on `bar` when this thread's copies so far have landed"
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Grid-wide barrier of the consumer warps (the producer warp never waits here).  One counter, zeroed between launches
// (decode_sample_kernel / the host before the first token); barrier k completes when it reaches (k + 1) G.
// `between` runs after the arrival and before the wait: loads issued there (key / value rows of the next phase) are not
// ordered by the release, so they neither delay the arrival nor wait for the barrier.
template <typename F>
__device__ __forceinline__ void grid_sync(Ctx& c, unsigned* bar, F&& between) {
    c.bar_target += (unsigned)c.G;
    consumer_sync();
    if (c.tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    between();
    if (c.tid == 0) {
        const long long t0 = clock64();
        unsigned v;
        while (true) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
            if ((int)(v - c.bar_target) >= 0) break;
            if (clock64() - t0 > 6000000000LL) __trap();     // seconds: a protocol bug must fail, not hang the GPU
        }
    }
    consumer_sync();
}
__device__ __forceinline__ void grid_sync(Ctx& c, unsigned* bar) {
    grid_sync(c, bar, [] {});
}

struct WPhase {
    const __nv_bfloat16* w;
    int N, K;
};
__device__ __forceinline__ WPhase weight_phase(const StackLayerW& lw, int i, int d, int f) {
    switch (i) {
        case 0: return {reinterpret_cast<const __nv_bfloat16*>(lw.wqkv), 3 * d, d};
        case 1: return {reinterpret_cast<const __nv_bfloat16*>(lw.wo), d, d};
        case 2: return {reinterpret_cast<const __nv_bfloat16*>(lw.wq2), d, d};
        case 3: return {reinterpret_cast<const __nv_bfloat16*>(lw.wo2), d, d};
        case 4: return {reinterpret_cast<const __nv_bfloat16*>(lw.w1), f, d};
        default: return {reinterpret_cast<const __nv_bfloat16*>(lw.w2), d, f};
    }
}
__device__ __forceinline__ int row_begin(int cta, int N, int G) { return (int)(((long long)cta * N) / G); }

// ------------------------------------------------------------------------------------------------ activation staging
__device__ __forceinline__ void unpack_h8(const uint4& u, float (&f)[8]) {
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __half22float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ void unpack_b8(const uint4& u, float (&f)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}

// The f16 residual row of sequence b as 8 values at chunk c: the stream itself, or -- first phase of the step -- the
// token embedding + position rounded to f16 exactly as decode_embed_kernel stores it.
__device__ __forceinline__ void load_x8(const DecStackParams& p, bool embed, int b, int chunk, int tok, int step, float (&v)[8]) {
    if (embed) {
        const uint4 e = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.emb) + (size_t)tok * p.d) + chunk);
        const float4 p0 = __ldg(reinterpret_cast<const float4*>(p.pos + (size_t)step * p.d) + 2 * chunk);
        const float4 p1 = __ldg(reinterpret_cast<const float4*>(p.pos + (size_t)step * p.d) + 2 * chunk + 1);
        float ev[8];
        unpack_b8(e, ev);
        const float pv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __half2float(__float2half_rn(ev[i] + pv[i]));
    } else {
        const uint4 u = __ldcg(reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(p.x) + (size_t)b * p.d) + chunk);
        unpack_h8(u, v);
    }
}

// One warp per sequence: LayerNorm(row) * gamma + beta -> bf16.  dst == nullptr: into the activation operand.
__device__ __forceinline__ void layer_norm_row(const DecStackParams& p, const Ctx& c, int b, bool embed, const float* gamma,
                                               const float* beta, __nv_bfloat16* dst) {
    const int n_chunks = p.d / 8;
    const int step = *p.step;
    const int tok = embed ? p.tokens[(size_t)b * p.tokens_ld + step] : 0;
    float v[kLnJ][8];
#pragma unroll
    for (int j = 0; j < kLnJ; ++j) {
        const int ch = c.lane + 32 * j;
        if (ch < n_chunks) {
            load_x8(p, embed, b, ch, tok, step, v[j]);
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[j][e] = 0.0f;
        }
    }
    float sum = 0.0f;
#pragma unroll
    for (int j = 0; j < kLnJ; ++j)
#pragma unroll
        for (int e = 0; e < 8; ++e) sum += v[j][e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)p.d;
    float sq = 0.0f;
#pragma unroll
    for (int j = 0; j < kLnJ; ++j)
        if (c.lane + 32 * j < n_chunks) {
#pragma unroll
            for (int e = 0; e < 8; ++e) sq += (v[j][e] - mean) * (v[j][e] - mean);
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / (float)p.d + 1e-5f);
#pragma unroll
    for (int j = 0; j < kLnJ; ++j) {
        const int ch = c.lane + 32 * j;
        if (ch < n_chunks) {
            // gamma / beta: shared memory (an aux stage of the ring)
            const float4 g0 = reinterpret_cast<const float4*>(gamma)[2 * ch], g1 = reinterpret_cast<const float4*>(gamma)[2 * ch + 1];
            const float4 b0 = reinterpret_cast<const float4*>(beta)[2 * ch], b1 = reinterpret_cast<const float4*>(beta)[2 * ch + 1];
            const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            uint4 o4;
            uint32_t* o = reinterpret_cast<uint32_t*>(&o4);
#pragma unroll
            for (int e = 0; e < 4; ++e)
                o[e] = pack_bf16x2((v[j][2 * e] - mean) * rstd * gv[2 * e] + bv[2 * e],
                                   (v[j][2 * e + 1] - mean) * rstd * gv[2 * e + 1] + bv[2 * e + 1]);
            if (dst) *reinterpret_cast<uint4*>(dst + 8 * ch) = o4;
            else *reinterpret_cast<uint4*>(c.act + (size_t)b * c.act_pitch + 16 * ch) = o4;
        }
    }
}

// bf16 rows [batch, K] from global memory (fc1's output)
__device__ __forceinline__ void stage_rows(const DecStackParams& p, const Ctx& c, const __nv_bfloat16* src, int K) {
    const int per_row = K / 8, total = p.batch * per_row;
#pragma unroll 4
    for (int i = c.tid; i < total; i += kConsumerThreads) {
        const int b = i / per_row, ch = i - b * per_row;
        *reinterpret_cast<uint4*>(c.act + (size_t)b * c.act_pitch + 16 * ch) = __ldcg(reinterpret_cast<const uint4*>(src + (size_t)b * K) + ch);
    }
    consumer_sync();
}

// ------------------------------------------------------------------------------------------------ ring, consumer side
// An aux stage (gamma, beta or a bias slice: one bulk copy) is acquired in ring order and released once read.
__device__ __forceinline__ const float* aux_acquire(Ctx& c, int& held) {
    mbar_wait(&c.full[c.stage], c.ph);
    held = c.stage;
    const float* ptr = reinterpret_cast<const float*>(c.ring + (size_t)c.stage * kStageBytes);
    if (++c.stage == c.ns) { c.stage = 0; c.ph ^= 1; }
    return ptr;
}
__device__ __forceinline__ void aux_release(const Ctx& c, int held) {
    __syncwarp();
    if (c.lane == 0) mbar_arrive(&c.empty[held]);
}
// bias slice of the phase -> s_bias (all consumer threads call; visible after the next consumer_sync)
__device__ __forceinline__ void take_bias(Ctx& c, int N) {
    const int n0 = row_begin(c.cta, N, c.G), n1 = row_begin(c.cta + 1, N, c.G);
    if (n0 >= n1) return;
    int held;
    const float* bs = aux_acquire(c, held);
    const int a0 = n0 & ~3, cnt = ((n1 + 3) & ~3) - a0;
    if (c.tid < cnt) c.s_bias[c.tid] = bs[c.tid];
    aux_release(c, held);
}

// ------------------------------------------------------------------------------------------------ attention phase
// Items (sequence, head, split) over the CTAs; 8 lanes per key / value row (16 bytes each), 32 rows per load instruction
// of the CTA, four of them in flight per lane for keys and values alike (one pass, online softmax per 8-lane group).
struct KvRegs {
    uint4 kk[4], vv[4];
};
struct KvPair {
    KvRegs a, b;                         // rows [base, base + 128) and [base + 128, base + 256)
};
struct AttnShape {
    const __nv_bfloat16 *kb, *vb;        // key / value base of the layer
    long long kv_rows;                   // rows per sequence
    int kv_ld, n_keys, S;
};
__device__ __forceinline__ void attn_item(const DecStackParams& p, const AttnShape& a, int item, int& b, int& h, int& j0, int& j1) {
    const int per = (a.n_keys + a.S - 1) / a.S;
    const int s = item % a.S, bh = item / a.S;
    b = bh / p.heads;
    h = bh - b * p.heads;
    j0 = s * per;
    j1 = min(a.n_keys, j0 + per);
}
// rows base + 32 u + gidx, u = 0 .. 3, below `limit`
__device__ __forceinline__ void attn_fetch(const Ctx& c, const AttnShape& a, int b, int h, int base, int limit, KvRegs& r) {
    const int sub = c.lane & 7, gidx = c.warp * 4 + (c.lane >> 3);
    const __nv_bfloat16* kbase = a.kb + (size_t)b * a.kv_rows * a.kv_ld + h * 64 + sub * 8;
    const __nv_bfloat16* vbase = a.vb + (size_t)b * a.kv_rows * a.kv_ld + h * 64 + sub * 8;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int j = base + u * 32 + gidx;
        if (j < limit) {
            r.kk[u] = __ldcg(reinterpret_cast<const uint4*>(kbase + (size_t)j * a.kv_ld));
            r.vv[u] = __ldcg(reinterpret_cast<const uint4*>(vbase + (size_t)j * a.kv_ld));
        } else {
            r.kk[u] = make_uint4(0, 0, 0, 0);
            r.vv[u] = make_uint4(0, 0, 0, 0);
        }
    }
}
// Before the grid barrier that opens the phase: the first 128 rows of the CTA's first item, except rows >= fresh_from
// (self-attention: the row this step appends is written by another CTA in the phase that is still running).
__device__ __forceinline__ void attn_prefetch(const DecStackParams& p, const Ctx& c, const AttnShape& a, int fresh_from, KvPair& r) {
    if (c.cta >= p.batch * p.heads * a.S) return;
    int b, h, j0, j1;
    attn_item(p, a, c.cta, b, h, j0, j1);
    attn_fetch(c, a, b, h, j0, min(j1, fresh_from), r.a);
    attn_fetch(c, a, b, h, j0 + 128, min(j1, fresh_from), r.b);
}

__device__ __forceinline__ void attention_phase(const DecStackParams& p, const Ctx& c, const AttnShape& a, const __nv_bfloat16* q,
                                                int q_ld, int fresh_from, KvPair& r) {
    const int H = p.heads, items = p.batch * H * a.S;
    const int grp = c.lane >> 3, sub = c.lane & 7, gidx = c.warp * 4 + grp;
    float* s_red = reinterpret_cast<float*>(c.act);           // [8 warps][66] (the activation operand is idle here)
    volatile unsigned& s_last = *c.s_flag;
    for (int item = c.cta; item < items; item += c.G) {
        int b, h, j0, j1;
        attn_item(p, a, item, b, h, j0, j1);
        if (p.done != nullptr && p.done[b] != 0) continue;      // a finished sequence stops reading its caches
        float qv[8];
        {
            const uint4 u = __ldcg(reinterpret_cast<const uint4*>(q + (size_t)b * q_ld + h * 64) + sub);
            unpack_b8(u, qv);
            const float sc = 0.125f * 1.4426950408889634f;
#pragma unroll
            for (int i = 0; i < 8; ++i) qv[i] *= sc;
        }
        if (item == c.cta) {
            // rows the prefetch had to leave out (written during the previous phase)
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j = j0 + u * 32 + gidx;
                if (j >= fresh_from && j < j1) {
                    const size_t at = ((size_t)b * a.kv_rows + j) * a.kv_ld + h * 64 + sub * 8;
                    const uint4 kx = __ldcg(reinterpret_cast<const uint4*>(a.kb + at));
                    const uint4 vx = __ldcg(reinterpret_cast<const uint4*>(a.vb + at));
                    if (u < 4) { r.a.kk[u & 3] = kx; r.a.vv[u & 3] = vx; }
                    else { r.b.kk[u & 3] = kx; r.b.vv[u & 3] = vx; }
                }
            }
        } else {
            attn_fetch(c, a, b, h, j0, j1, r.a);
            attn_fetch(c, a, b, h, j0 + 128, j1, r.b);
        }
        float m = -INFINITY, l = 0.0f, acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        auto consume = [&](const KvRegs& x, int base) {           // 128 rows: scores, online softmax, weighted values
            float dot[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float f[8];
                unpack_b8(x.kk[u], f);
                float dd = 0.0f;
#pragma unroll
                for (int i = 0; i < 8; ++i) dd = fmaf(qv[i], f[i], dd);
                dd += __shfl_xor_sync(0xffffffffu, dd, 1);
                dd += __shfl_xor_sync(0xffffffffu, dd, 2);
                dd += __shfl_xor_sync(0xffffffffu, dd, 4);
                dot[u] = (base + u * 32 + gidx < j1) ? dd : -INFINITY;
            }
            const float mn = fmaxf(fmaxf(m, fmaxf(dot[0], dot[1])), fmaxf(dot[2], dot[3]));
            if (mn != -INFINITY) {
                const float sc = exp2f(m - mn);                   // m = -inf: 0 (nothing accumulated yet)
                l *= sc;
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] *= sc;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float pj = exp2f(dot[u] - mn);          // masked rows: exp2(-inf) = 0
                    float f[8];
                    unpack_b8(x.vv[u], f);
                    l += pj;
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[i] = fmaf(pj, f[i], acc[i]);
                }
                m = mn;
            }
        };
        for (int base = j0; base < j1; base += 256) {             // CTA-uniform trip count (shuffles inside)
            consume(r.a, base);
            if (base + 256 < j1) attn_fetch(c, a, b, h, base + 256, j1, r.a);
            if (base + 128 < j1) {
                consume(r.b, base + 128);
                if (base + 384 < j1) attn_fetch(c, a, b, h, base + 384, j1, r.b);
            }
        }
        // the 4 groups of a warp (same `sub`, lane ^ 8, lane ^ 16), then the 8 warps
#pragma unroll
        for (int off = 8; off <= 16; off <<= 1) {
            const float mo = __shfl_xor_sync(0xffffffffu, m, off);
            const float M = fmaxf(m, mo);
            const float sc = (m == M) ? 1.0f : exp2f(m - M);      // (-inf, -inf) -> 1 on zeros
            l *= sc;
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] *= sc;
            l += __shfl_xor_sync(0xffffffffu, l, off);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
            m = M;
        }
        consumer_sync();                                          // the previous item's merge has read s_red
        if (grp == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) s_red[c.warp * 66 + sub * 8 + i] = acc[i];
            if (sub == 0) {
                s_red[c.warp * 66 + 64] = l;
                s_red[c.warp * 66 + 65] = m;
            }
        }
        consumer_sync();
        float o = 0.0f, ls = 0.0f, M = -INFINITY;
        if (c.tid < 64) {
#pragma unroll
            for (int w = 0; w < 8; ++w) M = fmaxf(M, s_red[w * 66 + 65]);
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                const float mw = s_red[w * 66 + 65];
                if (mw != -INFINITY) {
                    const float sc = exp2f(mw - M);
                    o = fmaf(sc, s_red[w * 66 + c.tid], o);
                    ls = fmaf(sc, s_red[w * 66 + 64], ls);
                }
            }
        }
        __nv_bfloat16* ctx = reinterpret_cast<__nv_bfloat16*>(p.ctx) + (size_t)b * p.d + h * 64;
        if (a.S == 1) {
            if (c.tid < 64) ctx[c.tid] = __float2bfloat16_rn(o / ls);
            continue;
        }
        // several splits: park the partial; the last split of (b, h) to arrive merges all of them in split order
        const int bh = b * H + h;
        float* pp = p.part + (size_t)item * kPartStride;          // item = (b H + h) S + s
        if (c.tid < 64) {
            __stcg(pp + c.tid, o);
            if (c.tid == 0) {
                __stcg(pp + 64, ls);
                __stcg(pp + 65, M);
            }
        }
        consumer_sync();
        if (c.tid == 0) {
            __threadfence();
            const unsigned old = atomicAdd(p.cnt + bh, 1u);
            s_last = (old == (unsigned)(a.S - 1)) ? 1u : 0u;
            if (s_last) {
                p.cnt[bh] = 0u;                                   // ready for the next attention phase
                __threadfence();
            }
        }
        consumer_sync();
        if (s_last && c.tid < 64) {
            const float* p0 = p.part + (size_t)bh * a.S * kPartStride;
            float ms[8], lsv[8], ov[8];
#pragma unroll
            for (int s = 0; s < 8; ++s)
                if (s < a.S) {
                    ms[s] = __ldcg(p0 + s * kPartStride + 65);
                    lsv[s] = __ldcg(p0 + s * kPartStride + 64);
                    ov[s] = __ldcg(p0 + s * kPartStride + c.tid);
                }
            float MM = -INFINITY;
#pragma unroll
            for (int s = 0; s < 8; ++s)
                if (s < a.S) MM = fmaxf(MM, ms[s]);
            float oo = 0.0f, ll = 0.0f;
#pragma unroll
            for (int s = 0; s < 8; ++s)
                if (s < a.S && ms[s] != -INFINITY) {
                    const float w = exp2f(ms[s] - MM);
                    oo = fmaf(w, ov[s], oo);
                    ll = fmaf(w, lsv[s], ll);
                }
            ctx[c.tid] = __float2bfloat16_rn(oo / ll);
        }
    }
    // (the grid barrier that follows opens with a consumer_sync: s_red is free before the next phase stages into it)
}

// ------------------------------------------------------------------------------------------------ weight phases
enum StackEpilogue { ST_QKV = 0, ST_RESID = 1, ST_BF16 = 2, ST_GELU = 3 };

// out rows [n0, n1) of one weight matrix against the staged activation operand; the stages arrive in the order the
// producer warp issues them (same loop nest).
template <int EPI>
__device__ __forceinline__ void weight_phase_run(const DecStackParams& p, Ctx& c, int N, int K, int layer, bool embed_resid,
                                                 __nv_bfloat16* out_bf16, int out_ld) {
    const int n0 = row_begin(c.cta, N, c.G), n1 = row_begin(c.cta + 1, N, c.G);
    const int g = c.lane >> 2, t = c.lane & 3;
    const int er = c.tid >> 3, eb = c.tid & 7;                    // epilogue: tile row, sequence (threads 0 .. 127)
    const int step = *p.step;
    for (int t0 = n0; t0 < n1; t0 += kTileRows) {
        // epilogue inputs first: their latency hides under the tile's stages
        const int n = t0 + er;
        const bool own = c.tid < 128 && n < n1 && eb < p.batch;
        float bias_v = 0.0f, resid_v = 0.0f;
        if (own) {
            bias_v = c.s_bias[n - (n0 & ~3)];
            if (EPI == ST_RESID) {
                if (embed_resid) {
                    const int tok = p.tokens[(size_t)eb * p.tokens_ld + step];
                    const float e = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.emb)[(size_t)tok * p.d + n]);
                    resid_v = __half2float(__float2half_rn(e + __ldg(p.pos + (size_t)step * p.d + n)));
                } else {
                    resid_v = __half2float(__ldcg(reinterpret_cast<const __half*>(p.x) + (size_t)eb * p.d + n));
                }
            }
        }
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f}, acc2[4] = {0.0f, 0.0f, 0.0f, 0.0f};      // two chains: the MMAs of a stage overlap
        for (int k0 = 0; k0 < K; k0 += kKBlock) {
            const int klen = min(kKBlock, K - k0);
            mbar_wait(&c.full[c.stage], c.ph);
            if (c.warp * 32 < klen) {
                const uint8_t* sa = c.ring + (size_t)c.stage * kStageBytes + g * kRowPitch + c.warp * 64 + t * 16;
                const uint4 alo = *reinterpret_cast<const uint4*>(sa);
                const uint4 ahi = *reinterpret_cast<const uint4*>(sa + 8 * kRowPitch);
                const uint4 bb = *reinterpret_cast<const uint4*>(c.act + (size_t)g * c.act_pitch + (size_t)(k0 + c.warp * 32) * 2 + t * 16);
                // k is permuted inside each 32-deep slice (thread t feeds physical k 8t .. 8t+7 into both operands): the
                // contraction does not care, and every operand read is one 16-byte load
                mma_16816(acc, alo.x, ahi.x, alo.y, ahi.y, bb.x, bb.y);
                mma_16816(acc2, alo.z, ahi.z, alo.w, ahi.w, bb.z, bb.w);
            }
            __syncwarp();
            if (c.lane == 0) mbar_arrive(&c.empty[c.stage]);
            if (++c.stage == c.ns) { c.stage = 0; c.ph ^= 1; }
        }
        float* red = c.red + c.red_buf * (8 * 128);
        *reinterpret_cast<float2*>(red + c.warp * 128 + g * 8 + 2 * t) = make_float2(acc[0] + acc2[0], acc[1] + acc2[1]);
        *reinterpret_cast<float2*>(red + c.warp * 128 + (g + 8) * 8 + 2 * t) = make_float2(acc[2] + acc2[2], acc[3] + acc2[3]);
        consumer_sync();
        if (own) {
            float v = 0.0f;
#pragma unroll
            for (int w = 0; w < 8; ++w) v += red[w * 128 + c.tid];
            v += bias_v;
            if (EPI == ST_QKV) {
                const __nv_bfloat16 o = __float2bfloat16_rn(v);
                const int d = p.d;
                if (n < d) {
                    out_bf16[(size_t)eb * out_ld + n] = o;
                } else {
                    __nv_bfloat16* cache = reinterpret_cast<__nv_bfloat16*>(n < 2 * d ? p.kc : p.vc);
                    cache[(((size_t)layer * p.max_batch + eb) * p.C + step) * d + (n < 2 * d ? n - d : n - 2 * d)] = o;
                }
            } else if (EPI == ST_RESID) {
                v += resid_v;
                reinterpret_cast<__half*>(p.x)[(size_t)eb * p.d + n] = __float2half_rn(fminf(fmaxf(v, -65504.0f), 65504.0f));
            } else {
                if (EPI == ST_GELU) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
                out_bf16[(size_t)eb * out_ld + n] = __float2bfloat16_rn(v);
            }
        }
        c.red_buf ^= 1;               // the next tile's partials go to the other buffer: one barrier per tile
    }
}

// LayerNorm-fed phase: gamma, beta and the bias slice arrive as aux stages ahead of the phase's weight stages
__device__ __forceinline__ void open_ln_phase(const DecStackParams& p, Ctx& c, bool embed, int N) {
    int hg, hb;
    const float* g = aux_acquire(c, hg);
    const float* b = aux_acquire(c, hb);
    take_bias(c, N);
    if (c.warp < p.batch) layer_norm_row(p, c, c.warp, embed, g, b, nullptr);
    aux_release(c, hg);
    aux_release(c, hb);
    consumer_sync();
}
__device__ __forceinline__ void open_rows_phase(const DecStackParams& p, Ctx& c, const __nv_bfloat16* src, int K, int N) {
    take_bias(c, N);
    stage_rows(p, c, src, K);
}

__global__ void __launch_bounds__(kStackThreads, 1) decode_stack_kernel(const DecStackParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    Ctx c;
    c.full = reinterpret_cast<uint64_t*>(smem);
    c.empty = c.full + kMaxStages;
    c.s_bias = reinterpret_cast<float*>(smem + 512);
    c.s_flag = reinterpret_cast<unsigned*>(smem + 448);
    c.act = smem + kBarBytes;
    const int kmax = p.d > p.f ? p.d : p.f;
    c.act_pitch = kmax * 2 + 64;
    c.red = reinterpret_cast<float*>(c.act + kMaxB * c.act_pitch);
    c.ring = reinterpret_cast<uint8_t*>(c.red) + kRedBytes;
    c.ns = p.n_stages;
    c.stage = 0;
    c.ph = 0;
    c.red_buf = 0;
    c.G = gridDim.x;
    c.cta = blockIdx.x;
    c.tid = threadIdx.x;
    c.warp = threadIdx.x >> 5;
    c.lane = threadIdx.x & 31;
    c.bar_target = 0;
    c.ti = 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < c.ns; ++i) {
            mbar_init(&c.full[i], 32);          // every producer lane arrives once its own copies have landed
            mbar_init(&c.empty[i], 8);
        }
        fence_mbar_init();
    }
    __syncthreads();
    const int d = p.d, f = p.f;

    if (c.warp == 8) {
        // ------------------------------------------------------------ producer: the token's weight stream, in order
        int stage = 0;
        uint32_t ph = 0;
        // The copies are per-lane 16-byte cp.async (one instruction moves a 512-byte row piece, each lane with its own
        // address), completion through cp.async.mbarrier.arrive.  cp.async.bulk was measured first: its operands live in
        // uniform registers, so a stage's 16 row copies cost ~480 cycles of issue whichever way they were spread over the
        // lanes, and the whole kernel ran at the producer's issue rate.
        auto emit = [&](const void* src, uint32_t bytes) {                // one aux stage: `bytes` contiguous bytes
            mbar_wait(&c.empty[stage], ph ^ 1);
            uint8_t* dst = c.ring + (size_t)stage * kStageBytes;
            for (uint32_t off = (uint32_t)c.lane * 16; off < bytes; off += 512)
                cp_async16(dst + off, reinterpret_cast<const char*>(src) + off);
            cp_async_arrive(&c.full[stage]);
            if (++stage == c.ns) { stage = 0; ph ^= 1; }
        };
        for (int l = 0; l < p.n_layers; ++l) {
            const StackLayerW& lw = p.layers[l];
            for (int i = 0; i < 6; ++i) {
                const WPhase wp = weight_phase(lw, i, d, f);
                const int n0 = row_begin(c.cta, wp.N, c.G), n1 = row_begin(c.cta + 1, wp.N, c.G);
                if (i == 0) { emit(lw.ln1_g, d * 4); emit(lw.ln1_b, d * 4); }
                if (i == 2) { emit(lw.ln2_g, d * 4); emit(lw.ln2_b, d * 4); }
                if (i == 4) { emit(lw.ln3_g, d * 4); emit(lw.ln3_b, d * 4); }
                if (n0 < n1) {
                    const float* bias = i == 0 ? lw.bqkv : i == 1 ? lw.bo : i == 2 ? lw.bq2 : i == 3 ? lw.bo2 : i == 4 ? lw.b1 : lw.b2;
                    const int a0 = n0 & ~3, a1 = (n1 + 3) & ~3;
                    emit(bias + a0, (uint32_t)(a1 - a0) * 4);
                }
                for (int t0 = n0; t0 < n1; t0 += kTileRows) {
                    const int rows = min(kTileRows, n1 - t0);
                    for (int k0 = 0; k0 < wp.K; k0 += kKBlock) {
                        const int klen = min(kKBlock, wp.K - k0);
                        mbar_wait(&c.empty[stage], ph ^ 1);
                        {
                            // fully unrolled, independent addresses: a single warp issues dependent instructions ~4 cycles
                            // apart, and a rolled loop (address chain + copy) held the producer to ~30 cycles per row copy
                            const uint32_t dst = smem_u32(c.ring) + (uint32_t)(stage * kStageBytes + c.lane * 16);
                            const char* src = reinterpret_cast<const char*>(wp.w + (size_t)t0 * wp.K + k0) + c.lane * 16;
                            const size_t pitch = (size_t)wp.K * 2;
                            if (c.lane * 16 < klen * 2) {
#pragma unroll
                                for (int r = 0; r < kTileRows; ++r)
                                    if (r < rows)
                                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + r * kRowPitch), "l"(src + r * pitch) : "memory");
                            }
                            cp_async_arrive(&c.full[stage]);
                        }
                        if (++stage == c.ns) { stage = 0; ph ^= 1; }
                    }
                }
            }
        }
        if (c.cta < p.batch) { emit(p.lnf_g, d * 4); emit(p.lnf_b, d * 4); }
        return;
    }

    // ---------------------------------------------------------------- consumers
    const int step = *p.step;
    const __nv_bfloat16* q_self = reinterpret_cast<const __nv_bfloat16*>(p.q);
    const __nv_bfloat16* q_cross = reinterpret_cast<const __nv_bfloat16*>(p.q2);
    const __nv_bfloat16* ctx = reinterpret_cast<const __nv_bfloat16*>(p.ctx);
    KvPair kv;
    for (int l = 0; l < p.n_layers; ++l) {
        AttnShape self_a, cross_a;
        {
            const size_t off = (size_t)l * p.max_batch * p.C * d;
            self_a.kb = reinterpret_cast<const __nv_bfloat16*>(p.kc) + off;
            self_a.vb = reinterpret_cast<const __nv_bfloat16*>(p.vc) + off;
            self_a.kv_rows = p.C; self_a.kv_ld = d; self_a.n_keys = step + 1;
            self_a.S = min(p.splits_self, (step + 128) / 128);        // a CTA covers 128 keys per load round anyway
            const __nv_bfloat16* xk = reinterpret_cast<const __nv_bfloat16*>(p.xkv) + (size_t)l * p.batch * p.A * 2 * d;
            cross_a.kb = xk; cross_a.vb = xk + d;
            cross_a.kv_rows = p.A; cross_a.kv_ld = 2 * d; cross_a.n_keys = p.A; cross_a.S = p.splits_cross;
        }
        // LN1 + QKV: q to the step buffer, k / v appended to the caches at position `step`
        STACK_TRACE(p, c);
        open_ln_phase(p, c, l == 0, 3 * d);
        STACK_TRACE(p, c);
        weight_phase_run<ST_QKV>(p, c, 3 * d, d, l, false, reinterpret_cast<__nv_bfloat16*>(p.q), d);
        STACK_TRACE(p, c);
        grid_sync(c, p.bar, [&] { attn_prefetch(p, c, self_a, step, kv); });
        STACK_TRACE(p, c);
        attention_phase(p, c, self_a, q_self, d, step, kv);
        STACK_TRACE(p, c);
        grid_sync(c, p.bar);
        STACK_TRACE(p, c);
        open_rows_phase(p, c, ctx, d, d);
        STACK_TRACE(p, c);
        weight_phase_run<ST_RESID>(p, c, d, d, l, l == 0, nullptr, 0);
        STACK_TRACE(p, c);
        grid_sync(c, p.bar);
        STACK_TRACE(p, c);
        open_ln_phase(p, c, false, d);
        STACK_TRACE(p, c);
        weight_phase_run<ST_BF16>(p, c, d, d, l, false, reinterpret_cast<__nv_bfloat16*>(p.q2), d);
        STACK_TRACE(p, c);
        grid_sync(c, p.bar, [&] { attn_prefetch(p, c, cross_a, p.A, kv); });
        STACK_TRACE(p, c);
        attention_phase(p, c, cross_a, q_cross, d, p.A, kv);
        STACK_TRACE(p, c);
        grid_sync(c, p.bar);
        STACK_TRACE(p, c);
        open_rows_phase(p, c, ctx, d, d);
        STACK_TRACE(p, c);
        weight_phase_run<ST_RESID>(p, c, d, d, l, false, nullptr, 0);
        STACK_TRACE(p, c);
        grid_sync(c, p.bar);
        STACK_TRACE(p, c);
        open_ln_phase(p, c, false, f);
        STACK_TRACE(p, c);
        weight_phase_run<ST_GELU>(p, c, f, d, l, false, reinterpret_cast<__nv_bfloat16*>(p.h), f);
        STACK_TRACE(p, c);
        grid_sync(c, p.bar);
        STACK_TRACE(p, c);
        open_rows_phase(p, c, reinterpret_cast<const __nv_bfloat16*>(p.h), f, d);
        STACK_TRACE(p, c);
        weight_phase_run<ST_RESID>(p, c, d, f, l, false, nullptr, 0);
        STACK_TRACE(p, c);
        grid_sync(c, p.bar);
        STACK_TRACE(p, c);
    }
    // final LayerNorm -> the logits projection's operand (one warp per sequence, CTA b)
    if (c.cta < p.batch) {
        int hg, hb;
        const float* g = aux_acquire(c, hg);
        const float* b = aux_acquire(c, hb);
        if (c.warp == 0) layer_norm_row(p, c, c.cta, false, g, b, reinterpret_cast<__nv_bfloat16*>(p.y) + (size_t)c.cta * d);
    }
}

}  // namespace

int decode_stack_stages(int d, int f) {
    const int kmax = d > f ? d : f;
    const int fixed = kBarBytes + kMaxB * (kmax * 2 + 64) + kRedBytes;
    int ns = (kSmemLimit - fixed) / kStageBytes;
    if (ns > kMaxStages) ns = kMaxStages;
    return ns;
}

bool decode_stack_supported(int d, int f, int heads, int batch, int sm_count) {
    // (bias slices of at most 64 floats: ceil(N / G) + 6 <= 64 for every N)
    const int nmax = (3 * d > f ? 3 * d : f);
    return batch >= 1 && batch <= kMaxB && d % 64 == 0 && f % 64 == 0 && d <= kLnJ * 256 && heads * 64 == d && heads <= kMaxHeads &&
           sm_count >= 8 && (nmax + sm_count - 1) / sm_count + 7 <= 64 && decode_stack_stages(d, f) >= 4;
}

size_t decode_stack_part_floats(int heads) { return (size_t)kMaxB * heads * 8 * kPartStride + (size_t)kMaxB * heads; }

cudaError_t decode_stack_launch(const DecStackParams& p0, int sm_count, cudaStream_t stream) {
    if (!decode_stack_supported(p0.d, p0.f, p0.heads, p0.batch, sm_count)) return cudaErrorInvalidValue;
    DecStackParams p = p0;
    p.n_stages = decode_stack_stages(p.d, p.f);
    const int items = p.batch * p.heads;
    int s = sm_count / items;
    s = s < 1 ? 1 : (s > 8 ? 8 : s);
    p.splits_self = s;
    p.splits_cross = s;
    p.cnt = reinterpret_cast<unsigned*>(p.part + (size_t)kMaxB * p.heads * 8 * kPartStride);     // zero at allocation, self-resetting
    const int kmax = p.d > p.f ? p.d : p.f;
    const int smem = kBarBytes + kMaxB * (kmax * 2 + 64) + kRedBytes + p.n_stages * kStageBytes;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(decode_stack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sm_count);
    cfg.blockDim = dim3(kStackThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;       // all CTAs co-resident, or the launch fails: the grid barrier is safe
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, decode_stack_kernel, p);
}

}  // namespace aries
