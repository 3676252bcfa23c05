// Token embedding and greedy sampling kernels of the decode step (row f1).
//
// decode_sample_kernel is the device-side mirror of one greedy step of ctranslate2.models.Whisper.generate as
// faster-whisper drives it for the reference (beam_size=1, temperature=0; ref: final_optimized_transcriber.py:432-441):
// logits processors SuppressTokensBegin (suppress_blank), SuppressTokens and ApplyTimestampRules, then argmax and the
// cumulative log-probability of the processed distribution (CT2 src/models/whisper.cc, GreedySearch) [unverified
// offline; restated in oracle/whisper_decoder.py apply_rules, which is pinned against HF's
// WhisperTimeStampLogitsProcessor].  Everything stays on the device so that the step can be replayed as a CUDA graph
// without a host round trip per token.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "layernorm.h"
#include "skinny.h"

namespace aries {

namespace {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__global__ void __launch_bounds__(128) decode_embed_kernel(const int* __restrict__ tokens, int tokens_ld, const int* step,
                                                           const __nv_bfloat16* __restrict__ emb,
                                                           const float* __restrict__ pos, __half* __restrict__ x, int d,
                                                           float2* __restrict__ stats, int stats_parts) {
    __shared__ float2 s_part[4];
    pdl_wait();
    pdl_trigger();
    const int b = blockIdx.x;
    const int t = *step;
    const int tok = tokens[(size_t)b * tokens_ld + t];
    const __nv_bfloat162* e = reinterpret_cast<const __nv_bfloat162*>(emb + (size_t)tok * d);
    const float2* p2 = reinterpret_cast<const float2*>(pos + (size_t)t * d);
    __half2* xo = reinterpret_cast<__half2*>(x + (size_t)b * d);
    float s1 = 0.0f, s2 = 0.0f;
    for (int i = threadIdx.x; i < d / 2; i += blockDim.x) {
        const float2 a = __bfloat1622float2(e[i]);
        const float2 c = p2[i];
        const __half2 h = __floats2half2_rn(a.x + c.x, a.y + c.y);
        xo[i] = h;
        const float2 r = __half22float2(h);
        s1 += r.x + r.y;
        s2 += r.x * r.x + r.y * r.y;
    }
    if (stats == nullptr) return;
    // folded LayerNorm (decoder.cu): the row's (sum, sum of squares) in slot 0 of the layout the residual GEMMs write
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = make_float2(s1, s2);
    __syncthreads();
    if (threadIdx.x < stats_parts) {
        float2 v = make_float2(0.0f, 0.0f);
        if (threadIdx.x == 0)
            v = make_float2(s_part[0].x + s_part[1].x + s_part[2].x + s_part[3].x, s_part[0].y + s_part[1].y + s_part[2].y + s_part[3].y);
        stats[(size_t)b * stats_parts + threadIdx.x] = v;
    }
}

// One warp per row; three passes over a row that stays in L1 (rows = sequences: a few KB in total).
__global__ void __launch_bounds__(128) decode_layernorm_kernel(const __half* __restrict__ x, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta,
                                                               __nv_bfloat16* __restrict__ y, int rows, int d) {
    pdl_wait();
    pdl_trigger();
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const __half2* xr = reinterpret_cast<const __half2*>(x + (size_t)row * d);
    float s = 0.0f;
    for (int i = lane; i < d / 2; i += 32) {
        const float2 v = __half22float2(xr[i]);
        s += v.x + v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / d;
    float q = 0.0f;
    for (int i = lane; i < d / 2; i += 32) {
        const float2 v = __half22float2(xr[i]);
        q += (v.x - mean) * (v.x - mean) + (v.y - mean) * (v.y - mean);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / d + 1e-5f);
    __nv_bfloat162* yr = reinterpret_cast<__nv_bfloat162*>(y + (size_t)row * d);
    const float2* g2 = reinterpret_cast<const float2*>(gamma);
    const float2* b2 = reinterpret_cast<const float2*>(beta);
    for (int i = lane; i < d / 2; i += 32) {
        const float2 v = __half22float2(xr[i]);
        const float2 g = g2[i], bb = b2[i];
        yr[i] = __floats2bfloat162_rn((v.x - mean) * rstd * g.x + bb.x, (v.y - mean) * rstd * g.y + bb.y);
    }
}

constexpr int kSampleThreads = 1024;     // one CTA per sequence; 16-byte loads, ~13 per thread and pass at 51866 ids

struct Best {
    float v;
    int i;
};
__device__ __forceinline__ Best better(Best a, Best b) {       // larger value wins, lower index on ties
    return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}
__device__ __forceinline__ Best warp_best(Best x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Best y;
        y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
        y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
        x = better(x, y);
    }
    return x;
}
__device__ __forceinline__ float warp_max(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}
__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

__global__ void __launch_bounds__(kSampleThreads) decode_sample_kernel(const SampleParams p) {
    __shared__ float s_f[4][kSampleThreads / 32];
    __shared__ Best s_b[2][kSampleThreads / 32];
    __shared__ unsigned s_last;

    pdl_wait();
    pdl_trigger();
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int kWarps = kSampleThreads / 32;
    if (p.stack_bar != nullptr && b == 0 && tid == 0) *p.stack_bar = 0u;     // the stack kernel of this step has exited
    const int t = *p.step;
    const int pos = t + 1;
    const float* lg = p.logits + (size_t)b * p.logits_ld;
    int* toks = p.tokens + (size_t)b * p.tokens_ld;
    const int V = p.vocab, tb = p.timestamp_begin;

    // rows are 16-byte aligned (logits_ld is a multiple of 128) and padded past the vocabulary, so every pass reads
    // float4 and masks n >= V
    const float4* lg4 = reinterpret_cast<const float4*>(lg);
    const int n4 = (V + 3) / 4;

    // ---- no-speech probability: softmax of the raw logits at the <|startoftranscript|> position
    if (t == p.sot_index[b]) {
        float m = -INFINITY;
#pragma unroll 4
        for (int i = tid; i < n4; i += kSampleThreads) {
            const float4 v = lg4[i];
            const int n = 4 * i;
            m = fmaxf(m, v.x);
            if (n + 1 < V) m = fmaxf(m, v.y);
            if (n + 2 < V) m = fmaxf(m, v.z);
            if (n + 3 < V) m = fmaxf(m, v.w);
        }
        m = warp_max(m);
        if (lane == 0) s_f[0][warp] = m;
        __syncthreads();
        m = s_f[0][0];
        for (int w = 1; w < kWarps; ++w) m = fmaxf(m, s_f[0][w]);
        float s = 0.0f;
#pragma unroll 4
        for (int i = tid; i < n4; i += kSampleThreads) {
            const float4 v = lg4[i];
            const int n = 4 * i;
            s += __expf(v.x - m);
            if (n + 1 < V) s += __expf(v.y - m);
            if (n + 2 < V) s += __expf(v.z - m);
            if (n + 3 < V) s += __expf(v.w - m);
        }
        s = warp_sum(s);
        if (lane == 0) s_f[1][warp] = s;
        __syncthreads();
        if (tid == 0) {
            float tot = 0.0f;
            for (int w = 0; w < kWarps; ++w) tot += s_f[1][w];
            p.no_speech_prob[b] = __expf(lg[p.no_speech] - m) / tot;
        }
        __syncthreads();
    }

    if (pos < p.max_length && pos >= p.prompt_len) {
        if (p.done[b]) {
            if (tid == 0) toks[pos] = p.eot;
        } else {
            const int n_sampled = pos - p.prompt_len;
            const int last = n_sampled >= 1 ? toks[pos - 1] : -1;
            const int penult = n_sampled >= 2 ? toks[pos - 2] : -1;
            const bool ts_on = p.use_timestamps[b] != 0;
            const bool first = n_sampled == 0;
            const bool last_ts = n_sampled >= 1 && last >= tb;
            const bool penult_ts = n_sampled < 2 || penult >= tb;
            const bool no_ts = ts_on && last_ts && penult_ts;           // a pair was just closed: text next
            const bool no_text = ts_on && last_ts && !penult_ts;        // an unpaired timestamp: timestamp or EOT next
            const int lt = p.last_timestamp[b];
            const int ts_floor = (ts_on && lt >= 0) ? ((last_ts && !penult_ts) ? lt : lt + 1) : tb;
            const int ts_ceil = (ts_on && first) ? tb + p.max_initial_timestamp_index : V - 1;
            auto allowed = [&](int n) -> bool {
                if ((p.suppress_bits[n >> 5] >> (n & 31)) & 1u) return false;
                if (first && p.suppress_blank && (n == p.blank_id || n == p.eot)) return false;
                if (!ts_on) return true;
                if (n == p.no_timestamps) return false;
                if (n >= tb) return !no_ts && n >= ts_floor && n <= ts_ceil;
                if (first) return false;
                if (no_text && n < p.eot) return false;
                return true;
            };
            // ---- pass A: maxima
            Best all{-INFINITY, 0x7fffffff}, ts{-INFINITY, 0x7fffffff};
            float text_max = -INFINITY;
#pragma unroll 2
            for (int i = tid; i < n4; i += kSampleThreads) {
                const float4 q = lg4[i];
                const float vv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int n = 4 * i + u;
                    if (n >= V || !allowed(n)) continue;
                    const float v = vv[u];
                    all = better(all, Best{v, n});
                    if (n >= tb) ts = better(ts, Best{v, n});
                    else text_max = fmaxf(text_max, v);
                }
            }
            all = warp_best(all);
            ts = warp_best(ts);
            text_max = warp_max(text_max);
            if (lane == 0) {
                s_b[0][warp] = all;
                s_b[1][warp] = ts;
                s_f[0][warp] = text_max;
            }
            __syncthreads();
            all = s_b[0][0];
            ts = s_b[1][0];
            text_max = s_f[0][0];
            for (int w = 1; w < kWarps; ++w) {
                all = better(all, s_b[0][w]);
                ts = better(ts, s_b[1][w]);
                text_max = fmaxf(text_max, s_f[0][w]);
            }
            // ---- pass B: log-sum-exp of everything allowed and of the timestamps
            float sum_all = 0.0f, sum_ts = 0.0f;
#pragma unroll 2
            for (int i = tid; i < n4; i += kSampleThreads) {
                const float4 q = lg4[i];
                const float vv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int n = 4 * i + u;
                    if (n >= V || !allowed(n)) continue;
                    const float v = vv[u];
                    sum_all += __expf(v - all.v);
                    if (n >= tb) sum_ts += __expf(v - ts.v);
                }
            }
            sum_all = warp_sum(sum_all);
            sum_ts = warp_sum(sum_ts);
            if (lane == 0) {
                s_f[2][warp] = sum_all;
                s_f[3][warp] = sum_ts;
            }
            __syncthreads();
            if (tid == 0) {
                float sa = 0.0f, st = 0.0f;
                for (int w = 0; w < kWarps; ++w) {
                    sa += s_f[2][w];
                    st += s_f[3][w];
                }
                const float ts_lse = (ts.v == -INFINITY) ? -INFINITY : ts.v + logf(st);
                int chosen = all.i;
                float lse = all.v + logf(sa);
                if (ts_on && ts_lse > text_max) {          // the timestamp mass beats every text token: sample a timestamp
                    chosen = ts.i;
                    lse = ts_lse;
                }
                if (chosen < 0 || chosen >= V) chosen = p.eot;       // every id forbidden (a degenerate suppress list): end
                if (p.argmax_out) p.argmax_out[(size_t)b * p.tokens_ld + pos] = chosen;
                int nxt = chosen;
                if (p.forced && n_sampled < p.n_forced) nxt = p.forced[(size_t)b * p.forced_ld + n_sampled];
                if (lse > -INFINITY) p.score[b] += lg[nxt] - lse;
                toks[pos] = nxt;
                if (nxt == p.eot) {
                    p.done[b] = 1;
                    atomicAdd(p.n_done, 1);
                } else if (nxt >= tb) {
                    p.last_timestamp[b] = nxt;
                }
            }
        }
    }
    // ---- the last sequence to finish advances the step counter
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        s_last = atomicAdd(p.ticket, 1u);
        if (s_last == (unsigned)(p.batch - 1)) {
            *p.ticket = 0u;
            *p.step = t + 1;
        }
    }
}

}  // namespace

cudaError_t decode_embed_launch(const int* tokens, int tokens_ld, const int* step, const void* emb_bf16, const float* pos,
                                void* x_f16, int batch, int d, int pdl, cudaStream_t stream, float2* stats, int stats_parts) {
    if (batch <= 0 || d % 2 != 0 || (stats && (stats_parts < 1 || stats_parts > 128))) return cudaErrorInvalidValue;
    const __nv_bfloat16* emb = reinterpret_cast<const __nv_bfloat16*>(emb_bf16);
    __half* x = reinterpret_cast<__half*>(x_f16);
    void* args[] = {&tokens, &tokens_ld, &step, &emb, &pos, &x, &d, &stats, &stats_parts};
    return launch_maybe_pdl(reinterpret_cast<const void*>(decode_embed_kernel), dim3(batch), dim3(128), 0, stream, args,
                            pdl != 0);
}

cudaError_t decode_layernorm_launch(const void* x_f16, const float* gamma, const float* beta, void* y_bf16, int rows, int d,
                                    int pdl, cudaStream_t stream) {
    if (rows <= 0 || d % 2 != 0) return cudaErrorInvalidValue;
    // Whisper widths (multiples of 128 the encoder kernel is instantiated for): the row stays in registers, one pass
    if (layernorm_supports(d)) return layernorm_launch_pdl(x_f16, gamma, beta, y_bf16, rows, d, 1e-5f, stream, pdl != 0);
    const __half* x = reinterpret_cast<const __half*>(x_f16);
    __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(y_bf16);
    void* args[] = {&x, &gamma, &beta, &y, &rows, &d};
    return launch_maybe_pdl(reinterpret_cast<const void*>(decode_layernorm_kernel), dim3((rows + 3) / 4), dim3(128), 0,
                            stream, args, pdl != 0);
}

cudaError_t decode_sample_launch(const SampleParams& p, cudaStream_t stream) {
    if (p.batch <= 0 || p.vocab <= 0) return cudaErrorInvalidValue;
    SampleParams q = p;
    void* args[] = {&q};
    return launch_maybe_pdl(reinterpret_cast<const void*>(decode_sample_kernel), dim3(p.batch), dim3(kSampleThreads), 0,
                            stream, args, p.pdl != 0);
}

}  // namespace aries
