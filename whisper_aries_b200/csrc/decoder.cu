// Whisper text decoder + greedy generate on sm_100a: weight preparation, the cross-attention key/value projection of
// the encoder output, and the per-token step replayed as one CUDA graph (row f1).
//
// Replaces ctranslate2.models.Whisper.generate(encoder_output, prompts, beam_size=1, ...) -> layers::WhisperDecoder +
// GreedySearch, which the reference reaches through model.transcribe right after the encoder (ref:
// final_optimized_transcriber.py:326 with beam_size=1, temperature=0 at :432-441) [upstream, unverified offline].
//
//   once per batch of windows:  K|V_l = enc_out W_kv,l^T + b   (the encoder's tcgen05 GEMM, M = B*1500, N = 2d) x n_layers
//   per token (356 kernels for large-v3, 260 with LayerNorm folded into the GEMMs at <= 8 windows; programmatic
//   dependent launch, one graph):
//     embed -> n_layers x { LN -> QKV (skinny GEMM) -> self-attention over the cache (+append) -> O (+residual)
//                           LN -> Q  (skinny GEMM) -> cross-attention over K|V_l            -> O (+residual)
//                           LN -> fc1 (+GELU)      -> fc2 (+residual) }
//     -> LN -> logits (skinny GEMM against the tied embedding) -> rules + argmax (device) -> token buffer
// The host only replays the graph and polls a "sequences finished" counter every 16 tokens.
// bf16 weights / activations, f32 accumulation, f16 residual stream -- the same numeric contract as the encoder.
#include <cuda_bf16.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "decoder.h"
#include "gemm.h"
#include "skinny.h"

namespace aries {

namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline unsigned short f32_to_bf16(float f) {
    unsigned u;
    std::memcpy(&u, &f, 4);
    if ((u & 0x7F800000u) == 0x7F800000u && (u & 0x7FFFFFu)) return (unsigned short)((u >> 16) | 0x40);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (unsigned short)(u >> 16);
}

// f32 -> f16 bits, round to nearest even (host; the folded LayerNorm weights gamma * W)
inline unsigned short f32_to_f16(float f) {
    unsigned u;
    std::memcpy(&u, &f, 4);
    const unsigned sign = (u >> 16) & 0x8000u;
    const int exp = (int)((u >> 23) & 0xFF) - 127 + 15;
    unsigned man = u & 0x7FFFFFu;
    if (((u >> 23) & 0xFF) == 0xFF) return (unsigned short)(sign | 0x7C00u | (man ? 0x200u : 0));
    if (exp >= 31) return (unsigned short)(sign | 0x7C00u);
    if (exp <= 0) {
        if (exp < -10) return (unsigned short)sign;
        man |= 0x800000u;
        const int shift = 14 - exp;
        unsigned h = man >> shift;
        const unsigned rem = man & ((1u << shift) - 1), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (h & 1))) ++h;
        return (unsigned short)(sign | h);
    }
    unsigned h = ((unsigned)exp << 10) | (man >> 13);
    const unsigned rem = man & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) ++h;
    return (unsigned short)(sign | h);
}
inline float f16_to_f32(unsigned short h) {
    const unsigned sign = (unsigned)(h & 0x8000u) << 16;
    int exp = (h >> 10) & 0x1F;
    unsigned man = h & 0x3FFu;
    unsigned u;
    if (exp == 0) {
        if (man == 0) { u = sign; }
        else {
            exp = 1;
            while (!(man & 0x400u)) { man <<= 1; --exp; }
            man &= 0x3FFu;
            u = sign | ((unsigned)(exp + 127 - 15) << 23) | (man << 13);
        }
    } else if (exp == 31) {
        u = sign | 0x7F800000u | (man << 13);
    } else {
        u = sign | ((unsigned)(exp + 127 - 15) << 23) | (man << 13);
    }
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

bool env_on(const char* name, bool dflt) {
    const char* e = getenv(name);
    return e ? (e[0] != '0') : dflt;
}

struct DecLayerW {
    const float *ln1_g, *ln1_b, *bqkv, *bo, *ln2_g, *ln2_b, *bq2, *bkv2, *bo2, *ln3_g, *ln3_b, *b1, *b2;
    CUtensorMap m_qkv, m_o, m_q2, m_kv2, m_o2, m_fc1, m_fc2;
    const void *wqkv, *wo, *wq2, *wo2, *w1, *w2;     // the same bf16 matrices as plain pointers (persistent stack kernel)
    // folded LayerNorm: f16(gamma * W) maps, c1[n] = sum_k f16(gamma_k W[n,k]), c2[n] = sum_k beta_k W[n,k] + b[n]
    CUtensorMap f_qkv, f_q2, f_fc1;
    const float *c1_qkv, *c2_qkv, *c1_q2, *c2_q2, *c1_fc1, *c2_fc1;
};

struct GraphKey {
    int batch, prompt_len, max_length, suppress_blank, blank_id, eot, no_speech, no_timestamps, timestamp_begin,
        max_initial, n_forced, want_argmax, pdl, fuse_ln, pdl_mask, stack, fold;
    bool operator==(const GraphKey& o) const { return std::memcmp(this, &o, sizeof(GraphKey)) == 0; }
};

cudaError_t map2d(CUtensorMap* m, const void* base, unsigned long long inner, unsigned long long rows, unsigned box_rows) {
    const unsigned long long dims[2] = {inner, rows};
    const unsigned long long strides[2] = {2, inner * 2};
    const unsigned box[2] = {64, box_rows};
    return make_tmap_bf16(m, base, 2, dims, strides, box);
}

}  // namespace

struct DecoderPlan {
    int device = 0, sm_count = 0;
    DecoderShapeC cfg{};
    int max_batch = 0, NB = 0, v_pad = 0;
    char* d_weights = nullptr;
    const void* emb = nullptr;
    const float *pos = nullptr, *lnf_g = nullptr, *lnf_b = nullptr;
    CUtensorMap m_proj{};
    std::vector<DecLayerW> layers;
    // activations of one step (rows padded to NB; the pad rows stay zero)
    char* d_act = nullptr;
    void *x = nullptr, *y = nullptr, *qkv = nullptr, *ctx = nullptr, *q2 = nullptr, *h = nullptr;
    float* logits = nullptr;
    CUtensorMap a_y{}, a_ctx{}, a_h{}, a_x{};
    float2* ln_stats = nullptr;                      // folded LayerNorm: [NB][16] partial (sum, sum of squares) of the stream
    bool fold_ok = false;
    // caches
    __nv_bfloat16 *kc = nullptr, *vc = nullptr;      // [layers][max_batch][n_text_ctx][d]
    __nv_bfloat16* xkv = nullptr;                    // [layers][batch * n_audio_ctx][2d], grown on demand
    size_t xkv_cap = 0;
    // decoding state
    char* d_state = nullptr;
    int *tokens = nullptr, *step = nullptr, *done = nullptr, *last_ts = nullptr, *n_done = nullptr, *sot_index = nullptr,
        *use_ts = nullptr, *forced = nullptr, *argmax = nullptr;
    unsigned *ticket = nullptr, *suppress_bits = nullptr, *stack_bar = nullptr;
    StackLayerW* d_stack_layers = nullptr;           // device copy of the per-layer pointers (persistent stack kernel)
    float* stack_part = nullptr;                     // attention partials of the stack kernel
    long long* stack_trace = nullptr;                // ARIES_STACK_TRACE diagnostics (managed memory)
    float *score = nullptr, *nsp = nullptr;
    int* h_ndone = nullptr;                          // pinned
    struct GraphEntry {
        GraphKey key;
        cudaGraphExec_t exec;
        int per_step;
    };
    std::vector<GraphEntry> graphs;                  // one step graph per (batch, prompt length, options); <= 8 kept
    cudaStream_t ds = nullptr;                       // decoding stream (graph capture needs a non-default stream)
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    float stats[5] = {0, 0, 0, 0, 0};
    std::string error;
};

const char* decoder_plan_error(const DecoderPlan* pl) { return pl->error.c_str(); }
const DecoderShapeC* decoder_plan_cfg(const DecoderPlan* pl) { return &pl->cfg; }
int decoder_plan_max_batch(const DecoderPlan* pl) { return pl->max_batch; }
void decoder_plan_last_stats(const DecoderPlan* pl, float out[5]) { std::memcpy(out, pl->stats, sizeof(pl->stats)); }

void decoder_plan_destroy(DecoderPlan* pl) {
    if (!pl) return;
    for (auto& g : pl->graphs) cudaGraphExecDestroy(g.exec);
    for (auto& e : pl->ev)
        if (e) cudaEventDestroy(e);
    if (pl->ds) cudaStreamDestroy(pl->ds);
    cudaFree(pl->d_weights);
    cudaFree(pl->d_act);
    cudaFree(pl->kc);
    cudaFree(pl->vc);
    cudaFree(pl->xkv);
    cudaFree(pl->d_state);
    cudaFree(pl->d_stack_layers);
    cudaFree(pl->stack_part);
    cudaFree(pl->stack_trace);
    cudaFree(pl->ln_stats);
    if (pl->h_ndone) cudaFreeHost(pl->h_ndone);
    delete pl;
}

cudaError_t decoder_plan_create(int device, int sm_count, const DecoderShapeC& cfg, const WeightView* weights,
                                int n_weights, int max_batch, DecoderPlan** out, std::string* why) {
    *out = nullptr;
    const int d = cfg.d_model, f = cfg.d_ffn, L = cfg.n_layers, V = cfg.vocab, C = cfg.n_text_ctx;
    if (d <= 0 || d % 64 != 0 || cfg.n_heads * 64 != d || f <= 0 || f % 64 != 0 || L <= 0 || V < 16 || C < 2 ||
        cfg.n_audio_ctx <= 0 || max_batch < 1 || max_batch > 128) {
        *why = "unsupported decoder shape (need d_model % 64 == 0, head_dim 64, d_ffn % 64 == 0, 1 <= max_batch <= 128)";
        return cudaErrorInvalidValue;
    }
    std::map<std::string, const WeightView*> by_name;
    for (int i = 0; i < n_weights; ++i) by_name[weights[i].name] = &weights[i];
    auto find = [&](const std::string& name, std::initializer_list<long long> shape, bool required) -> const float* {
        auto it = by_name.find(name);
        if (it == by_name.end()) {
            if (required && why->empty()) *why = "missing weight: " + name;
            return nullptr;
        }
        const WeightView* w = it->second;
        bool ok = w->ndim == (int)shape.size();
        int i = 0;
        for (long long s : shape) ok = ok && (i < w->ndim) && (w->shape[i++] == s);
        if (!ok || !w->data) {
            if (why->empty()) *why = "weight has the wrong shape: " + name;
            return nullptr;
        }
        return w->data;
    };

    DecoderPlan* pl = new DecoderPlan();
    pl->device = device;
    pl->sm_count = sm_count;
    pl->cfg = cfg;
    pl->max_batch = max_batch;
    pl->NB = (max_batch + 15) / 16 * 16;
    pl->v_pad = (V + 127) / 128 * 128;
    pl->layers.resize(L);

    std::vector<unsigned char> blob;
    auto reserve = [&](size_t bytes) {
        const size_t at = align_up(blob.size(), 256);
        blob.resize(at + bytes);
        return at;
    };
    bool ok = true;
    auto put_f32 = [&](const std::string& name, std::initializer_list<long long> shape, size_t n) -> size_t {
        const float* src = find(name, shape, true);
        if (!src) { ok = false; return 0; }
        const size_t at = reserve(n * 4);
        std::memcpy(blob.data() + at, src, n * 4);
        return at;
    };
    auto put_bf16 = [&](const std::string& name, std::initializer_list<long long> shape, size_t n) -> size_t {
        const float* src = find(name, shape, true);
        if (!src) { ok = false; return 0; }
        const size_t at = reserve(n * 2);
        unsigned short* dst = reinterpret_cast<unsigned short*>(blob.data() + at);
        for (size_t i = 0; i < n; ++i) dst[i] = f32_to_bf16(src[i]);
        return at;
    };

    struct Offs {
        size_t ln1_g, ln1_b, wqkv, bqkv, wo, bo, ln2_g, ln2_b, wq2, bq2, wkv2, bkv2, wo2, bo2, ln3_g, ln3_b, w1, b1, w2, b2;
        size_t f_qkv, c1_qkv, c2_qkv, f_q2, c1_q2, c2_q2, f_fc1, c1_fc1, c2_fc1;
    };
    // LayerNorm folded into the consuming Linear: W' = f16(gamma * W), c1 = row sums of W' (as rounded), c2 = W beta + b
    // (opt-in, ARIES_DECODE_FOLD_LN=1 when the handle is created: +0.94 GB of f16 copies and ~3 s of host time for large-v3)
    const bool want_fold = env_on("ARIES_DECODE_FOLD_LN", false);
    auto put_folded = [&](const std::string& wname, const std::string& bname, const std::string& gname, const std::string& bename,
                          long long n, long long k, size_t* o_w, size_t* o_c1, size_t* o_c2) {
        *o_w = *o_c1 = *o_c2 = 0;
        if (!want_fold) return;
        const float* w = find(wname, {n, k}, true);
        const float* b = find(bname, {n}, true);
        const float* g = find(gname, {k}, true);
        const float* be = find(bename, {k}, true);
        if (!w || !b || !g || !be) { ok = false; return; }
        *o_w = reserve((size_t)n * k * 2);
        *o_c1 = reserve((size_t)n * 4);
        *o_c2 = reserve((size_t)n * 4);
        unsigned short* dw = reinterpret_cast<unsigned short*>(blob.data() + *o_w);
        float* c1 = reinterpret_cast<float*>(blob.data() + *o_c1);
        float* c2 = reinterpret_cast<float*>(blob.data() + *o_c2);
        for (long long r = 0; r < n; ++r) {
            double s1 = 0.0, s2 = 0.0;
            for (long long c = 0; c < k; ++c) {
                // the model's weight IS the bf16-rounded value (what every other path multiplies by)
                const unsigned wb = (unsigned)f32_to_bf16(w[r * k + c]) << 16;
                float wv;
                std::memcpy(&wv, &wb, 4);
                const unsigned short h = f32_to_f16(g[c] * wv);
                dw[r * k + c] = h;
                s1 += (double)f16_to_f32(h);
                s2 += (double)be[c] * (double)wv;
            }
            c1[r] = (float)s1;
            c2[r] = (float)(s2 + (double)b[r]);
        }
    };
    std::vector<Offs> offs(L);
    const size_t o_emb = put_bf16("decoder/embeddings/weight", {V, d}, (size_t)V * d);
    size_t o_proj = o_emb;
    if (by_name.count("decoder/projection/weight")) o_proj = put_bf16("decoder/projection/weight", {V, d}, (size_t)V * d);
    const size_t o_pos = put_f32("decoder/position_encodings/encodings", {C, d}, (size_t)C * d);
    const size_t o_lnf_g = put_f32("decoder/layer_norm/gamma", {d}, d);
    const size_t o_lnf_b = put_f32("decoder/layer_norm/beta", {d}, d);
    for (int i = 0; i < L && ok; ++i) {
        const std::string p = "decoder/layer_" + std::to_string(i);
        Offs& o = offs[i];
        o.ln1_g = put_f32(p + "/self_attention/layer_norm/gamma", {d}, d);
        o.ln1_b = put_f32(p + "/self_attention/layer_norm/beta", {d}, d);
        o.wqkv = put_bf16(p + "/self_attention/linear_0/weight", {3 * d, d}, (size_t)3 * d * d);
        o.bqkv = put_f32(p + "/self_attention/linear_0/bias", {3 * d}, (size_t)3 * d);
        o.wo = put_bf16(p + "/self_attention/linear_1/weight", {d, d}, (size_t)d * d);
        o.bo = put_f32(p + "/self_attention/linear_1/bias", {d}, d);
        o.ln2_g = put_f32(p + "/attention/layer_norm/gamma", {d}, d);
        o.ln2_b = put_f32(p + "/attention/layer_norm/beta", {d}, d);
        o.wq2 = put_bf16(p + "/attention/linear_0/weight", {d, d}, (size_t)d * d);
        o.bq2 = put_f32(p + "/attention/linear_0/bias", {d}, d);
        o.wkv2 = put_bf16(p + "/attention/linear_1/weight", {2 * d, d}, (size_t)2 * d * d);
        o.bkv2 = put_f32(p + "/attention/linear_1/bias", {2 * d}, (size_t)2 * d);
        o.wo2 = put_bf16(p + "/attention/linear_2/weight", {d, d}, (size_t)d * d);
        o.bo2 = put_f32(p + "/attention/linear_2/bias", {d}, d);
        o.ln3_g = put_f32(p + "/ffn/layer_norm/gamma", {d}, d);
        o.ln3_b = put_f32(p + "/ffn/layer_norm/beta", {d}, d);
        o.w1 = put_bf16(p + "/ffn/linear_0/weight", {f, d}, (size_t)f * d);
        o.b1 = put_f32(p + "/ffn/linear_0/bias", {f}, f);
        o.w2 = put_bf16(p + "/ffn/linear_1/weight", {d, f}, (size_t)d * f);
        o.b2 = put_f32(p + "/ffn/linear_1/bias", {d}, d);
        put_folded(p + "/self_attention/linear_0/weight", p + "/self_attention/linear_0/bias", p + "/self_attention/layer_norm/gamma",
                   p + "/self_attention/layer_norm/beta", 3LL * d, d, &o.f_qkv, &o.c1_qkv, &o.c2_qkv);
        put_folded(p + "/attention/linear_0/weight", p + "/attention/linear_0/bias", p + "/attention/layer_norm/gamma",
                   p + "/attention/layer_norm/beta", d, d, &o.f_q2, &o.c1_q2, &o.c2_q2);
        put_folded(p + "/ffn/linear_0/weight", p + "/ffn/linear_0/bias", p + "/ffn/layer_norm/gamma", p + "/ffn/layer_norm/beta", f, d,
                   &o.f_fc1, &o.c1_fc1, &o.c2_fc1);
    }
    if (!ok) {
        delete pl;
        return cudaErrorInvalidValue;
    }

    cudaError_t e;
    auto fail = [&](cudaError_t err) {
        *why = cudaGetErrorString(err);
        decoder_plan_destroy(pl);
        return err;
    };
    if ((e = cudaMalloc(&pl->d_weights, blob.size())) != cudaSuccess) return fail(e);
    if ((e = cudaMemcpy(pl->d_weights, blob.data(), blob.size(), cudaMemcpyHostToDevice)) != cudaSuccess) return fail(e);
    if ((e = gemm_init_device()) != cudaSuccess) return fail(e);
    if ((e = skinny_init_device()) != cudaSuccess) return fail(e);

    char* base = pl->d_weights;
    auto F = [&](size_t off) { return reinterpret_cast<const float*>(base + off); };
    pl->emb = base + o_emb;
    pl->pos = F(o_pos);
    pl->lnf_g = F(o_lnf_g);
    pl->lnf_b = F(o_lnf_b);
    if ((e = map2d(&pl->m_proj, base + o_proj, d, V, 128)) != cudaSuccess) return fail(e);
    for (int i = 0; i < L; ++i) {
        DecLayerW& lw = pl->layers[i];
        const Offs& o = offs[i];
        lw.ln1_g = F(o.ln1_g); lw.ln1_b = F(o.ln1_b); lw.bqkv = F(o.bqkv); lw.bo = F(o.bo);
        lw.ln2_g = F(o.ln2_g); lw.ln2_b = F(o.ln2_b); lw.bq2 = F(o.bq2); lw.bkv2 = F(o.bkv2); lw.bo2 = F(o.bo2);
        lw.ln3_g = F(o.ln3_g); lw.ln3_b = F(o.ln3_b); lw.b1 = F(o.b1); lw.b2 = F(o.b2);
        lw.wqkv = base + o.wqkv; lw.wo = base + o.wo; lw.wq2 = base + o.wq2; lw.wo2 = base + o.wo2;
        lw.w1 = base + o.w1; lw.w2 = base + o.w2;
        if ((e = map2d(&lw.m_qkv, base + o.wqkv, d, 3ull * d, 128)) != cudaSuccess) return fail(e);
        if ((e = map2d(&lw.m_o, base + o.wo, d, d, 128)) != cudaSuccess) return fail(e);
        if ((e = map2d(&lw.m_q2, base + o.wq2, d, d, 128)) != cudaSuccess) return fail(e);
        if ((e = map2d(&lw.m_kv2, base + o.wkv2, d, 2ull * d, gemm_b_box_rows())) != cudaSuccess) return fail(e);
        if ((e = map2d(&lw.m_o2, base + o.wo2, d, d, 128)) != cudaSuccess) return fail(e);
        if ((e = map2d(&lw.m_fc1, base + o.w1, d, f, 128)) != cudaSuccess) return fail(e);
        if ((e = map2d(&lw.m_fc2, base + o.w2, f, d, 128)) != cudaSuccess) return fail(e);
        lw.c1_qkv = F(o.c1_qkv); lw.c2_qkv = F(o.c2_qkv); lw.c1_q2 = F(o.c1_q2); lw.c2_q2 = F(o.c2_q2);
        lw.c1_fc1 = F(o.c1_fc1); lw.c2_fc1 = F(o.c2_fc1);
        if (want_fold) {
            if ((e = map2d(&lw.f_qkv, base + o.f_qkv, d, 3ull * d, 128)) != cudaSuccess) return fail(e);
            if ((e = map2d(&lw.f_q2, base + o.f_q2, d, d, 128)) != cudaSuccess) return fail(e);
            if ((e = map2d(&lw.f_fc1, base + o.f_fc1, d, f, 128)) != cudaSuccess) return fail(e);
        }
    }

    // ---- step activations
    const size_t NB = pl->NB;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        const size_t at = off;
        off = align_up(off + bytes, 1024);
        return at;
    };
    const size_t a_x = take(NB * d * 2), a_y = take(NB * d * 2), a_qkv = take(NB * 3 * d * 2), a_ctx = take(NB * d * 2),
                 a_q2 = take(NB * d * 2), a_h = take(NB * f * 2), a_lg = take((size_t)max_batch * pl->v_pad * 4);
    if ((e = cudaMalloc(&pl->d_act, off)) != cudaSuccess) return fail(e);
    if ((e = cudaMemset(pl->d_act, 0, off)) != cudaSuccess) return fail(e);
    pl->x = pl->d_act + a_x; pl->y = pl->d_act + a_y; pl->qkv = pl->d_act + a_qkv; pl->ctx = pl->d_act + a_ctx;
    pl->q2 = pl->d_act + a_q2; pl->h = pl->d_act + a_h; pl->logits = reinterpret_cast<float*>(pl->d_act + a_lg);
    if ((e = map2d(&pl->a_y, pl->y, d, NB, (unsigned)NB)) != cudaSuccess) return fail(e);
    if ((e = map2d(&pl->a_ctx, pl->ctx, d, NB, (unsigned)NB)) != cudaSuccess) return fail(e);
    if ((e = map2d(&pl->a_h, pl->h, f, NB, (unsigned)NB)) != cudaSuccess) return fail(e);
    if ((e = map2d(&pl->a_x, pl->x, d, NB, (unsigned)NB)) != cudaSuccess) return fail(e);      // the f16 stream as an operand
    if ((e = cudaMalloc(&pl->ln_stats, NB * 16 * sizeof(float2))) != cudaSuccess) return fail(e);
    if ((e = cudaMemset(pl->ln_stats, 0, NB * 16 * sizeof(float2))) != cudaSuccess) return fail(e);
    // the fold needs the residual GEMMs' cluster reduction path (splits > 1) and <= 16 feature tiles of 128 per row
    pl->fold_ok = want_fold && (d + 127) / 128 <= 16 && skinny_pick_splits(d, d, sm_count) > 1 &&
                  skinny_pick_splits(d, f, sm_count) > 1;

    const size_t cache_bytes = (size_t)L * max_batch * C * d * 2;
    if ((e = cudaMalloc(&pl->kc, cache_bytes)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc(&pl->vc, cache_bytes)) != cudaSuccess) return fail(e);

    // ---- decoding state
    off = 0;
    const size_t B = max_batch;
    const size_t s_tok = take(B * C * 4), s_forced = take(B * C * 4), s_argmax = take(B * C * 4), s_step = take(4),
                 s_done = take(B * 4), s_lts = take(B * 4), s_ndone = take(4), s_sot = take(B * 4), s_uts = take(B * 4),
                 s_ticket = take(4), s_sbar = take(4), s_bits = take(((size_t)V + 31) / 32 * 4),
                 s_score = take(B * 4), s_nsp = take(B * 4);
    if ((e = cudaMalloc(&pl->d_state, off)) != cudaSuccess) return fail(e);
    if ((e = cudaMemset(pl->d_state, 0, off)) != cudaSuccess) return fail(e);
    char* s = pl->d_state;
    pl->tokens = (int*)(s + s_tok); pl->forced = (int*)(s + s_forced); pl->argmax = (int*)(s + s_argmax);
    pl->step = (int*)(s + s_step); pl->done = (int*)(s + s_done); pl->last_ts = (int*)(s + s_lts);
    pl->n_done = (int*)(s + s_ndone); pl->sot_index = (int*)(s + s_sot); pl->use_ts = (int*)(s + s_uts);
    pl->stack_bar = (unsigned*)(s + s_sbar);
    pl->ticket = (unsigned*)(s + s_ticket); pl->suppress_bits = (unsigned*)(s + s_bits);
    pl->score = (float*)(s + s_score); pl->nsp = (float*)(s + s_nsp);
    {
        std::vector<StackLayerW> sl(L);
        for (int i = 0; i < L; ++i) {
            const DecLayerW& lw = pl->layers[i];
            sl[i] = StackLayerW{lw.wqkv, lw.wo, lw.wq2, lw.wo2, lw.w1, lw.w2, lw.ln1_g, lw.ln1_b, lw.bqkv, lw.bo, lw.ln2_g,
                                lw.ln2_b, lw.bq2, lw.bo2, lw.ln3_g, lw.ln3_b, lw.b1, lw.b2};
        }
        if ((e = cudaMalloc(&pl->d_stack_layers, L * sizeof(StackLayerW))) != cudaSuccess) return fail(e);
        if ((e = cudaMemcpy(pl->d_stack_layers, sl.data(), L * sizeof(StackLayerW), cudaMemcpyHostToDevice)) != cudaSuccess)
            return fail(e);
        const size_t pf = decode_stack_part_floats(cfg.n_heads) * 4;
        if ((e = cudaMalloc(&pl->stack_part, pf)) != cudaSuccess) return fail(e);
        if ((e = cudaMemset(pl->stack_part, 0, pf)) != cudaSuccess) return fail(e);
    }
    if (getenv("ARIES_STACK_TRACE") && (e = cudaMallocManaged(&pl->stack_trace, 4096 * 8)) != cudaSuccess) return fail(e);
    if ((e = cudaMallocHost(&pl->h_ndone, 4)) != cudaSuccess) return fail(e);
    for (auto& ev : pl->ev)
        if ((e = cudaEventCreate(&ev)) != cudaSuccess) return fail(e);
    if ((e = cudaStreamCreateWithFlags(&pl->ds, cudaStreamNonBlocking)) != cudaSuccess) return fail(e);
    *out = pl;
    return cudaSuccess;
}

#define ARIES_TRY(expr, what)                                                         \
    do {                                                                              \
        cudaError_t _e = (expr);                                                      \
        if (_e != cudaSuccess) {                                                      \
            pl->error = std::string(what) + ": " + cudaGetErrorString(_e);            \
            return _e;                                                                \
        }                                                                             \
    } while (0)

namespace {

// One decode step: consumes tokens[:, *step], writes tokens[:, *step + 1], increments *step.
cudaError_t run_step(DecoderPlan* pl, int batch, const SampleParams& sp, int xsplits, bool pdl, bool fuse_ln, bool stack,
                     bool fold, cudaStream_t stream, int* launches) {
    const auto& c = pl->cfg;
    const int d = c.d_model, f = c.d_ffn, C = c.n_text_ctx, A = c.n_audio_ctx;
    int n = 0;
    // Which kernel classes take a programmatic edge: 1 = skinny GEMMs, 2 = attention, 4 = the rest.  Measured per step
    // (tests/gpu_diag_decode.py pdlmask): 64 windows 4.74 ms with none, 4.50 with 5, 5.72 with 7 -- an attention grid
    // (1280 CTAs) launched early only holds SM slots while it waits; 16 windows 2.99 / 2.75 / 2.77.
    // With 1-2 windows every grid is tiny and 7 wins by 3 % (1.55 vs 1.59 ms).  Default: 7 for <= 2 windows, else 5.
    const char* mask_env = getenv("ARIES_DECODE_PDL_MASK");
    const int mask = mask_env ? atoi(mask_env) : (batch <= 2 ? 7 : 5);
    const bool pdl_g = pdl && (mask & 1), pdl_a = pdl && (mask & 2), pdl_o = pdl && (mask & 4);
    // LayerNorm folded into the consuming GEMM's operand load (<= 8 sequences): 3 launches fewer per layer
    auto ln_skinny = [&](int epi, const CUtensorMap& w, int N, const float* g, const float* b, const float* bias, void* out,
                         int ldo) {
        SkinnyParams q{};
        q.B = batch; q.NB = 16; q.N = N; q.K = d;
        q.splits = skinny_pick_splits_ln(N, d, pl->sm_count);
        q.bias = bias; q.out = out; q.ldo = ldo; q.pdl = pdl_g;
        q.ln_x = pl->x; q.ln_gamma = g; q.ln_beta = b;
        ++n;
        return skinny_launch(epi, w, w, q, stream);
    };
    auto skinny = [&](int epi, const CUtensorMap& w, const CUtensorMap& xin, int N, int K, const float* bias, void* out,
                      int ldo) {
        SkinnyParams q{};
        q.B = batch; q.NB = pl->NB; q.N = N; q.K = K;
        q.splits = skinny_pick_splits(N, K, pl->sm_count);
        q.bias = bias; q.out = out; q.ldo = ldo; q.pdl = pdl_g;
        // folded LayerNorm: every residual GEMM leaves the partial row sums of the stream it has just written
        if (fold && epi == SK_BIAS_RESID_F16) q.stats_out = pl->ln_stats;
        ++n;
        return skinny_launch(epi, w, xin, q, stream);
    };
    // LayerNorm folded into the GEMM (any batch size): the f16 stream is the operand, W' = f16(gamma W), the epilogue
    // applies mean / rstd from the partial sums (skinny.h SK_LNF_*)
    const int stat_parts = (d + 127) / 128;
    auto fold_skinny = [&](int epi, const CUtensorMap& w, int N, const float* c1, const float* c2, void* out, int ldo) {
        SkinnyParams q{};
        q.B = batch; q.NB = pl->NB; q.N = N; q.K = d;
        q.splits = skinny_pick_splits(N, d, pl->sm_count);
        q.bias = c2; q.c1 = c1; q.stats_in = pl->ln_stats; q.stats_parts = stat_parts; q.ln_dim = d;
        q.out = out; q.ldo = ldo; q.pdl = pdl_g;
        ++n;
        return skinny_launch(epi, w, pl->a_x, q, stream);
    };
    if (stack) {
        // <= 8 sequences: embedding, every layer and the final LayerNorm in one persistent cooperative kernel
        DecStackParams q{};
        q.layers = pl->d_stack_layers; q.n_layers = c.n_layers; q.d = d; q.f = f; q.heads = c.n_heads; q.batch = batch;
        q.max_batch = pl->max_batch; q.C = C; q.A = A;
        q.tokens = pl->tokens; q.tokens_ld = C; q.step = pl->step; q.done = pl->done; q.emb = pl->emb; q.pos = pl->pos;
        q.x = pl->x; q.q = pl->qkv; q.q2 = pl->q2; q.h = pl->h; q.y = pl->y; q.ctx = pl->ctx; q.kc = pl->kc; q.vc = pl->vc; q.xkv = pl->xkv;
        q.part = pl->stack_part; q.bar = pl->stack_bar; q.lnf_g = pl->lnf_g; q.lnf_b = pl->lnf_b;
        if (const char* tr = getenv("ARIES_STACK_TRACE")) {      // diagnostics: per-phase clock64 stamps of one CTA
            q.trace = pl->stack_trace;                           // allocated at plan creation (not during capture)
            q.trace_cta = atoi(tr);
        }
        ARIES_TRY(decode_stack_launch(q, pl->sm_count, stream), "decoder stack");
        ++n;
        ARIES_TRY(skinny(SK_LOGITS_F32, pl->m_proj, pl->a_y, c.vocab, d, nullptr, pl->logits, pl->v_pad), "logits");
        SampleParams sp2 = sp;
        sp2.pdl = pdl_o;
        sp2.stack_bar = pl->stack_bar;
        ARIES_TRY(decode_sample_launch(sp2, stream), "sampling");
        ++n;
        *launches = n;
        return cudaSuccess;
    }
    ARIES_TRY(decode_embed_launch(pl->tokens, C, pl->step, pl->emb, pl->pos, pl->x, batch, d, pdl_o, stream,
                                  fold ? pl->ln_stats : nullptr, stat_parts), "embed");
    ++n;
    const __nv_bfloat16* qkv = reinterpret_cast<const __nv_bfloat16*>(pl->qkv);
    for (int l = 0; l < c.n_layers; ++l) {
        const DecLayerW& lw = pl->layers[l];
        if (fold) {
            ARIES_TRY(fold_skinny(SK_LNF_BF16, lw.f_qkv, 3 * d, lw.c1_qkv, lw.c2_qkv, pl->qkv, 3 * d), "folded LN + qkv projection");
        } else if (fuse_ln) {
            ARIES_TRY(ln_skinny(SK_BIAS_BF16, lw.m_qkv, 3 * d, lw.ln1_g, lw.ln1_b, lw.bqkv, pl->qkv, 3 * d), "LN + qkv projection");
        } else {
            ARIES_TRY(decode_layernorm_launch(pl->x, lw.ln1_g, lw.ln1_b, pl->y, batch, d, pdl_o, stream), "layer norm 1");
            ARIES_TRY(skinny(SK_BIAS_BF16, lw.m_qkv, pl->a_y, 3 * d, d, lw.bqkv, pl->qkv, 3 * d), "qkv projection");
            ++n;
        }
        DecAttnParams a{};
        a.batch = batch; a.heads = c.n_heads;
        a.q = qkv; a.q_ld = 3 * d;
        a.k = pl->kc + (size_t)l * pl->max_batch * C * d;
        a.v = pl->vc + (size_t)l * pl->max_batch * C * d;
        a.kv_rows = C; a.kv_ld = d;
        a.new_k = qkv + d; a.new_v = qkv + 2 * d; a.new_ld = 3 * d;
        a.step = pl->step; a.n_keys_fixed = 0;
        a.out = pl->ctx; a.out_ld = d; a.splits = 1; a.pdl = pdl_a; a.done = pl->done;
        ARIES_TRY(decode_attention_launch(a, stream), "self-attention");
        ARIES_TRY(skinny(SK_BIAS_RESID_F16, lw.m_o, pl->a_ctx, d, d, lw.bo, pl->x, d), "self-attention output");
        if (fold) {
            ARIES_TRY(fold_skinny(SK_LNF_BF16, lw.f_q2, d, lw.c1_q2, lw.c2_q2, pl->q2, d), "folded LN + cross-attention query");
        } else if (fuse_ln) {
            ARIES_TRY(ln_skinny(SK_BIAS_BF16, lw.m_q2, d, lw.ln2_g, lw.ln2_b, lw.bq2, pl->q2, d), "LN + cross-attention query");
        } else {
            ARIES_TRY(decode_layernorm_launch(pl->x, lw.ln2_g, lw.ln2_b, pl->y, batch, d, pdl_o, stream), "layer norm 2");
            ARIES_TRY(skinny(SK_BIAS_BF16, lw.m_q2, pl->a_y, d, d, lw.bq2, pl->q2, d), "cross-attention query");
            ++n;
        }
        DecAttnParams x{};
        x.batch = batch; x.heads = c.n_heads;
        x.q = pl->q2; x.q_ld = d;
        __nv_bfloat16* kv = pl->xkv + (size_t)l * batch * A * 2 * d;
        x.k = kv; x.v = kv + d; x.kv_rows = A; x.kv_ld = 2 * d;
        x.n_keys_fixed = A;
        x.out = pl->ctx; x.out_ld = d; x.splits = xsplits;
        x.pdl = pdl_a; x.done = pl->done;
        ARIES_TRY(decode_attention_launch(x, stream), "cross-attention");
        ARIES_TRY(skinny(SK_BIAS_RESID_F16, lw.m_o2, pl->a_ctx, d, d, lw.bo2, pl->x, d), "cross-attention output");
        if (fold) {
            ARIES_TRY(fold_skinny(SK_LNF_GELU_BF16, lw.f_fc1, f, lw.c1_fc1, lw.c2_fc1, pl->h, f), "folded LN + fc1");
        } else if (fuse_ln) {
            ARIES_TRY(ln_skinny(SK_BIAS_GELU_BF16, lw.m_fc1, f, lw.ln3_g, lw.ln3_b, lw.b1, pl->h, f), "LN + fc1");
        } else {
            ARIES_TRY(decode_layernorm_launch(pl->x, lw.ln3_g, lw.ln3_b, pl->y, batch, d, pdl_o, stream), "layer norm 3");
            ARIES_TRY(skinny(SK_BIAS_GELU_BF16, lw.m_fc1, pl->a_y, f, d, lw.b1, pl->h, f), "fc1");
            ++n;
        }
        ARIES_TRY(skinny(SK_BIAS_RESID_F16, lw.m_fc2, pl->a_h, d, f, lw.b2, pl->x, d), "fc2");
        n += 2;
    }
    ARIES_TRY(decode_layernorm_launch(pl->x, pl->lnf_g, pl->lnf_b, pl->y, batch, d, pdl_o, stream), "final layer norm");
    ARIES_TRY(skinny(SK_LOGITS_F32, pl->m_proj, pl->a_y, c.vocab, d, nullptr, pl->logits, pl->v_pad), "logits");
    SampleParams sp2 = sp;
    sp2.pdl = pdl_o;
    ARIES_TRY(decode_sample_launch(sp2, stream), "sampling");
    n += 2;                       // (final LayerNorm + sampling; the skinny lambdas count themselves)
    *launches = n;
    return cudaSuccess;
}

}  // namespace

cudaError_t decoder_generate(DecoderPlan* pl, const void* enc_out, int batch, const int* prompts, int prompt_len,
                             const GenerateOptsC& o, int* tokens_out, int* lengths, float* scores, float* no_speech_prob,
                             cudaStream_t caller_stream) {
    const auto& c = pl->cfg;
    const int d = c.d_model, C = c.n_text_ctx, A = c.n_audio_ctx, V = c.vocab, L = c.n_layers;
    // everything runs on the plan's own stream, ordered after the caller's (the encoder output may still be in flight)
    cudaStream_t stream = pl->ds;
    ARIES_TRY(cudaEventRecord(pl->ev[3], caller_stream), "event");
    ARIES_TRY(cudaStreamWaitEvent(stream, pl->ev[3], 0), "stream wait");
    auto bad_id = [&](int id) { return id < 0 || id >= V; };
    if (!enc_out || !prompts || !tokens_out || batch < 1 || batch > pl->max_batch || prompt_len < 1 ||
        o.max_length > C || prompt_len >= o.max_length || bad_id(o.eot) || bad_id(o.no_speech) ||
        bad_id(o.timestamp_begin) || bad_id(o.blank_id) || bad_id(o.no_timestamps) || o.n_suppress < 0 || o.n_forced < 0 ||
        o.n_forced > C || (o.n_suppress && !o.suppress_tokens)) {
        pl->error = "invalid generate arguments (batch <= max_batch, 1 <= prompt_len < max_length <= n_text_ctx, token ids "
                    "inside the vocabulary)";
        return cudaErrorInvalidValue;
    }
    for (int i = 0; i < batch * prompt_len; ++i)
        if (bad_id(prompts[i])) {
            pl->error = "prompt token id outside the vocabulary";
            return cudaErrorInvalidValue;
        }
    // ---- cross-attention keys / values of every layer (once per batch of windows)
    const size_t xkv_bytes = (size_t)L * batch * A * 2 * d * 2;
    if (xkv_bytes > pl->xkv_cap) {
        // the cached step graphs have the old cache's address baked in: they go with it
        for (auto& g : pl->graphs) cudaGraphExecDestroy(g.exec);
        pl->graphs.clear();
        cudaFree(pl->xkv);
        pl->xkv = nullptr;
        pl->xkv_cap = 0;
        ARIES_TRY(cudaMalloc(&pl->xkv, xkv_bytes), "cudaMalloc (cross-attention cache)");
        pl->xkv_cap = xkv_bytes;
    }
    // ---- host-side staging of the per-sequence state
    std::vector<int> h_tok((size_t)batch * C, o.eot), h_sot(batch, 0), h_uts(batch, 1), h_lts(batch, -1);
    for (int b = 0; b < batch; ++b)
        for (int i = 0; i < prompt_len; ++i) {
            const int t = prompts[(size_t)b * prompt_len + i];
            h_tok[(size_t)b * C + i] = t;
            if (t == o.sot) h_sot[b] = i;
            if (t == o.no_timestamps) h_uts[b] = 0;
        }
    std::vector<unsigned> h_bits(((size_t)V + 31) / 32, 0u);
    for (int i = 0; i < o.n_suppress; ++i) {
        const int t = o.suppress_tokens[i];
        if (bad_id(t)) {
            pl->error = "suppress_tokens id outside the vocabulary";
            return cudaErrorInvalidValue;
        }
        h_bits[t >> 5] |= 1u << (t & 31);
    }
    // every argument is validated before the first enqueue: an early return must not leave copies from these stack-local
    // vectors in flight
    std::vector<int> h_forced;
    if (o.forced && o.n_forced > 0) {
        h_forced.assign(o.forced, o.forced + (size_t)batch * o.n_forced);
        for (int t : h_forced)
            if (bad_id(t)) {
                pl->error = "forced token id outside the vocabulary";
                return cudaErrorInvalidValue;
            }
    }
    ARIES_TRY(cudaMemcpyAsync(pl->tokens, h_tok.data(), h_tok.size() * 4, cudaMemcpyHostToDevice, stream), "upload tokens");
    ARIES_TRY(cudaMemcpyAsync(pl->sot_index, h_sot.data(), batch * 4, cudaMemcpyHostToDevice, stream), "upload state");
    ARIES_TRY(cudaMemcpyAsync(pl->use_ts, h_uts.data(), batch * 4, cudaMemcpyHostToDevice, stream), "upload state");
    ARIES_TRY(cudaMemcpyAsync(pl->last_ts, h_lts.data(), batch * 4, cudaMemcpyHostToDevice, stream), "upload state");
    ARIES_TRY(cudaMemcpyAsync(pl->suppress_bits, h_bits.data(), h_bits.size() * 4, cudaMemcpyHostToDevice, stream),
              "upload suppress mask");
    if (!h_forced.empty())
        ARIES_TRY(cudaMemcpyAsync(pl->forced, h_forced.data(), h_forced.size() * 4, cudaMemcpyHostToDevice, stream),
                  "upload forced tokens");
    ARIES_TRY(cudaMemsetAsync(pl->step, 0, 4, stream), "memset");
    ARIES_TRY(cudaMemsetAsync(pl->n_done, 0, 4, stream), "memset");
    ARIES_TRY(cudaMemsetAsync(pl->ticket, 0, 4, stream), "memset");
    ARIES_TRY(cudaMemsetAsync(pl->stack_bar, 0, 4, stream), "memset");
    ARIES_TRY(cudaMemsetAsync(pl->done, 0, batch * 4, stream), "memset");
    ARIES_TRY(cudaMemsetAsync(pl->score, 0, batch * 4, stream), "memset");
    ARIES_TRY(cudaMemsetAsync(pl->nsp, 0, batch * 4, stream), "memset");
    if (o.argmax_out) ARIES_TRY(cudaMemsetAsync(pl->argmax, 0xFF, (size_t)batch * C * 4, stream), "memset");
    // the host vectors above must outlive their asynchronous copies
    ARIES_TRY(cudaStreamSynchronize(stream), "synchronise (state upload)");

    ARIES_TRY(cudaEventRecord(pl->ev[0], stream), "event");
    CUtensorMap a_enc;
    ARIES_TRY(map2d(&a_enc, enc_out, d, (unsigned long long)batch * A, 128), "tensor map (encoder output)");
    int kv_launches = 0;
    for (int l = 0; l < L; ++l) {
        GemmParams g{};
        g.M = batch * A; g.N = 2 * d; g.K = d; g.a_cols = d;
        g.p_in = g.M; g.t_valid = g.M; g.p_out = g.M; g.row_off = 0; g.ldo = 2 * d;
        g.bias = pl->layers[l].bkv2;
        g.out = pl->xkv + (size_t)l * batch * A * 2 * d;
        ARIES_TRY(gemm_launch(EPI_BIAS_BF16, a_enc, pl->layers[l].m_kv2, g, pl->sm_count, stream), "cross-attention K|V projection");
        ++kv_launches;
    }
    ARIES_TRY(cudaEventRecord(pl->ev[1], stream), "event");

    // ---- the per-token step
    SampleParams sp{};
    sp.batch = batch; sp.vocab = V; sp.logits = pl->logits; sp.logits_ld = pl->v_pad;
    sp.tokens = pl->tokens; sp.tokens_ld = C; sp.prompt_len = prompt_len; sp.max_length = o.max_length;
    sp.step = pl->step; sp.suppress_bits = pl->suppress_bits;
    sp.suppress_blank = o.suppress_blank; sp.blank_id = o.blank_id;
    sp.eot = o.eot; sp.no_speech = o.no_speech; sp.no_timestamps = o.no_timestamps; sp.timestamp_begin = o.timestamp_begin;
    sp.max_initial_timestamp_index = o.max_initial_timestamp_index;
    sp.sot_index = pl->sot_index; sp.use_timestamps = pl->use_ts; sp.done = pl->done; sp.last_timestamp = pl->last_ts;
    sp.score = pl->score; sp.no_speech_prob = pl->nsp; sp.n_done = pl->n_done; sp.ticket = pl->ticket;
    sp.forced = (o.forced && o.n_forced > 0) ? pl->forced : nullptr;
    sp.forced_ld = o.n_forced; sp.n_forced = o.n_forced;
    sp.argmax_out = o.argmax_out ? pl->argmax : nullptr;

    // cross-attention: ~4 CTAs per SM in flight; the splits of a (sequence, head) form a cluster of <= 8 CTAs
    int want = (4 * pl->sm_count + batch * c.n_heads - 1) / (batch * c.n_heads);
    const int xsplits = want < 1 ? 1 : (want > 8 ? 8 : want);
    // fused LayerNorm: small batch only (two rows per warp), and only where every fused GEMM fits (d <= 1280)
    const bool fuse_ln = env_on("ARIES_DECODE_FUSED_LN", true) && batch <= 8 && d <= 1280 &&
                         skinny_pick_splits_ln(3 * d, d, pl->sm_count) > 0 && skinny_pick_splits_ln(d, d, pl->sm_count) > 0 &&
                         skinny_pick_splits_ln(c.d_ffn, d, pl->sm_count) > 0;
    // persistent stack kernel (<= 8 sequences): opt-in with ARIES_DECODE_STACK=1 -- measured SLOWER than the launch-per-op
    // step on B200 (1.69 vs 1.54 ms per token at 1 window, 3.4 vs 2.0 at 8; decode_stack.cu header)
    const bool stack = env_on("ARIES_DECODE_STACK", false) &&
                       decode_stack_supported(d, c.d_ffn, c.n_heads, batch, pl->sm_count);
    // LayerNorm folded into the GEMMs by algebra (every batch size).  Opt-in (ARIES_DECODE_FOLD_LN=1 at handle creation
    // AND at the call): measured neutral -- 1.60 / 1.95 / 4.45 ms per token at 1 / 8 / 64 windows against 1.57 / 2.01 /
    // 4.38 -- because the LayerNorm launches it removes were already hidden behind programmatic dependent launch
    const bool fold = env_on("ARIES_DECODE_FOLD_LN", false) && pl->fold_ok && !stack;
    const bool use_graph = env_on("ARIES_DECODE_GRAPH", true) && !o.logits_out;
    // programmatic dependent launch (GEMMs and the small kernels; see run_step): -5 .. -12 % per step at every batch size
    bool pdl = env_on("ARIES_DECODE_PDL", true);
    sp.pdl = pdl;
    int per_step = 0;

    cudaGraphExec_t graph = nullptr;
    if (use_graph) {
        GraphKey key{batch, prompt_len, o.max_length, o.suppress_blank, o.blank_id, o.eot, o.no_speech, o.no_timestamps,
                     o.timestamp_begin, o.max_initial_timestamp_index, o.n_forced, o.argmax_out ? 1 : 0, pdl ? 1 : 0,
                     fuse_ln ? 1 : 0, getenv("ARIES_DECODE_PDL_MASK") ? atoi(getenv("ARIES_DECODE_PDL_MASK")) : -1,
                     stack ? 1 : 0, fold ? 1 : 0};
        for (auto& g : pl->graphs)
            if (g.key == key) {
                graph = g.exec;
                per_step = g.per_step;
            }
        if (!graph) {
            for (int attempt = 0; attempt < 2 && !graph; ++attempt) {
                sp.pdl = pdl;
                cudaGraph_t g = nullptr;
                ARIES_TRY(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal), "begin capture");
                cudaError_t e1 = run_step(pl, batch, sp, xsplits, pdl, fuse_ln, stack, fold, stream, &per_step);
                cudaError_t e2 = cudaStreamEndCapture(stream, &g);
                cudaError_t e3 = (e1 == cudaSuccess && e2 == cudaSuccess) ? cudaGraphInstantiate(&graph, g, 0) : cudaErrorUnknown;
                if (g) cudaGraphDestroy(g);
                if (e3 != cudaSuccess) {
                    graph = nullptr;
                    cudaGetLastError();
                    if (!pdl) {
                        if (e1 == cudaSuccess) pl->error = std::string("graph capture of the decode step failed: ") +
                                                           cudaGetErrorString(e2 != cudaSuccess ? e2 : e3);
                        return e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3);
                    }
                    pdl = false;        // programmatic edges refused by this driver: capture again with plain edges
                }
            }
            // (the entry is filed under the REQUESTED key, so a fallback without programmatic edges is found again)
            if (pl->graphs.size() >= 8) {
                cudaGraphExecDestroy(pl->graphs.front().exec);
                pl->graphs.erase(pl->graphs.begin());
            }
            pl->graphs.push_back({key, graph, per_step});
        }
    }

    const int last_step = o.max_length - 2;      // consuming position max_length - 2 fills the last position
    int steps_run = 0;
    for (int t = 0; t <= last_step; ++t) {
        if (use_graph) {
            ARIES_TRY(cudaGraphLaunch(graph, stream), "graph launch");
        } else {
            cudaError_t e = run_step(pl, batch, sp, xsplits, pdl, fuse_ln, stack, fold, stream, &per_step);
            if (e != cudaSuccess) return e;
            if (o.logits_out)
                ARIES_TRY(cudaMemcpy2DAsync(o.logits_out + (size_t)t * batch * V, (size_t)V * 4, pl->logits,
                                            (size_t)pl->v_pad * 4, (size_t)V * 4, batch, cudaMemcpyDeviceToHost, stream),
                          "download logits");
        }
        ++steps_run;
        if (t + 1 >= prompt_len && ((t + 1 - prompt_len) % 16 == 15 || t == last_step)) {
            ARIES_TRY(cudaMemcpyAsync(pl->h_ndone, pl->n_done, 4, cudaMemcpyDeviceToHost, stream), "poll");
            ARIES_TRY(cudaStreamSynchronize(stream), "synchronise (decode loop)");
            if (*pl->h_ndone >= batch) break;
        }
    }
    ARIES_TRY(cudaEventRecord(pl->ev[2], stream), "event");

    // ---- results
    std::vector<int> h_out((size_t)batch * C);
    ARIES_TRY(cudaMemcpyAsync(h_out.data(), pl->tokens, h_out.size() * 4, cudaMemcpyDeviceToHost, stream), "download tokens");
    if (scores) ARIES_TRY(cudaMemcpyAsync(scores, pl->score, batch * 4, cudaMemcpyDeviceToHost, stream), "download scores");
    if (no_speech_prob)
        ARIES_TRY(cudaMemcpyAsync(no_speech_prob, pl->nsp, batch * 4, cudaMemcpyDeviceToHost, stream), "download");
    std::vector<int> h_arg;
    if (o.argmax_out) {
        h_arg.resize((size_t)batch * C);
        ARIES_TRY(cudaMemcpyAsync(h_arg.data(), pl->argmax, h_arg.size() * 4, cudaMemcpyDeviceToHost, stream), "download");
    }
    ARIES_TRY(cudaStreamSynchronize(stream), "synchronise (results)");
    const int filled = steps_run + 1 > prompt_len ? steps_run + 1 : prompt_len;   // positions holding a prompt / sampled token
    for (int b = 0; b < batch; ++b) {
        int n = 0;
        for (int i = 0; i < o.max_length; ++i) {
            const int t = (i < filled) ? h_out[(size_t)b * C + i] : o.eot;
            tokens_out[(size_t)b * o.max_length + i] = t;
            if (o.argmax_out) o.argmax_out[(size_t)b * o.max_length + i] = h_arg[(size_t)b * C + i];
        }
        for (int i = prompt_len; i < filled && i < o.max_length; ++i) {
            if (h_out[(size_t)b * C + i] == o.eot) break;
            ++n;
        }
        if (lengths) lengths[b] = n;
    }
    if (pl->stack_trace && getenv("ARIES_STACK_TRACE")) {
        // last step: 17 stamps per layer (see decode_stack_kernel): phase start, staged, weights done / barrier done
        const long long* t = pl->stack_trace;
        const int per_layer = 23;   // LN1 QKV sync | attn sync | ctx O sync | LN2 Q2 sync | xattn sync | ctx O2 sync | LN3 fc1 sync | h fc2 sync
        for (int l = 0; l < L && l < 3; ++l) {
            fprintf(stderr, "stack trace layer %d (cycles):", l);
            for (int i = 1; i < per_layer; ++i) fprintf(stderr, " %lld", t[l * per_layer + i] - t[l * per_layer + i - 1]);
            fprintf(stderr, "\n");
        }
        fprintf(stderr, "stack trace: whole stack %lld cycles\n", t[L * per_layer - 1] - t[0]);
    }
    float ms_kv = 0, ms_loop = 0;
    cudaEventElapsedTime(&ms_kv, pl->ev[0], pl->ev[1]);
    cudaEventElapsedTime(&ms_loop, pl->ev[1], pl->ev[2]);
    pl->stats[0] = ms_kv;
    pl->stats[1] = ms_loop;
    pl->stats[2] = (float)steps_run;
    pl->stats[3] = (float)per_step;
    pl->stats[4] = (float)kv_launches;
    return cudaSuccess;
}

cudaError_t decoder_detect_language(DecoderPlan* pl, const void* enc_out, int batch, const GenerateOptsC& ids,
                                    const int* lang_ids, int n_lang, float* probs, cudaStream_t stream) {
    const int V = pl->cfg.vocab;
    if (!lang_ids || !probs || n_lang < 1) {
        pl->error = "detect_language: no language ids";
        return cudaErrorInvalidValue;
    }
    for (int i = 0; i < n_lang; ++i)
        if (lang_ids[i] < 0 || lang_ids[i] >= V) {
            pl->error = "detect_language: language id outside the vocabulary";
            return cudaErrorInvalidValue;
        }
    // one step of the ordinary decode path on the prompt [<|startoftranscript|>] with the logits brought to the host
    GenerateOptsC o = ids;
    o.max_length = 2;
    o.suppress_blank = 0;
    o.suppress_tokens = nullptr;
    o.n_suppress = 0;
    o.forced = nullptr;
    o.n_forced = 0;
    o.argmax_out = nullptr;
    std::vector<float> logits((size_t)batch * V);
    o.logits_out = logits.data();
    std::vector<int> prompts(batch, ids.sot), toks((size_t)batch * 2);
    cudaError_t e = decoder_generate(pl, enc_out, batch, prompts.data(), 1, o, toks.data(), nullptr, nullptr, nullptr, stream);
    if (e != cudaSuccess) return e;
    for (int b = 0; b < batch; ++b) {
        const float* lg = logits.data() + (size_t)b * V;
        float m = lg[lang_ids[0]];
        for (int i = 1; i < n_lang; ++i) m = lg[lang_ids[i]] > m ? lg[lang_ids[i]] : m;
        double sum = 0.0;
        for (int i = 0; i < n_lang; ++i) sum += std::exp((double)(lg[lang_ids[i]] - m));
        for (int i = 0; i < n_lang; ++i) probs[(size_t)b * n_lang + i] = (float)(std::exp((double)(lg[lang_ids[i]] - m)) / sum);
    }
    return cudaSuccess;
}

}  // namespace aries
