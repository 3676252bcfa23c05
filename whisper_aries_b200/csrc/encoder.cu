// Whisper encoder forward on sm_100a: weight preparation and the per-call launch sequence.
//
// Replaces CTranslate2 4.6.0 layers::WhisperEncoder::operator() behind ctranslate2.models.Whisper.encode (SURVEY.md
// rows a-6..a-8), which the reference reaches through model.transcribe (ref: final_optimized_transcriber.py:326).
//
//   mel f32 [B, n_mels, frames] --(transpose, bf16)--> [B, 3002, c_pad]               (zero row before / after)
//   conv1 k3 s1 p1 + GELU   = implicit GEMM, K = 3 c_pad, A rows overlap in memory   -> c1 bf16 [B, 3002, d]
//   conv2 k3 s2 p1 + GELU + positions = implicit GEMM, K = 3 d over row PAIRS         -> x  f32  [B*1500, d]
//   n_layers x { QKV GEMM with LayerNorm folded in (V stored transposed) -> fused attention -> O GEMM (+bias +residual)
//                fc1 GEMM with LayerNorm folded in (+GELU)               -> fc2 GEMM (+bias +residual) }
//   final LN -> bf16 [B, 1500, d]
// LayerNorm never runs as a pass of its own inside the layers (north_star (2)): the epilogue that writes the residual
// stream (conv2, O, fc2) also leaves per-row (sum, sum of squares) partials, the consuming GEMM multiplies the f16
// stream itself by gamma-scaled f16 weights and applies  rstd (acc - mean c1) + c2  in its epilogue (gemm.h).  Only
// ln_post, whose output IS the result, is a LayerNorm kernel.
// bf16 operands, f32 accumulation in TMEM, f16 residual stream (what CTranslate2's float16 mode keeps; 8x finer
// than bf16, and half the HBM traffic of f32 for the two residual epilogues and LayerNorm), f32 LayerNorm statistics
// and softmax.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "attention.h"
#include "encoder.h"
#include "gemm.h"
#include "layernorm.h"
#include "logmel.h"
#include "profiler.h"

namespace aries {

namespace {

constexpr int kFramesIn = 3000;
constexpr int kRowsPadded = kFramesIn + 2;
constexpr float kLnEps = 1e-5f;

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline unsigned short f32_to_bf16(float f) {
    unsigned u;
    std::memcpy(&u, &f, 4);
    if ((u & 0x7F800000u) == 0x7F800000u && (u & 0x7FFFFFu)) return (unsigned short)((u >> 16) | 0x40);   // NaN
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (unsigned short)(u >> 16);
}

struct LayerW {
    // c1_* = row sums of the gamma-scaled f16 weights, c2_* = beta folded through the weights + bias (gemm.h, EPI_LN_*)
    const float *c1_qkv, *c2_qkv, *bo, *c1_fc1, *c2_fc1, *b2;
    CUtensorMap m_qkv, m_o, m_fc1, m_fc2;
};

}  // namespace

struct EncoderPlan {
    int device = 0;
    int sm_count = 0;
    EncoderShapeC cfg{};
    int c_pad = 0;
    int t_pad = 0;
    char* d_weights = nullptr;       // one allocation
    const float *conv1_b = nullptr, *conv2_b = nullptr, *pos = nullptr, *lnf_g = nullptr, *lnf_b = nullptr;
    CUtensorMap m_conv1{}, m_conv2{};
    std::vector<LayerW> layers;
    // activation tensor maps, cached per (workspace, batch)
    const void* cached_ws = nullptr;
    int cached_batch = 0;
    CUtensorMap a_mel{}, a_c1{}, a_x{}, a_ctx{}, a_h{};
    AttnMaps a_attn{};
    int last_launches = 0;
    Profiler prof;
    std::string error;
};

namespace {

struct WsLayout {
    size_t melT, x, stats, ctx, qk, vt, h, total;
};

WsLayout ws_layout(const EncoderPlan& pl, int batch) {
    const auto& c = pl.cfg;
    const size_t B = (size_t)batch, T = (size_t)c.n_ctx, d = (size_t)c.d_model;
    WsLayout w{};
    size_t off = 0;
    auto take = [&](size_t bytes) {
        const size_t at = off;
        off = align_up(off + bytes, 1024);
        return at;
    };
    w.melT = take(B * kRowsPadded * pl.c_pad * 2);
    w.x = take(B * T * d * 2);
    w.stats = take(B * T * (size_t)gemm_stats_parts((int)d) * sizeof(float2));
    w.ctx = take(B * T * d * 2);
    w.qk = take(B * T * 2 * d * 2);
    w.vt = take(B * d * pl.t_pad * 2);
    const size_t hb = B * T * (size_t)c.d_ffn * 2, c1 = B * kRowsPadded * d * 2;
    w.h = take(hb > c1 ? hb : c1);
    w.total = off;
    return w;
}

cudaError_t map2d(CUtensorMap* m, const void* base, unsigned long long inner, unsigned long long rows, unsigned box_rows) {
    const unsigned long long dims[2] = {inner, rows};
    const unsigned long long strides[2] = {2, inner * 2};
    const unsigned box[2] = {64, box_rows};
    return make_tmap_bf16(m, base, 2, dims, strides, box);
}

}  // namespace

const char* encoder_plan_error(const EncoderPlan* pl) { return pl->error.c_str(); }
int encoder_plan_last_launches(const EncoderPlan* pl) { return pl->last_launches; }
const EncoderShapeC* encoder_plan_cfg(const EncoderPlan* pl) { return &pl->cfg; }

size_t encoder_workspace_bytes(const EncoderPlan* pl, int batch) {
    if (batch <= 0) return 0;
    return ws_layout(*pl, batch).total;
}

void encoder_plan_destroy(EncoderPlan* pl) {
    if (!pl) return;
    cudaFree(pl->d_weights);
    delete pl;
}

cudaError_t encoder_plan_create(int device, int sm_count, const EncoderShapeC& cfg, const WeightView* weights,
                                int n_weights, EncoderPlan** out, std::string* why) {
    *out = nullptr;
    const int d = cfg.d_model, f = cfg.d_ffn, L = cfg.n_layers, T = cfg.n_ctx;
    if (d <= 0 || d % 128 != 0 || d > 2048 || cfg.n_heads * 64 != d || f % 128 != 0 || f <= 0 || L <= 0 ||
        cfg.n_mels <= 0 || cfg.n_mels > 256 || T != 1500) {
        *why = "unsupported encoder shape (need d_model % 128 == 0, d_model <= 2048, head_dim 64, d_ffn % 128 == 0, "
               "n_ctx 1500)";
        return cudaErrorInvalidValue;
    }
    if (gemm_stats_parts(d) > 16) {
        *why = "unsupported d_model: the LayerNorm hand-over between the GEMMs carries at most 16 column slices per row "
               "(d_model a multiple of 256 up to 2048, or of 128 up to 1024)";
        return cudaErrorInvalidValue;
    }
    std::map<std::string, const WeightView*> by_name;
    for (int i = 0; i < n_weights; ++i) by_name[weights[i].name] = &weights[i];
    auto need = [&](const std::string& name, std::initializer_list<long long> shape) -> const float* {
        auto it = by_name.find(name);
        if (it == by_name.end()) {
            if (why->empty()) *why = "missing weight: " + name;
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

    EncoderPlan* pl = new EncoderPlan();
    pl->device = device;
    pl->sm_count = sm_count;
    pl->cfg = cfg;
    pl->c_pad = (cfg.n_mels + 63) / 64 * 64;
    pl->t_pad = (T + 7) / 8 * 8;          // 1504: 16-byte row pitch for the V^T tensor map
    pl->layers.resize(L);
    const int cp = pl->c_pad;

    // ---- host staging: bf16 matrices and f32 vectors in one blob
    std::vector<unsigned char> blob;
    auto reserve = [&](size_t bytes) {
        const size_t at = align_up(blob.size(), 256);
        blob.resize(at + bytes);
        return at;
    };
    auto put_f32 = [&](const float* src, size_t n) {
        const size_t at = reserve(n * 4);
        std::memcpy(blob.data() + at, src, n * 4);
        return at;
    };
    auto put_bf16 = [&](const float* src, size_t n) {
        const size_t at = reserve(n * 2);
        unsigned short* dst = reinterpret_cast<unsigned short*>(blob.data() + at);
        for (size_t i = 0; i < n; ++i) dst[i] = f32_to_bf16(src[i]);
        return at;
    };
    // conv weight [out, in, 3] -> implicit-GEMM B [out, 3 * in_pad] with k = tap * in_pad + channel
    auto put_conv = [&](const float* src, int n_out, int n_in, int in_pad) {
        const size_t at = reserve((size_t)n_out * 3 * in_pad * 2);
        unsigned short* dst = reinterpret_cast<unsigned short*>(blob.data() + at);
        std::memset(dst, 0, (size_t)n_out * 3 * in_pad * 2);
        for (int o = 0; o < n_out; ++o)
            for (int c = 0; c < n_in; ++c)
                for (int tap = 0; tap < 3; ++tap)
                    dst[((size_t)o * 3 + tap) * in_pad + c] = f32_to_bf16(src[((size_t)o * n_in + c) * 3 + tap]);
        return at;
    };

    // LayerNorm fold (gemm.h): W' = gamma (.) W rounded to f16, c1[n] = sum_k W'[n,k] (of the ROUNDED values, so that the
    // mean cancels exactly), c2[n] = sum_k beta[k] W[n,k] + bias[n]; sums in f64.
    auto put_ln_folded = [&](const float* w, const float* bias, const float* gamma, const float* beta, int n_out, int n_in,
                             size_t* o_w, size_t* o_c1, size_t* o_c2) {
        *o_w = reserve((size_t)n_out * n_in * 2);
        *o_c1 = reserve((size_t)n_out * 4);
        *o_c2 = reserve((size_t)n_out * 4);
        __half* dst = reinterpret_cast<__half*>(blob.data() + *o_w);
        float* c1 = reinterpret_cast<float*>(blob.data() + *o_c1);
        float* c2 = reinterpret_cast<float*>(blob.data() + *o_c2);
        for (int n = 0; n < n_out; ++n) {
            double s1 = 0.0, s2 = 0.0;
            const float* row = w + (size_t)n * n_in;
            for (int k = 0; k < n_in; ++k) {
                const __half h = __float2half_rn(row[k] * gamma[k]);
                dst[(size_t)n * n_in + k] = h;
                s1 += (double)__half2float(h);
                s2 += (double)beta[k] * (double)row[k];
            }
            c1[n] = (float)s1;
            c2[n] = (float)(s2 + (double)bias[n]);
        }
    };
    struct Offs {
        size_t wqkv, c1_qkv, c2_qkv, wo, bo, w1, c1_fc1, c2_fc1, w2, b2;
    };
    std::vector<Offs> offs(L);
    size_t o_conv1w = 0, o_conv1b = 0, o_conv2w = 0, o_conv2b = 0, o_pos = 0, o_lnf_g = 0, o_lnf_b = 0;
    bool ok = true;
    {
        const float* w;
        ok &= (w = need("encoder/conv1/weight", {d, cfg.n_mels, 3})) != nullptr;
        if (w) o_conv1w = put_conv(w, d, cfg.n_mels, cp);
        ok &= (w = need("encoder/conv1/bias", {d})) != nullptr;
        if (w) o_conv1b = put_f32(w, d);
        ok &= (w = need("encoder/conv2/weight", {d, d, 3})) != nullptr;
        if (w) o_conv2w = put_conv(w, d, d, d);
        ok &= (w = need("encoder/conv2/bias", {d})) != nullptr;
        if (w) o_conv2b = put_f32(w, d);
        ok &= (w = need("encoder/position_encodings/encodings", {T, d})) != nullptr;
        if (w) o_pos = put_f32(w, (size_t)T * d);
        ok &= (w = need("encoder/layer_norm/gamma", {d})) != nullptr;
        if (w) o_lnf_g = put_f32(w, d);
        ok &= (w = need("encoder/layer_norm/beta", {d})) != nullptr;
        if (w) o_lnf_b = put_f32(w, d);
    }
    for (int i = 0; i < L && ok; ++i) {
        const std::string p = "encoder/layer_" + std::to_string(i);
        const float* w;
        Offs& o = offs[i];
        const float *g1 = need(p + "/self_attention/layer_norm/gamma", {d}), *be1 = need(p + "/self_attention/layer_norm/beta", {d});
        const float *wq = need(p + "/self_attention/linear_0/weight", {3 * d, d}), *bq = need(p + "/self_attention/linear_0/bias", {3 * d});
        ok &= g1 && be1 && wq && bq;
        if (g1 && be1 && wq && bq) put_ln_folded(wq, bq, g1, be1, 3 * d, d, &o.wqkv, &o.c1_qkv, &o.c2_qkv);
        ok &= (w = need(p + "/self_attention/linear_1/weight", {d, d})) != nullptr;
        if (w) o.wo = put_bf16(w, (size_t)d * d);
        ok &= (w = need(p + "/self_attention/linear_1/bias", {d})) != nullptr;
        if (w) o.bo = put_f32(w, d);
        const float *g2 = need(p + "/ffn/layer_norm/gamma", {d}), *be2 = need(p + "/ffn/layer_norm/beta", {d});
        const float *w1 = need(p + "/ffn/linear_0/weight", {f, d}), *b1 = need(p + "/ffn/linear_0/bias", {f});
        ok &= g2 && be2 && w1 && b1;
        if (g2 && be2 && w1 && b1) put_ln_folded(w1, b1, g2, be2, f, d, &o.w1, &o.c1_fc1, &o.c2_fc1);
        ok &= (w = need(p + "/ffn/linear_1/weight", {d, f})) != nullptr;
        if (w) o.w2 = put_bf16(w, (size_t)d * f);
        ok &= (w = need(p + "/ffn/linear_1/bias", {d})) != nullptr;
        if (w) o.b2 = put_f32(w, d);
    }
    if (!ok) {
        delete pl;
        return cudaErrorInvalidValue;
    }

    cudaError_t e;
    auto fail = [&](cudaError_t err) {
        *why = cudaGetErrorString(err);
        encoder_plan_destroy(pl);
        return err;
    };
    if ((e = cudaMalloc(&pl->d_weights, blob.size())) != cudaSuccess) return fail(e);
    if ((e = cudaMemcpy(pl->d_weights, blob.data(), blob.size(), cudaMemcpyHostToDevice)) != cudaSuccess) return fail(e);
    if ((e = gemm_init_device()) != cudaSuccess) return fail(e);
    if ((e = attention_init_device()) != cudaSuccess) return fail(e);

    char* base = pl->d_weights;
    auto F = [&](size_t off) { return reinterpret_cast<const float*>(base + off); };
    pl->conv1_b = F(o_conv1b);
    pl->conv2_b = F(o_conv2b);
    pl->pos = F(o_pos);
    pl->lnf_g = F(o_lnf_g);
    pl->lnf_b = F(o_lnf_b);
    if ((e = map2d(&pl->m_conv1, base + o_conv1w, 3ull * cp, d, gemm_b_box_rows())) != cudaSuccess) return fail(e);
    if ((e = map2d(&pl->m_conv2, base + o_conv2w, 3ull * d, d, gemm_b_box_rows())) != cudaSuccess) return fail(e);
    for (int i = 0; i < L; ++i) {
        LayerW& lw = pl->layers[i];
        const Offs& o = offs[i];
        lw.c1_qkv = F(o.c1_qkv); lw.c2_qkv = F(o.c2_qkv); lw.bo = F(o.bo);
        lw.c1_fc1 = F(o.c1_fc1); lw.c2_fc1 = F(o.c2_fc1); lw.b2 = F(o.b2);
        if ((e = map2d(&lw.m_qkv, base + o.wqkv, d, 3ull * d, gemm_b_box_rows())) != cudaSuccess) return fail(e);
        if ((e = map2d(&lw.m_o, base + o.wo, d, d, gemm_b_box_rows())) != cudaSuccess) return fail(e);
        if ((e = map2d(&lw.m_fc1, base + o.w1, d, f, gemm_b_box_rows())) != cudaSuccess) return fail(e);
        if ((e = map2d(&lw.m_fc2, base + o.w2, f, d, gemm_b_box_rows())) != cudaSuccess) return fail(e);
    }
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

cudaError_t encoder_run(EncoderPlan* pl, const float* mel, int batch, int frames, void* out_bf16, void* workspace,
                        size_t ws_bytes, cudaStream_t stream) {
    const auto& c = pl->cfg;
    const int d = c.d_model, f = c.d_ffn, T = c.n_ctx, cp = pl->c_pad;
    // mel == nullptr: the time-major conv1 operand (encoder_workspace_conv1_operand) has already been written by the
    // fused log-mel kernel; otherwise mel is f32 [batch, n_mels, frames] in upstream's layout and is transposed here
    if (batch <= 0 || (mel && (frames <= 0 || frames > kFramesIn))) {
        pl->error = "invalid batch / frames";
        return cudaErrorInvalidValue;
    }
    const WsLayout w = ws_layout(*pl, batch);
    if (!workspace || ws_bytes < w.total || (reinterpret_cast<uintptr_t>(workspace) & 255)) {
        pl->error = "workspace missing, too small or not 256-byte aligned";
        return cudaErrorInvalidValue;
    }
    char* ws = static_cast<char*>(workspace);
    void* melT = ws + w.melT;
    void* x = ws + w.x;                                  // residual stream, f16
    float2* stats = reinterpret_cast<float2*>(ws + w.stats);   // LayerNorm partials of the residual stream's rows
    const int parts = gemm_stats_parts(d);
    void* ctx = ws + w.ctx;
    void* qk = ws + w.qk;
    void* vt = ws + w.vt;
    void* h = ws + w.h;
    void* c1 = h;                              // conv1 output is dead before the first fc1 writes h
    const long long M = (long long)batch * T;

    if (pl->cached_ws != workspace || pl->cached_batch != batch) {
        ARIES_TRY(map2d(&pl->a_mel, melT, cp, (unsigned long long)batch * kRowsPadded, 128), "tensor map (mel)");
        ARIES_TRY(map2d(&pl->a_c1, c1, 2ull * d, (unsigned long long)batch * (kRowsPadded / 2), 128), "tensor map (conv1 out)");
        ARIES_TRY(map2d(&pl->a_x, x, d, M, 128), "tensor map (x)");
        ARIES_TRY(map2d(&pl->a_ctx, ctx, d, M, 128), "tensor map (ctx)");
        ARIES_TRY(map2d(&pl->a_h, h, f, M, 128), "tensor map (h)");
        ARIES_TRY(attention_make_maps(qk, vt, batch, T, d, c.n_heads, pl->t_pad, &pl->a_attn), "tensor map (attention)");
        pl->cached_ws = workspace;
        pl->cached_batch = batch;
    }
    int launches = 0;

    // zero rows before / after each item's frames (conv padding = 1) in both time-major buffers
    ARIES_TRY(cudaMemset2DAsync(melT, (size_t)kRowsPadded * cp * 2, 0, (size_t)cp * 2, batch, stream), "memset");
    ARIES_TRY(cudaMemset2DAsync(static_cast<char*>(melT) + (size_t)(kRowsPadded - 1) * cp * 2, (size_t)kRowsPadded * cp * 2,
                                0, (size_t)cp * 2, batch, stream), "memset");
    ARIES_TRY(cudaMemset2DAsync(c1, (size_t)kRowsPadded * d * 2, 0, (size_t)d * 2, batch, stream), "memset");
    ARIES_TRY(cudaMemset2DAsync(static_cast<char*>(c1) + (size_t)(kRowsPadded - 1) * d * 2, (size_t)kRowsPadded * d * 2, 0,
                                (size_t)d * 2, batch, stream), "memset");

    if (mel) {
        pl->prof.begin(KC_TRANSPOSE, stream);
        ARIES_TRY(mel_to_time_major(mel, batch, c.n_mels, frames, melT, cp, stream), "mel transpose");
        pl->prof.end(stream);
        ++launches;
    }

    GemmParams g{};
    // conv1: row r = b * 3002 + t reads padded rows t, t+1, t+2 (taps 0..2); writes padded row t + 1
    g = GemmParams{};
    g.M = batch * kRowsPadded; g.N = d; g.K = 3 * cp; g.a_cols = cp;
    g.p_in = kRowsPadded; g.t_valid = kFramesIn; g.p_out = kRowsPadded; g.row_off = 1; g.ldo = d;
    g.bias = pl->conv1_b; g.out = c1;
    pl->prof.begin(KC_CONV1, stream);
    ARIES_TRY(gemm_launch(EPI_BIAS_GELU_BF16, pl->a_mel, pl->m_conv1, g, pl->sm_count, stream), "conv1");
    pl->prof.end(stream);
    ++launches;
    // conv2 (stride 2): row r = b * 1501 + t reads the row PAIR t (taps 0, 1) and the first half of pair t + 1 (tap 2)
    g = GemmParams{};
    g.M = batch * (kRowsPadded / 2); g.N = d; g.K = 3 * d; g.a_cols = 2 * d;
    g.p_in = kRowsPadded / 2; g.t_valid = T; g.p_out = T; g.row_off = 0; g.ldo = d;
    g.bias = pl->conv2_b; g.pos = pl->pos; g.out = x; g.stats_out = stats;
    pl->prof.begin(KC_CONV2, stream);
    ARIES_TRY(gemm_launch(EPI_BIAS_GELU_POS_F16, pl->a_c1, pl->m_conv2, g, pl->sm_count, stream), "conv2");
    pl->prof.end(stream);
    ++launches;

    auto plain = [&](int N, int K) {
        GemmParams q{};
        q.M = (int)M; q.N = N; q.K = K; q.a_cols = K;
        q.p_in = (int)M; q.t_valid = (int)M; q.p_out = (int)M; q.row_off = 0; q.ldo = N;
        return q;
    };
    AttnParams ap{batch, T, d, c.n_heads, ctx};
    for (int i = 0; i < c.n_layers; ++i) {
        const LayerW& lw = pl->layers[i];
        g = plain(3 * d, d);
        g.p_in = T; g.t_valid = T; g.p_out = T; g.ldo = 2 * d;
        g.bias = lw.c2_qkv; g.c1 = lw.c1_qkv; g.stats_in = stats; g.stats_parts = parts; g.ln_dim = d; g.ln_eps = kLnEps;
        g.out = qk; g.out2 = vt; g.n_split = 2 * d; g.t_pad = pl->t_pad;
        pl->prof.begin(KC_QKV, stream);
        ARIES_TRY(gemm_launch(EPI_LN_QKV_SPLIT_BF16, pl->a_x, lw.m_qkv, g, pl->sm_count, stream), "qkv projection");
        pl->prof.end(stream);
        pl->prof.begin(KC_ATTENTION, stream);
        ARIES_TRY(attention_launch(pl->a_attn, ap, stream), "attention");
        pl->prof.end(stream);
        g = plain(d, d);
        g.bias = lw.bo; g.resid = x; g.out = x; g.stats_out = stats;
        pl->prof.begin(KC_OPROJ, stream);
        ARIES_TRY(gemm_launch(EPI_BIAS_RESID_F16, pl->a_ctx, lw.m_o, g, pl->sm_count, stream), "output projection");
        pl->prof.end(stream);
        g = plain(f, d);
        g.bias = lw.c2_fc1; g.c1 = lw.c1_fc1; g.stats_in = stats; g.stats_parts = parts; g.ln_dim = d; g.ln_eps = kLnEps;
        g.out = h;
        pl->prof.begin(KC_FC1, stream);
        ARIES_TRY(gemm_launch(EPI_LN_GELU_BF16, pl->a_x, lw.m_fc1, g, pl->sm_count, stream), "fc1");
        pl->prof.end(stream);
        g = plain(d, f);
        g.bias = lw.b2; g.resid = x; g.out = x; g.stats_out = stats;
        pl->prof.begin(KC_FC2, stream);
        ARIES_TRY(gemm_launch(EPI_BIAS_RESID_F16, pl->a_h, lw.m_fc2, g, pl->sm_count, stream), "fc2");
        pl->prof.end(stream);
        launches += 5;
    }
    pl->prof.begin(KC_LAYERNORM, stream);
    ARIES_TRY(layernorm_launch(x, pl->lnf_g, pl->lnf_b, out_bf16, M, d, kLnEps, stream), "final layer norm");
    pl->prof.end(stream);
    ++launches;
    pl->last_launches = launches;
    return cudaSuccess;
}

Profiler* encoder_plan_profiler(EncoderPlan* pl) { return &pl->prof; }

void* encoder_workspace_conv1_operand(const EncoderPlan* pl, void* workspace, int batch, int* c_pad) {
    *c_pad = pl->c_pad;
    return static_cast<char*>(workspace) + ws_layout(*pl, batch).melT;
}

}  // namespace aries
