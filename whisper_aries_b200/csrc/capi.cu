// extern "C" surface of libaries_b200.so (include/aries_b200.h, include/aries_b200_test.h).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/aries_b200.h"
#include "../../include/aries_b200_test.h"
#include "attention.h"
#include "decoder.h"
#include "encoder.h"
#include "gemm.h"
#include "layernorm.h"
#include "logmel.h"
#include "profiler.h"
#include "skinny.h"
#include "vad.h"

namespace {

thread_local std::string g_error;

int fail(int code, const std::string& msg) {
    g_error = msg;
    return code;
}
int fail_cuda(const char* what, cudaError_t e) {
    g_error = std::string(what) + ": " + cudaGetErrorString(e);
    cudaGetLastError();   // clear the sticky-free error state
    return e == cudaErrorMemoryAllocation ? ARIES_ENOMEM : (e == cudaErrorInvalidValue ? ARIES_EINVAL : ARIES_ECUDA);
}

constexpr unsigned kMagicCtx = 0xA51E5001u, kMagicMel = 0xA51E5002u, kMagicEnc = 0xA51E5003u, kMagicDec = 0xA51E5004u;

}  // namespace

struct aries_ctx {
    unsigned magic;
    int device;
    int sm_count;
    bool kernels_ready;
    long long* d_chunks = nullptr;     // aries_collect_chunks: device copy of (starts | offsets), grown on demand
    size_t chunks_cap = 0;
};

struct aries_mel {
    unsigned magic;
    aries_ctx* ctx;
    aries::LogmelPlan* plan;
    int last_launches;
    aries::Profiler* prof;      // borrowed from the encoder during aries_encode_pcm when profiling is on
    // scratch for the host-buffer variant
    float* d_pcm;
    size_t pcm_cap;
    float* d_out;
    size_t out_cap;
};

struct aries_encoder {
    unsigned magic;
    aries_ctx* ctx;
    aries::EncoderPlan* plan;
    // scratch for the host-buffer variant
    void* d_ws;
    size_t ws_cap;
    float* d_mel;
    size_t mel_cap;
    void* d_out;
    size_t out_cap;
};

struct aries_decoder {
    unsigned magic;
    aries_ctx* ctx;
    aries::DecoderPlan* plan;
};

namespace {

int use(const aries_ctx* ctx) {
    if (!ctx || ctx->magic != kMagicCtx) return fail(ARIES_ESTATE, "invalid context handle");
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return fail_cuda("cudaSetDevice", e);
    return ARIES_OK;
}

int ensure_kernels(aries_ctx* ctx) {
    if (ctx->kernels_ready) return ARIES_OK;
    cudaError_t e;
    if ((e = aries::gemm_init_device()) != cudaSuccess) return fail_cuda("gemm_init_device", e);
    if ((e = aries::attention_init_device()) != cudaSuccess) return fail_cuda("attention_init_device", e);
    if ((e = aries::skinny_init_device()) != cudaSuccess) return fail_cuda("skinny_init_device", e);
    ctx->kernels_ready = true;
    return ARIES_OK;
}

template <class T>
int grow(T** ptr, size_t* cap, size_t bytes) {
    if (bytes <= *cap) return ARIES_OK;
    cudaFree(*ptr);
    *ptr = nullptr;
    *cap = 0;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(ptr), bytes);
    if (e != cudaSuccess) return fail_cuda("cudaMalloc", e);
    *cap = bytes;
    return ARIES_OK;
}

}  // namespace

extern "C" {

int aries_abi_version(void) { return 101; }

const char* aries_last_error(void) { return g_error.c_str(); }

int aries_init(int device, aries_ctx** out) {
    if (!out) return fail(ARIES_EINVAL, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(ARIES_ECUDA, std::string("no CUDA device is usable (this library has no CPU fallback): ") +
                                     (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    if (device < 0 || device >= n) return fail(ARIES_EINVAL, "device index out of range");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail_cuda("cudaGetDeviceProperties", e);
    if (prop.major != 10)
        return fail(ARIES_ECUDA, "device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                                     "; this library is compiled for sm_100a (B200) only");
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail_cuda("cudaSetDevice", e);
    aries_ctx* c = new (std::nothrow) aries_ctx();
    if (c) {
        c->magic = kMagicCtx;
        c->device = device;
        c->sm_count = prop.multiProcessorCount;
        c->kernels_ready = false;
    }
    if (!c) return fail(ARIES_ENOMEM, "out of host memory");
    *out = c;
    return ARIES_OK;
}

int aries_destroy(aries_ctx* ctx) {
    if (!ctx) return ARIES_OK;
    if (ctx->magic != kMagicCtx) return fail(ARIES_ESTATE, "invalid context handle");
    ctx->magic = 0;
    if (ctx->d_chunks && cudaSetDevice(ctx->device) == cudaSuccess) cudaFree(ctx->d_chunks);
    delete ctx;
    return ARIES_OK;
}

int aries_device(const aries_ctx* ctx) { return (ctx && ctx->magic == kMagicCtx) ? ctx->device : -1; }
int aries_sm_count(const aries_ctx* ctx) { return (ctx && ctx->magic == kMagicCtx) ? ctx->sm_count : -1; }

// ------------------------------------------------------------------------------------------------ log-mel
int aries_logmel_create(aries_ctx* ctx, int n_mels, const float* mel_filters, aries_mel** out) {
    if (!out) return fail(ARIES_EINVAL, "out is NULL");
    *out = nullptr;
    int rc = use(ctx);
    if (rc) return rc;
    if (!mel_filters) return fail(ARIES_EINVAL, "mel_filters is NULL");
    aries::LogmelPlan* plan = nullptr;
    const char* why = "";
    cudaError_t e = aries::logmel_plan_create(ctx->device, ctx->sm_count, n_mels, mel_filters, &plan, &why);
    if (e != cudaSuccess) {
        g_error = std::string("aries_logmel_create: ") + why;
        cudaGetLastError();
        return e == cudaErrorInvalidValue ? ARIES_EINVAL : ARIES_ECUDA;
    }
    aries_mel* m = new (std::nothrow) aries_mel{kMagicMel, ctx, plan, 0, nullptr, nullptr, 0, nullptr, 0};
    if (!m) {
        aries::logmel_plan_destroy(plan);
        return fail(ARIES_ENOMEM, "out of host memory");
    }
    *out = m;
    return ARIES_OK;
}

int aries_logmel_destroy(aries_mel* mel) {
    if (!mel) return ARIES_OK;
    if (mel->magic != kMagicMel) return fail(ARIES_ESTATE, "invalid log-mel handle");
    use(mel->ctx);
    aries::logmel_plan_destroy(mel->plan);
    cudaFree(mel->d_pcm);
    cudaFree(mel->d_out);
    mel->magic = 0;
    delete mel;
    return ARIES_OK;
}

int64_t aries_logmel_num_frames(int64_t n_samples, int padding) {
    if (n_samples < 0 || padding < 0) return 0;
    return (n_samples + padding) / 160;
}

int aries_logmel_run(aries_mel* mel, const float* pcm_dev, int batch, int64_t n_samples, int64_t pcm_stride,
                     int padding, float* out_dev, int frames_out, void* stream) {
    if (!mel || mel->magic != kMagicMel) return fail(ARIES_ESTATE, "invalid log-mel handle");
    int rc = use(mel->ctx);
    if (rc) return rc;
    if (batch <= 0 || n_samples <= 0 || padding < 0 || frames_out < 0 || pcm_stride < n_samples || !pcm_dev || !out_dev)
        return fail(ARIES_EINVAL, "aries_logmel_run: need batch > 0, n_samples > 0, padding >= 0, frames_out >= 0, "
                                  "pcm_stride >= n_samples and non-NULL buffers");
    if ((n_samples + padding) / 160 > (1 << 24)) return fail(ARIES_EINVAL, "aries_logmel_run: signal too long");
    cudaError_t e = aries::logmel_run(mel->plan, pcm_dev, batch, n_samples, pcm_stride, padding, out_dev, frames_out,
                                      static_cast<cudaStream_t>(stream), &mel->last_launches, mel->prof);
    if (e != cudaSuccess) return fail_cuda("aries_logmel_run", e);
    return ARIES_OK;
}

int aries_logmel_run_host(aries_mel* mel, const float* pcm_host, int batch, int64_t n_samples, int padding,
                          float* out_host, int frames_out) {
    if (!mel || mel->magic != kMagicMel) return fail(ARIES_ESTATE, "invalid log-mel handle");
    int rc = use(mel->ctx);
    if (rc) return rc;
    if (batch <= 0 || n_samples <= 0 || !pcm_host || !out_host || frames_out < 0)
        return fail(ARIES_EINVAL, "aries_logmel_run_host: bad arguments");
    const size_t in_bytes = (size_t)batch * n_samples * 4;
    const size_t out_bytes = (size_t)batch * aries::logmel_plan_n_mels(mel->plan) * frames_out * 4;
    if ((rc = grow(&mel->d_pcm, &mel->pcm_cap, in_bytes))) return rc;
    if ((rc = grow(&mel->d_out, &mel->out_cap, out_bytes ? out_bytes : 4))) return rc;
    cudaError_t e;
    if ((e = cudaMemcpyAsync(mel->d_pcm, pcm_host, in_bytes, cudaMemcpyHostToDevice, 0)) != cudaSuccess)
        return fail_cuda("H2D copy", e);
    if ((rc = aries_logmel_run(mel, mel->d_pcm, batch, n_samples, n_samples, padding, mel->d_out, frames_out, nullptr)))
        return rc;
    if (out_bytes && (e = cudaMemcpyAsync(out_host, mel->d_out, out_bytes, cudaMemcpyDeviceToHost, 0)) != cudaSuccess)
        return fail_cuda("D2H copy", e);
    if ((e = cudaStreamSynchronize(0)) != cudaSuccess) return fail_cuda("aries_logmel_run_host", e);
    return ARIES_OK;
}

int aries_logmel_last_launches(const aries_mel* mel) { return (mel && mel->magic == kMagicMel) ? mel->last_launches : -1; }

// ------------------------------------------------------------------------------------------------ encoder
int aries_encoder_create(aries_ctx* ctx, const aries_encoder_cfg* cfg, const aries_weight_desc* weights,
                         int n_weights, aries_encoder** out) {
    if (!out) return fail(ARIES_EINVAL, "out is NULL");
    *out = nullptr;
    int rc = use(ctx);
    if (rc) return rc;
    if (!cfg || !weights || n_weights <= 0) return fail(ARIES_EINVAL, "aries_encoder_create: cfg / weights missing");
    std::vector<aries::WeightView> views(n_weights);
    for (int i = 0; i < n_weights; ++i) {
        if (!weights[i].name || weights[i].ndim < 1 || weights[i].ndim > 4)
            return fail(ARIES_EINVAL, "aries_encoder_create: malformed weight descriptor");
        views[i].name = weights[i].name;
        views[i].data = weights[i].data;
        views[i].ndim = weights[i].ndim;
        for (int k = 0; k < 4; ++k) views[i].shape[k] = k < weights[i].ndim ? weights[i].shape[k] : 1;
    }
    aries::EncoderShapeC shape{cfg->n_mels, cfg->d_model, cfg->n_heads, cfg->n_layers, cfg->d_ffn, cfg->n_ctx};
    aries::EncoderPlan* plan = nullptr;
    std::string why;
    cudaError_t e = aries::encoder_plan_create(ctx->device, ctx->sm_count, shape, views.data(), n_weights, &plan, &why);
    if (e != cudaSuccess) {
        g_error = "aries_encoder_create: " + why;
        cudaGetLastError();
        return e == cudaErrorInvalidValue ? ARIES_EINVAL : (e == cudaErrorMemoryAllocation ? ARIES_ENOMEM : ARIES_ECUDA);
    }
    aries_encoder* h = new (std::nothrow) aries_encoder{kMagicEnc, ctx, plan, nullptr, 0, nullptr, 0, nullptr, 0};
    if (!h) {
        aries::encoder_plan_destroy(plan);
        return fail(ARIES_ENOMEM, "out of host memory");
    }
    *out = h;
    return ARIES_OK;
}

int aries_encoder_destroy(aries_encoder* enc) {
    if (!enc) return ARIES_OK;
    if (enc->magic != kMagicEnc) return fail(ARIES_ESTATE, "invalid encoder handle");
    use(enc->ctx);
    aries::encoder_plan_destroy(enc->plan);
    cudaFree(enc->d_ws);
    cudaFree(enc->d_mel);
    cudaFree(enc->d_out);
    enc->magic = 0;
    delete enc;
    return ARIES_OK;
}

size_t aries_encoder_workspace_bytes(const aries_encoder* enc, int batch) {
    if (!enc || enc->magic != kMagicEnc || batch <= 0) return 0;
    return aries::encoder_workspace_bytes(enc->plan, batch);
}

static int check_features(const aries_encoder* enc, int batch, int frames) {
    if (batch <= 0 || frames <= 0 || frames > 3000) {
        const aries::EncoderShapeC* c = aries::encoder_plan_cfg(enc->plan);
        return fail(ARIES_EINVAL, "Invalid input features shape: expected an input with shape (batch, " +
                                      std::to_string(c->n_mels) + ", <=3000), but got (" + std::to_string(batch) + ", " +
                                      std::to_string(c->n_mels) + ", " + std::to_string(frames) + ")");
    }
    return ARIES_OK;
}

int aries_encoder_run(aries_encoder* enc, const float* mel_dev, int batch, int frames, void* out_dev, void* workspace,
                      size_t workspace_bytes, void* stream) {
    if (!enc || enc->magic != kMagicEnc) return fail(ARIES_ESTATE, "invalid encoder handle");
    int rc = use(enc->ctx);
    if (rc) return rc;
    if ((rc = check_features(enc, batch, frames))) return rc;
    if (!mel_dev || !out_dev) return fail(ARIES_EINVAL, "aries_encoder_run: NULL buffer");
    cudaError_t e = aries::encoder_run(enc->plan, mel_dev, batch, frames, out_dev, workspace, workspace_bytes,
                                       static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) {
        g_error = std::string("aries_encoder_run: ") + aries::encoder_plan_error(enc->plan);
        cudaGetLastError();
        return e == cudaErrorInvalidValue ? ARIES_EINVAL : ARIES_ECUDA;
    }
    return ARIES_OK;
}

int aries_encoder_run_host(aries_encoder* enc, const float* mel_host, int batch, int frames, uint16_t* out_host) {
    if (!enc || enc->magic != kMagicEnc) return fail(ARIES_ESTATE, "invalid encoder handle");
    int rc = use(enc->ctx);
    if (rc) return rc;
    if ((rc = check_features(enc, batch, frames))) return rc;
    if (!mel_host || !out_host) return fail(ARIES_EINVAL, "aries_encoder_run_host: NULL buffer");
    const aries::EncoderShapeC* c = aries::encoder_plan_cfg(enc->plan);
    const size_t in_bytes = (size_t)batch * c->n_mels * frames * 4;
    const size_t out_bytes = (size_t)batch * c->n_ctx * c->d_model * 2;
    const size_t ws = aries::encoder_workspace_bytes(enc->plan, batch);
    if ((rc = grow(&enc->d_mel, &enc->mel_cap, in_bytes))) return rc;
    if ((rc = grow(&enc->d_out, &enc->out_cap, out_bytes))) return rc;
    if ((rc = grow(&enc->d_ws, &enc->ws_cap, ws))) return rc;
    cudaError_t e;
    if ((e = cudaMemcpyAsync(enc->d_mel, mel_host, in_bytes, cudaMemcpyHostToDevice, 0)) != cudaSuccess)
        return fail_cuda("H2D copy", e);
    if ((rc = aries_encoder_run(enc, enc->d_mel, batch, frames, enc->d_out, enc->d_ws, enc->ws_cap, nullptr))) return rc;
    if ((e = cudaMemcpyAsync(out_host, enc->d_out, out_bytes, cudaMemcpyDeviceToHost, 0)) != cudaSuccess)
        return fail_cuda("D2H copy", e);
    if ((e = cudaStreamSynchronize(0)) != cudaSuccess) return fail_cuda("aries_encoder_run_host", e);
    return ARIES_OK;
}

int aries_encode_pcm(aries_encoder* enc, aries_mel* mel, const float* pcm_dev, int batch, int64_t n_samples,
                     int64_t pcm_stride, void* out_dev, void* workspace, size_t workspace_bytes, void* stream) {
    if (!enc || enc->magic != kMagicEnc) return fail(ARIES_ESTATE, "invalid encoder handle");
    if (!mel || mel->magic != kMagicMel) return fail(ARIES_ESTATE, "invalid log-mel handle");
    if (mel->ctx != enc->ctx) return fail(ARIES_ESTATE, "log-mel and encoder handles belong to different contexts");
    int rc = use(enc->ctx);
    if (rc) return rc;
    const aries::EncoderShapeC* c = aries::encoder_plan_cfg(enc->plan);
    if (aries::logmel_plan_n_mels(mel->plan) != c->n_mels)
        return fail(ARIES_EINVAL, "aries_encode_pcm: the log-mel handle's n_mels differs from the encoder's");
    if (batch <= 0 || n_samples <= 0 || n_samples > 480000)
        return fail(ARIES_EINVAL, "aries_encode_pcm: need batch > 0 and 0 < n_samples <= 480000 (one 30-second window)");
    const size_t need = aries::encoder_workspace_bytes(enc->plan, batch);
    if (!workspace || workspace_bytes < need) return fail(ARIES_EINVAL, "aries_encode_pcm: workspace too small");
    if (!pcm_dev || pcm_stride < n_samples) return fail(ARIES_EINVAL, "aries_encode_pcm: NULL pcm or pcm_stride < n_samples");
    if (reinterpret_cast<uintptr_t>(workspace) & 255) return fail(ARIES_EINVAL, "aries_encode_pcm: workspace must be 256-byte aligned");
    // one launch: PCM -> log-mel -> bf16 time-major conv1 operand inside the workspace (the f32 mel never exists in HBM)
    int c_pad = 0;
    void* operand = aries::encoder_workspace_conv1_operand(enc->plan, workspace, batch, &c_pad);
    aries::Profiler* prof = aries::encoder_plan_profiler(enc->plan);
    cudaError_t e = aries::logmel_run_time_major(mel->plan, pcm_dev, batch, n_samples, pcm_stride, 160, operand, 3000, c_pad,
                                                 static_cast<cudaStream_t>(stream), &mel->last_launches,
                                                 prof->enabled() ? prof : nullptr);
    if (e == cudaErrorInvalidValue)
        return fail(ARIES_EINVAL, "aries_encode_pcm: the fused path supports n_mels <= 128 (every Whisper size); use "
                                  "aries_logmel_run + aries_encoder_run for wider filter banks");
    if (e != cudaSuccess) return fail_cuda("aries_encode_pcm (log-mel)", e);
    e = aries::encoder_run(enc->plan, nullptr, batch, 3000, out_dev, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) {
        g_error = std::string("aries_encode_pcm: ") + aries::encoder_plan_error(enc->plan);
        cudaGetLastError();
        return e == cudaErrorInvalidValue ? ARIES_EINVAL : ARIES_ECUDA;
    }
    return ARIES_OK;
}

// ------------------------------------------------------------------------------------------------ VAD front end (row f4)
int64_t aries_vad_num_windows(int64_t n_samples) { return n_samples < 0 ? 0 : aries::vad_num_windows(n_samples); }

int aries_vad_speech_timestamps(const float* probs_host, int64_t n_windows, int64_t audio_len, const aries_vad_opts* opts,
                                int64_t* starts, int64_t* ends, int cap, int* n_out) {
    if (!probs_host || n_windows < 0 || audio_len < 0 || !opts || !n_out || cap < 0 || (cap > 0 && (!starts || !ends)))
        return fail(ARIES_EINVAL, "aries_vad_speech_timestamps: bad arguments");
    if (!(opts->threshold > 0.0f) || opts->min_speech_duration_ms < 0 || opts->min_silence_duration_ms < 0 ||
        opts->speech_pad_ms < 0)
        return fail(ARIES_EINVAL, "aries_vad_speech_timestamps: threshold must be > 0 and the durations >= 0");
    aries::VadOpts o;
    o.threshold = opts->threshold;
    o.neg_threshold = opts->neg_threshold;
    o.min_speech_duration_ms = opts->min_speech_duration_ms;
    o.max_speech_duration_s = opts->max_speech_duration_s;
    o.min_silence_duration_ms = opts->min_silence_duration_ms;
    o.speech_pad_ms = opts->speech_pad_ms;
    std::vector<long long> s, e;
    aries::vad_speech_timestamps(probs_host, n_windows, audio_len, o, &s, &e);
    *n_out = (int)s.size();
    if ((int)s.size() > cap) return fail(ARIES_EINVAL, "aries_vad_speech_timestamps: more chunks than `cap`");
    for (size_t i = 0; i < s.size(); ++i) {
        starts[i] = s[i];
        ends[i] = e[i];
    }
    return ARIES_OK;
}

int aries_vad_energy_probs(aries_ctx* ctx, const float* pcm_dev, int64_t n_samples, float center_db, float width_db,
                           float* probs_dev, void* stream) {
    int rc = use(ctx);
    if (rc) return rc;
    if (!pcm_dev || !probs_dev || n_samples <= 0 || !(width_db > 0.0f))
        return fail(ARIES_EINVAL, "aries_vad_energy_probs: need non-NULL buffers, n_samples > 0 and width_db > 0");
    cudaError_t e = aries::vad_energy_probs(pcm_dev, n_samples, center_db, width_db, probs_dev, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail_cuda("aries_vad_energy_probs", e);
    return ARIES_OK;
}

int aries_collect_chunks(aries_ctx* ctx, const float* pcm_dev, int64_t n_samples, const int64_t* starts_host,
                         const int64_t* ends_host, int n_chunks, float* out_dev, int64_t out_cap, int64_t* out_len,
                         void* stream) {
    int rc = use(ctx);
    if (rc) return rc;
    if (n_chunks < 0 || !out_len || (n_chunks > 0 && (!pcm_dev || !starts_host || !ends_host)))
        return fail(ARIES_EINVAL, "aries_collect_chunks: bad arguments");
    std::vector<long long> tab(2 * (size_t)(n_chunks > 0 ? n_chunks : 1));
    long long total = 0;
    for (int k = 0; k < n_chunks; ++k) {
        if (starts_host[k] < 0 || ends_host[k] < starts_host[k] || ends_host[k] > n_samples)
            return fail(ARIES_EINVAL, "aries_collect_chunks: chunk outside [0, n_samples]");
        tab[k] = starts_host[k];
        tab[n_chunks + k] = total;
        total += ends_host[k] - starts_host[k];
    }
    *out_len = total;
    if (total == 0) return ARIES_OK;
    if (!out_dev || out_cap < total) return fail(ARIES_EINVAL, "aries_collect_chunks: out_dev is NULL or smaller than the chunks");
    if ((rc = grow(&ctx->d_chunks, &ctx->chunks_cap, tab.size() * sizeof(long long)))) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // pageable source: the copy is staged before the call returns, so `tab` may go out of scope
    cudaError_t e = cudaMemcpyAsync(ctx->d_chunks, tab.data(), (size_t)2 * n_chunks * sizeof(long long), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return fail_cuda("aries_collect_chunks (table upload)", e);
    e = aries::collect_chunks(pcm_dev, ctx->d_chunks, ctx->d_chunks + n_chunks, n_chunks, total, out_dev, ctx->sm_count, st);
    if (e != cudaSuccess) return fail_cuda("aries_collect_chunks", e);
    return ARIES_OK;
}

int aries_pcm_s16_to_f32(aries_ctx* ctx, const int16_t* pcm_s16_dev, float* out_dev, int64_t n, void* stream) {
    int rc = use(ctx);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!pcm_s16_dev || !out_dev))) return fail(ARIES_EINVAL, "aries_pcm_s16_to_f32: NULL buffer or negative length");
    cudaError_t e = aries::pcm_s16_to_f32(reinterpret_cast<const short*>(pcm_s16_dev), out_dev, n, ctx->sm_count,
                                          static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail_cuda("aries_pcm_s16_to_f32", e);
    return ARIES_OK;
}

int aries_encoder_last_launches(const aries_encoder* enc) {
    return (enc && enc->magic == kMagicEnc) ? aries::encoder_plan_last_launches(enc->plan) : -1;
}

int aries_encoder_set_profiling(aries_encoder* enc, int on) {
    if (!enc || enc->magic != kMagicEnc) return fail(ARIES_ESTATE, "invalid encoder handle");
    aries::encoder_plan_profiler(enc->plan)->enable(on != 0);
    return ARIES_OK;
}

int aries_encoder_collect_profile(aries_encoder* enc, float* ms, int* counts, int n) {
    if (!enc || enc->magic != kMagicEnc) return fail(ARIES_ESTATE, "invalid encoder handle");
    if (!ms || !counts || n < aries::KC_COUNT) return fail(ARIES_EINVAL, "aries_encoder_collect_profile: need n >= 11");
    int rc = use(enc->ctx);
    if (rc) return rc;
    cudaError_t e = aries::encoder_plan_profiler(enc->plan)->collect(ms, counts);
    if (e != cudaSuccess) return fail_cuda("aries_encoder_collect_profile", e);
    return ARIES_OK;
}


// ------------------------------------------------------------------------------------------------ decoder (row f1)
int aries_decoder_create(aries_ctx* ctx, const aries_decoder_cfg* cfg, const aries_weight_desc* weights, int n_weights,
                         int max_batch, aries_decoder** out) {
    if (!out) return fail(ARIES_EINVAL, "out is NULL");
    *out = nullptr;
    int rc = use(ctx);
    if (rc) return rc;
    if (!cfg || !weights || n_weights <= 0) return fail(ARIES_EINVAL, "aries_decoder_create: NULL argument");
    std::vector<aries::WeightView> views(n_weights);
    for (int i = 0; i < n_weights; ++i) {
        if (!weights[i].name || !weights[i].data || weights[i].ndim < 1 || weights[i].ndim > 4)
            return fail(ARIES_EINVAL, "aries_decoder_create: malformed weight descriptor");
        views[i].name = weights[i].name;
        views[i].data = weights[i].data;
        views[i].ndim = weights[i].ndim;
        for (int k = 0; k < 4; ++k) views[i].shape[k] = weights[i].shape[k];
    }
    aries::DecoderShapeC c{cfg->vocab, cfg->d_model, cfg->n_heads, cfg->n_layers, cfg->d_ffn, cfg->n_text_ctx, cfg->n_audio_ctx};
    aries::DecoderPlan* plan = nullptr;
    std::string why;
    cudaError_t e = aries::decoder_plan_create(ctx->device, ctx->sm_count, c, views.data(), n_weights, max_batch, &plan, &why);
    if (e != cudaSuccess) {
        g_error = "aries_decoder_create: " + why;
        cudaGetLastError();
        return e == cudaErrorInvalidValue ? ARIES_EINVAL : (e == cudaErrorMemoryAllocation ? ARIES_ENOMEM : ARIES_ECUDA);
    }
    aries_decoder* d = new (std::nothrow) aries_decoder{kMagicDec, ctx, plan};
    if (!d) {
        aries::decoder_plan_destroy(plan);
        return fail(ARIES_ENOMEM, "out of host memory");
    }
    *out = d;
    return ARIES_OK;
}

int aries_decoder_destroy(aries_decoder* dec) {
    if (!dec) return ARIES_OK;
    if (dec->magic != kMagicDec) return fail(ARIES_ESTATE, "invalid decoder handle");
    use(dec->ctx);
    aries::decoder_plan_destroy(dec->plan);
    dec->magic = 0;
    delete dec;
    return ARIES_OK;
}

namespace {
int generate_impl(aries_decoder* dec, const void* enc_out_dev, int batch, const int32_t* prompts, int prompt_len,
                  const aries_generate_opts* opts, const int32_t* forced, int n_forced, int32_t* tokens_out,
                  int32_t* lengths, int32_t* argmax_out, float* logits_out, float* scores, float* no_speech_prob,
                  void* stream) {
    if (!dec || dec->magic != kMagicDec) return fail(ARIES_ESTATE, "invalid decoder handle");
    int rc = use(dec->ctx);
    if (rc) return rc;
    if (!opts) return fail(ARIES_EINVAL, "aries_decoder_generate: opts is NULL");
    aries::GenerateOptsC o{};
    o.max_length = opts->max_length; o.suppress_blank = opts->suppress_blank; o.blank_id = opts->blank_id;
    o.eot = opts->eot; o.sot = opts->sot; o.no_speech = opts->no_speech; o.no_timestamps = opts->no_timestamps;
    o.timestamp_begin = opts->timestamp_begin; o.max_initial_timestamp_index = opts->max_initial_timestamp_index;
    o.suppress_tokens = opts->suppress_tokens; o.n_suppress = opts->n_suppress;
    o.forced = forced; o.n_forced = n_forced; o.argmax_out = argmax_out; o.logits_out = logits_out;
    cudaError_t e = aries::decoder_generate(dec->plan, enc_out_dev, batch, prompts, prompt_len, o, tokens_out, lengths,
                                            scores, no_speech_prob, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) {
        g_error = std::string("aries_decoder_generate: ") + aries::decoder_plan_error(dec->plan);
        cudaGetLastError();
        return e == cudaErrorInvalidValue ? ARIES_EINVAL : (e == cudaErrorMemoryAllocation ? ARIES_ENOMEM : ARIES_ECUDA);
    }
    return ARIES_OK;
}
}  // namespace

int aries_decoder_generate(aries_decoder* dec, const void* enc_out_dev, int batch, const int32_t* prompts, int prompt_len,
                           const aries_generate_opts* opts, int32_t* tokens_out, int32_t* lengths, float* scores,
                           float* no_speech_prob, void* stream) {
    return generate_impl(dec, enc_out_dev, batch, prompts, prompt_len, opts, nullptr, 0, tokens_out, lengths, nullptr,
                         nullptr, scores, no_speech_prob, stream);
}

int aries_decoder_detect_language(aries_decoder* dec, const void* enc_out_dev, int batch, const aries_generate_opts* opts,
                                  const int32_t* lang_ids, int n_lang, float* probs_out, void* stream) {
    if (!dec || dec->magic != kMagicDec) return fail(ARIES_ESTATE, "invalid decoder handle");
    int rc = use(dec->ctx);
    if (rc) return rc;
    if (!opts || !enc_out_dev) return fail(ARIES_EINVAL, "aries_decoder_detect_language: NULL argument");
    aries::GenerateOptsC o{};
    o.eot = opts->eot; o.sot = opts->sot; o.no_speech = opts->no_speech; o.no_timestamps = opts->no_timestamps;
    o.timestamp_begin = opts->timestamp_begin; o.blank_id = opts->blank_id;
    o.max_initial_timestamp_index = opts->max_initial_timestamp_index;
    cudaError_t e = aries::decoder_detect_language(dec->plan, enc_out_dev, batch, o, lang_ids, n_lang, probs_out,
                                                   static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) {
        g_error = std::string("aries_decoder_detect_language: ") + aries::decoder_plan_error(dec->plan);
        cudaGetLastError();
        return e == cudaErrorInvalidValue ? ARIES_EINVAL : (e == cudaErrorMemoryAllocation ? ARIES_ENOMEM : ARIES_ECUDA);
    }
    return ARIES_OK;
}

int aries_decoder_last_stats(const aries_decoder* dec, float* out, int n) {
    if (!dec || dec->magic != kMagicDec) return fail(ARIES_ESTATE, "invalid decoder handle");
    if (!out || n < 5) return fail(ARIES_EINVAL, "aries_decoder_last_stats: need n >= 5");
    aries::decoder_plan_last_stats(dec->plan, out);
    return ARIES_OK;
}

// ------------------------------------------------------------------------------------------------ test hooks
int aries_test_decoder_generate(aries_decoder* dec, const void* enc_out_dev, int batch, const int32_t* prompts,
                                int prompt_len, const aries_generate_opts* opts, const int32_t* forced, int n_forced,
                                int32_t* tokens_out, int32_t* argmax_out, float* logits_out, float* scores,
                                float* no_speech_prob, void* stream) {
    return generate_impl(dec, enc_out_dev, batch, prompts, prompt_len, opts, forced, n_forced, tokens_out, nullptr,
                         argmax_out, logits_out, scores, no_speech_prob, stream);
}

int aries_test_skinny_gemm(aries_ctx* ctx, int epi, int B, int N, int K, const void* x, const void* w, const float* bias,
                           void* out, int ldo, int splits, void* stream) {
    int rc = use(ctx);
    if (rc) return rc;
    if ((rc = ensure_kernels(ctx))) return rc;
    if (B <= 0 || B > 128 || N <= 0 || K <= 0 || K % 64 || epi < 0 || epi >= aries::SK_COUNT)
        return fail(ARIES_EINVAL, "aries_test_skinny_gemm: bad shape");
    const int NB = (B + 15) / 16 * 16;
    CUtensorMap tw, tx;
    const unsigned long long dw[2] = {(unsigned long long)K, (unsigned long long)N}, st[2] = {2, (unsigned long long)K * 2};
    const unsigned long long dx[2] = {(unsigned long long)K, (unsigned long long)NB};
    const unsigned bw[2] = {64, 128}, bx[2] = {64, (unsigned)NB};
    cudaError_t e;
    if ((e = aries::make_tmap_bf16(&tw, w, 2, dw, st, bw)) != cudaSuccess) return fail_cuda("tensor map W", e);
    if ((e = aries::make_tmap_bf16(&tx, x, 2, dx, st, bx)) != cudaSuccess) return fail_cuda("tensor map X", e);
    aries::SkinnyParams p{};
    p.B = B; p.NB = NB; p.N = N; p.K = K;
    p.splits = splits > 0 ? splits : aries::skinny_pick_splits(N, K, ctx->sm_count);
    p.bias = bias; p.out = out; p.ldo = ldo; p.pdl = 0;
    if ((e = aries::skinny_launch(epi, tw, tx, p, static_cast<cudaStream_t>(stream))) != cudaSuccess)
        return fail_cuda("skinny_launch", e);
    return ARIES_OK;
}

int aries_test_skinny_gemm_ln(aries_ctx* ctx, int epi, int B, int N, int K, const void* x_f16, const float* gamma,
                              const float* beta, const void* w, const float* bias, void* out, int ldo, int splits,
                              void* stream) {
    int rc = use(ctx);
    if (rc) return rc;
    if ((rc = ensure_kernels(ctx))) return rc;
    if (B <= 0 || B > 8 || N <= 0 || K <= 0 || K % 64 || epi < 0 || epi >= aries::SK_COUNT || !x_f16 || !gamma || !beta)
        return fail(ARIES_EINVAL, "aries_test_skinny_gemm_ln: bad shape");
    CUtensorMap tw;
    const unsigned long long dw[2] = {(unsigned long long)K, (unsigned long long)N}, st[2] = {2, (unsigned long long)K * 2};
    const unsigned bw[2] = {64, 128};
    cudaError_t e;
    if ((e = aries::make_tmap_bf16(&tw, w, 2, dw, st, bw)) != cudaSuccess) return fail_cuda("tensor map W", e);
    aries::SkinnyParams p{};
    p.B = B; p.NB = 16; p.N = N; p.K = K;
    p.splits = splits > 0 ? splits : aries::skinny_pick_splits_ln(N, K, ctx->sm_count);
    p.bias = bias; p.out = out; p.ldo = ldo; p.pdl = 0;
    p.ln_x = x_f16; p.ln_gamma = gamma; p.ln_beta = beta;
    if (p.splits < 1) return fail(ARIES_EINVAL, "aries_test_skinny_gemm_ln: K too deep for the fused variant");
    if ((e = aries::skinny_launch(epi, tw, tw, p, static_cast<cudaStream_t>(stream))) != cudaSuccess)
        return fail_cuda("skinny_launch (fused LayerNorm)", e);
    return ARIES_OK;
}

int aries_test_skinny_gemm_folded(aries_ctx* ctx, int B, int N, int K, const void* resid_x_f16, const void* ctx_in,
                                  const void* w_o, const float* bias_o, const void* w_folded_f16, const float* c1, const float* c2,
                                  int gelu, void* out_bf16, float* stats, void* stream) {
    int rc = use(ctx);
    if (rc) return rc;
    if ((rc = ensure_kernels(ctx))) return rc;
    if (B <= 0 || B > 128 || N <= 0 || K <= 0 || K % 64 || !resid_x_f16 || !ctx_in || !w_o || !bias_o || !w_folded_f16 || !c1 || !c2 ||
        !out_bf16 || !stats)
        return fail(ARIES_EINVAL, "aries_test_skinny_gemm_folded: bad arguments");
    const int parts = (K + 127) / 128;
    if (parts > 16 || aries::skinny_pick_splits(K, K, ctx->sm_count) < 2)
        return fail(ARIES_EINVAL, "aries_test_skinny_gemm_folded: needs K <= 2048 and a split residual GEMM");
    const int NB = (B + 15) / 16 * 16;
    CUtensorMap t_wo, t_ctx, t_wf, t_x;
    const unsigned long long st[2] = {2, (unsigned long long)K * 2};
    const unsigned long long d_wo[2] = {(unsigned long long)K, (unsigned long long)K}, d_x[2] = {(unsigned long long)K, (unsigned long long)NB};
    const unsigned long long d_wf[2] = {(unsigned long long)K, (unsigned long long)N};
    const unsigned bw[2] = {64, 128}, bx[2] = {64, (unsigned)NB};
    cudaError_t e;
    if ((e = aries::make_tmap_bf16(&t_wo, w_o, 2, d_wo, st, bw)) != cudaSuccess) return fail_cuda("tensor map", e);
    if ((e = aries::make_tmap_bf16(&t_ctx, ctx_in, 2, d_x, st, bx)) != cudaSuccess) return fail_cuda("tensor map", e);
    if ((e = aries::make_tmap_bf16(&t_wf, w_folded_f16, 2, d_wf, st, bw)) != cudaSuccess) return fail_cuda("tensor map", e);
    if ((e = aries::make_tmap_bf16(&t_x, resid_x_f16, 2, d_x, st, bx)) != cudaSuccess) return fail_cuda("tensor map", e);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // (1) the producing side: x += ctx W_o^T + b (f16 stream, in place) and the partial row sums of what it stored
    aries::SkinnyParams a{};
    a.B = B; a.NB = NB; a.N = K; a.K = K;
    a.splits = aries::skinny_pick_splits(K, K, ctx->sm_count);
    a.bias = bias_o; a.out = const_cast<void*>(resid_x_f16); a.ldo = K;
    a.stats_out = reinterpret_cast<float2*>(stats);
    if ((e = aries::skinny_launch(aries::SK_BIAS_RESID_F16, t_wo, t_ctx, a, s)) != cudaSuccess) return fail_cuda("skinny_launch (residual + statistics)", e);
    // (2) the consuming side: out = [gelu](rstd (x W'^T - mean c1) + c2)
    aries::SkinnyParams q{};
    q.B = B; q.NB = NB; q.N = N; q.K = K;
    q.splits = aries::skinny_pick_splits(N, K, ctx->sm_count);
    q.bias = c2; q.c1 = c1; q.stats_in = reinterpret_cast<const float2*>(stats); q.stats_parts = parts; q.ln_dim = K;
    q.out = out_bf16; q.ldo = N;
    if ((e = aries::skinny_launch(gelu ? aries::SK_LNF_GELU_BF16 : aries::SK_LNF_BF16, t_wf, t_x, q, s)) != cudaSuccess)
        return fail_cuda("skinny_launch (folded LayerNorm)", e);
    return ARIES_OK;
}

int aries_test_decode_attention(aries_ctx* ctx, const void* q, int q_ld, void* k, void* v, int64_t kv_rows, int kv_ld,
                                const void* new_k, const void* new_v, int new_ld, const int* step_dev, int n_keys_fixed,
                                int batch, int heads, void* out, int out_ld, int splits, void* stream) {
    int rc = use(ctx);
    if (rc) return rc;
    if (batch <= 0 || heads <= 0 || splits < 1 || splits > 8) return fail(ARIES_EINVAL, "aries_test_decode_attention: bad shape");
    aries::DecAttnParams p{};
    p.batch = batch; p.heads = heads; p.q = q; p.q_ld = q_ld; p.k = k; p.v = v; p.kv_rows = kv_rows; p.kv_ld = kv_ld;
    p.new_k = new_k; p.new_v = new_v; p.new_ld = new_ld; p.step = step_dev; p.n_keys_fixed = n_keys_fixed;
    p.out = out; p.out_ld = out_ld; p.splits = splits; p.pdl = 0;
    cudaError_t e = aries::decode_attention_launch(p, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail_cuda("decode_attention_launch", e);
    return ARIES_OK;
}

int aries_test_gemm(aries_ctx* ctx, int epi, int M, int N, int K, const void* a, const void* b, const float* bias,
                    const void* resid, const float* pos, int pos_rows, void* out, void* out2, int n_split,
                    int t_rows, int t_pad, void* stream) {
    int rc = use(ctx);
    if (rc) return rc;
    if ((rc = ensure_kernels(ctx))) return rc;
    if (M <= 0 || N % 128 || K % 64 || epi < 0 || epi >= aries::EPI_COUNT) return fail(ARIES_EINVAL, "aries_test_gemm: bad shape");
    CUtensorMap ta, tb;
    const unsigned long long da[2] = {(unsigned long long)K, (unsigned long long)M}, sa[2] = {2, (unsigned long long)K * 2};
    const unsigned long long db[2] = {(unsigned long long)K, (unsigned long long)N};
    const unsigned ba[2] = {64, 128}, bb[2] = {64, (unsigned)aries::gemm_b_box_rows()};
    cudaError_t e;
    if ((e = aries::make_tmap_bf16(&ta, a, 2, da, sa, ba)) != cudaSuccess) return fail_cuda("tensor map A", e);
    if ((e = aries::make_tmap_bf16(&tb, b, 2, db, sa, bb)) != cudaSuccess) return fail_cuda("tensor map B", e);
    aries::GemmParams p{};
    p.M = M; p.N = N; p.K = K; p.a_cols = K;
    p.p_in = M; p.t_valid = M; p.p_out = M; p.row_off = 0; p.ldo = N;
    p.bias = bias; p.resid = resid; p.pos = pos; p.out = out;
    if (epi == aries::EPI_BIAS_GELU_POS_F16) { p.p_in = pos_rows; p.t_valid = pos_rows; p.p_out = pos_rows; }
    if (epi == aries::EPI_QKV_SPLIT_BF16) {
        p.p_in = t_rows; p.t_valid = t_rows; p.p_out = t_rows; p.ldo = n_split;
        p.out2 = out2; p.n_split = n_split; p.t_pad = t_pad;
    }
    if ((e = aries::gemm_launch(epi, ta, tb, p, ctx->sm_count, static_cast<cudaStream_t>(stream))) != cudaSuccess)
        return fail_cuda("gemm_launch", e);
    return ARIES_OK;
}

int aries_test_gemm_ln(aries_ctx* ctx, int epi, int M, int N, int K, const void* a, const void* b, const float* bias,
                       const float* c1, const void* stats_in, int stats_parts, int ln_dim, const void* resid, void* out,
                       void* out2, int n_split, int t_rows, int t_pad, void* stats_out, void* stream) {
    int rc = use(ctx);
    if (rc) return rc;
    if ((rc = ensure_kernels(ctx))) return rc;
    if (M <= 0 || N % 128 || K % 64 || epi < 0 || epi >= aries::EPI_COUNT) return fail(ARIES_EINVAL, "aries_test_gemm_ln: bad shape");
    CUtensorMap ta, tb;
    const unsigned long long da[2] = {(unsigned long long)K, (unsigned long long)M}, sa[2] = {2, (unsigned long long)K * 2};
    const unsigned long long db[2] = {(unsigned long long)K, (unsigned long long)N};
    const unsigned ba[2] = {64, 128}, bb[2] = {64, (unsigned)aries::gemm_b_box_rows()};
    cudaError_t e;
    if ((e = aries::make_tmap_bf16(&ta, a, 2, da, sa, ba)) != cudaSuccess) return fail_cuda("tensor map A", e);
    if ((e = aries::make_tmap_bf16(&tb, b, 2, db, sa, bb)) != cudaSuccess) return fail_cuda("tensor map B", e);
    aries::GemmParams p{};
    p.M = M; p.N = N; p.K = K; p.a_cols = K;
    p.p_in = M; p.t_valid = M; p.p_out = M; p.row_off = 0; p.ldo = N;
    p.bias = bias; p.resid = resid; p.out = out;
    p.c1 = c1; p.stats_in = static_cast<const float2*>(stats_in); p.stats_parts = stats_parts; p.ln_dim = ln_dim;
    p.ln_eps = 1e-5f; p.stats_out = static_cast<float2*>(stats_out);
    if (const char* tr = getenv("ARIES_GEMM_TRACE")) {            // test hook: device u64 [tiles][8], see gemm.h
        p.trace = reinterpret_cast<unsigned long long*>(strtoull(tr, nullptr, 0));
        p.trace_tiles = 48;
    }
    if (epi == aries::EPI_LN_QKV_SPLIT_BF16) {
        p.p_in = t_rows; p.t_valid = t_rows; p.p_out = t_rows; p.ldo = n_split;
        p.out2 = out2; p.n_split = n_split; p.t_pad = t_pad;
    }
    if ((e = aries::gemm_launch(epi, ta, tb, p, ctx->sm_count, static_cast<cudaStream_t>(stream))) != cudaSuccess)
        return fail_cuda("gemm_launch", e);
    return ARIES_OK;
}

int aries_test_gemm_stats_parts(int N) { return aries::gemm_stats_parts(N); }

int aries_test_layernorm(aries_ctx* ctx, const void* x, const float* gamma, const float* beta, void* y, int64_t rows,
                         int d, void* stream) {
    int rc = use(ctx);
    if (rc) return rc;
    cudaError_t e = aries::layernorm_launch(x, gamma, beta, y, rows, d, 1e-5f, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail_cuda("layernorm_launch", e);
    return ARIES_OK;
}

int aries_test_attention(aries_ctx* ctx, const void* qk, const void* vt, int batch, int T, int n_heads, int t_pad,
                         void* out, void* stream) {
    int rc = use(ctx);
    if (rc) return rc;
    if ((rc = ensure_kernels(ctx))) return rc;
    aries::AttnMaps maps;
    cudaError_t e = aries::attention_make_maps(qk, vt, batch, T, n_heads * 64, n_heads, t_pad, &maps);
    if (e != cudaSuccess) return fail_cuda("attention tensor maps", e);
    aries::AttnParams p{batch, T, n_heads * 64, n_heads, out};
    if ((e = aries::attention_launch(maps, p, static_cast<cudaStream_t>(stream))) != cudaSuccess)
        return fail_cuda("attention_launch", e);
    return ARIES_OK;
}

int aries_test_attention_trace(aries_ctx* ctx, unsigned long long* host, size_t count) {
    int rc = use(ctx);
    if (rc) return rc;
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = aries::attention_read_trace(host, count);
    if (e != cudaSuccess) return fail_cuda("attention_read_trace", e);
    return ARIES_OK;
}

}  // extern "C"
