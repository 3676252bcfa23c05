/* aries_b200.h — C ABI of the B200-native log-mel + Whisper-encoder path (and, widened, the greedy decoder).
 *
 * Drop-in boundary for the hot path Whisper-Aries obtains from faster-whisper 1.1.1 / CTranslate2 4.6.0
 * (requirements.txt:12,9), which the reference enters at final_optimized_transcriber.py:326
 * (model.transcribe(chunk_audio, ...); warm-up at :189; model built at :179-182; large-v3 forced at
 * conversation_transcriber.py:72-77).  Upstream, that call runs
 *     FeatureExtractor.__call__(waveform, padding=160)            -> aries_logmel_run / aries_logmel_run_host
 *     WhisperModel.encode(features) -> ctranslate2 Whisper.encode -> aries_encoder_run / aries_encoder_run_host
 *     ctranslate2 Whisper.detect_language(encoder_output)         -> aries_decoder_detect_language      
 *     ctranslate2 Whisper.generate(encoder_output, prompts, ...)  -> aries_decoder_generate             (row f1)
 * per 30-second window.  Each entry point below names the upstream interface it replaces.
 *
 * Conventions
 *   - plain C types only; every handle is opaque; no exceptions cross the boundary.
 *   - return value: 0 (ARIES_OK) or a negative ARIES_E* code; aries_last_error() gives the message of the last
 *     failure on the calling thread.  The Python shim maps ARIES_EINVAL -> ValueError (upstream raises
 *     ValueError("Invalid input features shape...") from std::invalid_argument), everything else -> RuntimeError.
 *   - inputs are borrowed for the duration of the call; outputs are caller-allocated; weights are copied at create.
 *   - "dev" pointers are device memory on the context's GPU; the *_host variants take host memory (pageable or
 *     pinned), do the copies on the context's stream and return after synchronising.
 *   - stream arguments are cudaStream_t passed as void* (NULL = the legacy default stream).  Device-pointer entry
 *     points are stream-ordered and do not synchronise the host.
 *   - threading: one in-flight call per handle; distinct handles (also on different GPUs) are independent — the
 *     reference's model of one model object per worker thread (final_optimized_transcriber.py:256-262).
 *   - there is no CPU fallback: every entry point fails with ARIES_ECUDA when no sm_100 device is usable.
 */
#ifndef ARIES_B200_H_
#define ARIES_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ARIES_API __attribute__((visibility("default")))
#else
#define ARIES_API
#endif

#define ARIES_OK 0
#define ARIES_EINVAL (-1)   /* bad argument / shape (-> ValueError) */
#define ARIES_ECUDA (-2)    /* CUDA runtime or driver failure (-> RuntimeError) */
#define ARIES_ENOMEM (-3)   /* allocation failure */
#define ARIES_ESTATE (-4)   /* handle used on the wrong context / after destroy */

typedef struct aries_ctx aries_ctx;
typedef struct aries_mel aries_mel;
typedef struct aries_encoder aries_encoder;

/* ABI version of this header (major * 100 + minor). */
ARIES_API int aries_abi_version(void);
/* Message of the last failed call on this thread ("" if none). Never NULL. */
ARIES_API const char* aries_last_error(void);

/* One context per GPU. Replaces: the device selection inside WhisperModel(device=..., device_index=...)
 * (final_optimized_transcriber.py:179-182; round-robin placement in "Yasmeen's code/complete_fixed_whisper.py":179-184). */
ARIES_API int aries_init(int device, aries_ctx** out);
ARIES_API int aries_destroy(aries_ctx* ctx);
ARIES_API int aries_device(const aries_ctx* ctx);
ARIES_API int aries_sm_count(const aries_ctx* ctx);

/* ---------------------------------------------------------------------------------------------- log-mel
 * Replaces faster_whisper.feature_extractor.FeatureExtractor (n_fft 400, hop 160, sampling rate 16 kHz).
 * mel_filters: host f32 [n_mels, 201] exactly as FeatureExtractor.get_mel_filters returns it. */
ARIES_API int aries_logmel_create(aries_ctx* ctx, int n_mels, const float* mel_filters, aries_mel** out);
ARIES_API int aries_logmel_destroy(aries_mel* mel);
/* (n_samples + padding) / 160: the frame count FeatureExtractor.__call__ returns (last STFT frame dropped). */
ARIES_API int64_t aries_logmel_num_frames(int64_t n_samples, int padding);

/* FeatureExtractor.__call__(waveform, padding) for `batch` equally long signals.
 *   pcm_dev    device f32, signal i starts at pcm_dev + i * pcm_stride (floats)
 *   out_dev    device f32 [batch, n_mels, frames_out]
 *   frames_out frames stored per signal: pass aries_logmel_num_frames() for the upstream result, or 3000 for
 *              the window WhisperModel.encode consumes (pad_or_trim: zero-filled past the last real frame).
 *              The clamp maximum is always taken over ALL frames of the call, as upstream does. */
ARIES_API int aries_logmel_run(aries_mel* mel, const float* pcm_dev, int batch, int64_t n_samples, int64_t pcm_stride,
                     int padding, float* out_dev, int frames_out, void* stream);
/* Same, host buffers (pcm_host contiguous [batch, n_samples], out_host [batch, n_mels, frames_out]). */
ARIES_API int aries_logmel_run_host(aries_mel* mel, const float* pcm_host, int batch, int64_t n_samples, int padding,
                          float* out_host, int frames_out);

/* ---------------------------------------------------------------------------------------------- encoder
 * Replaces ctranslate2.models.Whisper.encode (layers::WhisperEncoder): conv1(k3,s1,p1)+GELU, conv2(k3,s2,p1)+GELU,
 * + position encodings, n_layers pre-norm blocks (MHA with q scaled by head_dim^-0.5 and no key bias, GELU MLP),
 * final LayerNorm (eps 1e-5).  bf16 tensor-core arithmetic, f32 accumulation / LN statistics / softmax, f16 residual stream. */
typedef struct aries_encoder_cfg {
    int32_t n_mels;     /* 80 | 128 */
    int32_t d_model;    /* multiple of 128 */
    int32_t n_heads;    /* d_model / 64 */
    int32_t n_layers;
    int32_t d_ffn;      /* multiple of 128 */
    int32_t n_ctx;      /* 1500 */
} aries_encoder_cfg;

/* One tensor of the CTranslate2 Whisper model ("encoder/conv1/weight", "encoder/layer_0/self_attention/linear_0/weight",
 * ...), host f32, C-contiguous. */
typedef struct aries_weight_desc {
    const char* name;
    const float* data;
    int32_t ndim;
    int64_t shape[4];
} aries_weight_desc;

ARIES_API int aries_encoder_create(aries_ctx* ctx, const aries_encoder_cfg* cfg, const aries_weight_desc* weights,
                         int n_weights, aries_encoder** out);
ARIES_API int aries_encoder_destroy(aries_encoder* enc);
/* Bytes of device scratch aries_encoder_run needs for `batch` windows (0 on error). */
ARIES_API size_t aries_encoder_workspace_bytes(const aries_encoder* enc, int batch);

/* Whisper.encode(features).
 *   mel_dev    device f32 [batch, n_mels, frames], frames <= 3000 (shorter inputs are zero-padded like CT2 pads)
 *   out_dev    device bf16 [batch, 1500, d_model]
 *   workspace  device scratch, 256-byte aligned, >= aries_encoder_workspace_bytes(enc, batch) */
ARIES_API int aries_encoder_run(aries_encoder* enc, const float* mel_dev, int batch, int frames, void* out_dev,
                      void* workspace, size_t workspace_bytes, void* stream);
/* Same with host buffers: mel_host f32 [batch, n_mels, frames] -> out_host bf16 (raw uint16) [batch, 1500, d_model];
 * device scratch is owned (and grown on demand) by the handle. */
ARIES_API int aries_encoder_run_host(aries_encoder* enc, const float* mel_host, int batch, int frames, uint16_t* out_host);

/* Fused front end + encoder (additive extension): PCM -> log-mel -> encoder without the mel tensor leaving the GPU.
 * Equivalent to encode(pad_or_trim(feature_extractor(pcm))[:, :3000]) for 30-second windows (n_samples <= 480000). */
ARIES_API int aries_encode_pcm(aries_encoder* enc, aries_mel* mel, const float* pcm_dev, int batch, int64_t n_samples,
                     int64_t pcm_stride, void* out_dev, void* workspace, size_t workspace_bytes, void* stream);

/* PCM staging (row f3, additive): int16 samples (what ffmpeg's pcm_s16le decode yields, ref: utils.py:116) -> float32 / 32768,
 * the conversion faster-whisper's decode_audio does on the host, here on the device so that the host->device copy of the hot
 * path carries 2 bytes per sample.  pcm_s16_dev / out_dev: device pointers, n samples; stream-ordered. */
ARIES_API int aries_pcm_s16_to_f32(aries_ctx* ctx, const int16_t* pcm_s16_dev, float* out_dev, int64_t n, void* stream);

/* Number of kernels the last aries_logmel_run / aries_encoder_run / aries_encode_pcm on this handle launched. */
ARIES_API int aries_logmel_last_launches(const aries_mel* mel);
ARIES_API int aries_encoder_last_launches(const aries_encoder* enc);

/* Per-kernel timing for the benchmark's roofline report (no upstream counterpart). When on, every kernel launched by
 * aries_encoder_run / aries_encode_pcm is bracketed by CUDA events on the launching stream.
 * aries_encoder_collect_profile synchronises on them and ADDS, per kernel class, the elapsed milliseconds to ms[i]
 * and the number of launches to counts[i]; n must be >= ARIES_KERNEL_CLASSES. Class order: log-mel tiles,
 * log-mel clamp, mel transpose, conv1, conv2, layer norm, qkv projection, attention, output projection, fc1, fc2. */
#define ARIES_KERNEL_CLASSES 11
ARIES_API int aries_encoder_set_profiling(aries_encoder* enc, int on);
ARIES_API int aries_encoder_collect_profile(aries_encoder* enc, float* ms, int* counts, int n);

/* ---------------------------------------------------------------------------------------------- decoder (row f1)
 * Replaces ctranslate2.models.Whisper.generate(encoder_output, prompts, beam_size=1, max_length=448, suppress_blank=True,
 * suppress_tokens=[...], max_initial_timestamp_index=50, return_scores=True, return_no_speech_prob=True) as faster-whisper
 * calls it for the reference's decode knobs beam_size=1 / best_of=1 / temperature=0 (final_optimized_transcriber.py:432-441,
 * reached from model.transcribe at :326): layers::WhisperDecoder + GreedySearch with the logits processors
 * SuppressTokensBegin, SuppressTokens and ApplyTimestampRules.  Greedy only (the reference never asks for a beam).
 * The encoder output stays on the GPU: enc_out_dev is what aries_encoder_run / aries_encode_pcm wrote. */
typedef struct aries_decoder aries_decoder;

typedef struct aries_decoder_cfg {
    int32_t vocab;        /* 51866 (large-v3) | 51865 */
    int32_t d_model;      /* multiple of 64 */
    int32_t n_heads;      /* d_model / 64 */
    int32_t n_layers;
    int32_t d_ffn;        /* multiple of 64 */
    int32_t n_text_ctx;   /* 448 */
    int32_t n_audio_ctx;  /* 1500 */
} aries_decoder_cfg;

/* Token ids come from the converted model's tokenizer (the caller owns the tokenizer, as upstream does). */
typedef struct aries_generate_opts {
    int32_t max_length;                    /* total positions, prompt included; <= n_text_ctx */
    int32_t suppress_blank;                /* forbid " " and EOT at the first sampled position */
    int32_t blank_id;
    int32_t eot, sot, no_speech, no_timestamps, timestamp_begin;
    int32_t max_initial_timestamp_index;   /* 50 = 1.0 s */
    const int32_t* suppress_tokens;        /* ids that are never sampled (upstream's -1 already expanded) */
    int32_t n_suppress;
} aries_generate_opts;

/* Weights: CTranslate2 Whisper variables "decoder/embeddings/weight", "decoder/position_encodings/encodings",
 * "decoder/layer_N/{self_attention,attention,ffn}/...", "decoder/layer_norm/{gamma,beta}" and optionally
 * "decoder/projection/weight" (tied to the embedding when absent); host f32, copied (as bf16 / f32) at create.
 * max_batch (<= 128) sizes the self-attention cache: n_layers * max_batch * n_text_ctx * d_model * 4 bytes. */
ARIES_API int aries_decoder_create(aries_ctx* ctx, const aries_decoder_cfg* cfg, const aries_weight_desc* weights,
                         int n_weights, int max_batch, aries_decoder** out);
ARIES_API int aries_decoder_destroy(aries_decoder* dec);

/* Whisper.generate for `batch` windows.
 *   enc_out_dev     device bf16 [batch, n_audio_ctx, d_model]
 *   prompts         host int32 [batch, prompt_len], e.g. <|startoftranscript|> <|en|> <|transcribe|>; the timestamp
 *                   rules apply to a sequence unless its prompt contains <|notimestamps|>
 *   tokens_out      host int32 [batch, max_length]: the prompt, the sampled ids, then EOT
 *   lengths         host int32 [batch]: sampled ids before EOT (upstream's sequences_ids[0] length)   (may be NULL)
 *   scores          host f32 [batch]: sum of the log-probabilities of the sampled ids                  (may be NULL)
 *   no_speech_prob  host f32 [batch]: P(<|nospeech|>) at the <|startoftranscript|> position            (may be NULL)
 * Ordered after the work already queued on `stream`; returns after the results are in host memory. */
ARIES_API int aries_decoder_generate(aries_decoder* dec, const void* enc_out_dev, int batch, const int32_t* prompts,
                           int prompt_len, const aries_generate_opts* opts, int32_t* tokens_out, int32_t* lengths,
                           float* scores, float* no_speech_prob, void* stream);

/* ctranslate2.models.Whisper.detect_language(encoder_output), which faster-whisper calls when language=None -- the
 * reference's default ("auto" -> None, final_optimized_transcriber.py:433; result recorded at :350-351): one decoder step
 * on <|startoftranscript|> and a softmax over the language-token logits.
 *   lang_ids   host int32 [n_lang]: the language token ids (large-v3: 50259 .. 50358)
 *   probs_out  host f32 [batch, n_lang], in the order of lang_ids (upstream sorts them by probability: the shim does) */
ARIES_API int aries_decoder_detect_language(aries_decoder* dec, const void* enc_out_dev, int batch,
                                  const aries_generate_opts* opts, const int32_t* lang_ids, int n_lang, float* probs_out,
                                  void* stream);

/* Timing of the last aries_decoder_generate (CUDA events on the decoding stream), n >= 5:
 * out[0] cross-attention K|V projection ms, out[1] decode loop ms, out[2] steps run, out[3] kernels per step,
 * out[4] kernels of the K|V projection phase. */
ARIES_API int aries_decoder_last_stats(const aries_decoder* dec, float* out, int n);

/* ------------------------------------------------------------------------------------------------------------------
 * Row f4: the voice-activity filter in front of the feature extractor.  The reference always passes
 * vad_filter=True (final_optimized_transcriber.py:440, whitelisted at :318), so upstream's transcribe() runs
 *     speech_chunks = get_speech_timestamps(audio, VadOptions())     -> aries_vad_speech_timestamps
 *     audio = np.concatenate(collect_chunks(audio, speech_chunks))   -> aries_collect_chunks
 * before FeatureExtractor.__call__.  The per-window speech probabilities come from upstream's Silero network, whose
 * trained weights ship inside the faster-whisper wheel and are not available offline: they are an INPUT here (any model
 * plugs in); aries_vad_energy_probs is a labelled stand-in (log-energy logistic), not Silero. */
typedef struct aries_vad_opts {      /* faster_whisper.vad.VadOptions (1.1.1) */
    float threshold;                 /* 0.5 */
    float neg_threshold;             /* < 0 = None -> max(threshold - 0.15, 0.01) */
    int32_t min_speech_duration_ms;  /* 0 */
    float max_speech_duration_s;     /* <= 0 or inf = unlimited */
    int32_t min_silence_duration_ms; /* 2000 */
    int32_t speech_pad_ms;           /* 400 */
} aries_vad_opts;

/* Number of 512-sample windows upstream evaluates for n_samples of audio (it pads with 512 - n % 512 samples). */
ARIES_API int64_t aries_vad_num_windows(int64_t n_samples);

/* Replaces: the state machine + padding pass of faster_whisper.vad.get_speech_timestamps.  probs_host: f32
 * [n_windows] speech probabilities; writes up to `cap` chunks as sample ranges [starts[k], ends[k]) and their count to
 * *n_out (ARIES_EINVAL if cap is too small; *n_out then holds the needed count).  Host-only: needs no GPU. */
ARIES_API int aries_vad_speech_timestamps(const float* probs_host, int64_t n_windows, int64_t audio_len,
                                          const aries_vad_opts* opts, int64_t* starts, int64_t* ends, int cap, int* n_out);

/* Stand-in probability model on the device (NOT Silero): p = sigmoid((20 log10(rms of the window + 1e-10) - center_db)
 * / width_db); pcm_dev f32 [n_samples] -> probs_dev f32 [aries_vad_num_windows(n_samples)].  Stream-ordered. */
ARIES_API int aries_vad_energy_probs(aries_ctx* ctx, const float* pcm_dev, int64_t n_samples, float center_db,
                                     float width_db, float* probs_dev, void* stream);

/* Replaces: faster_whisper.vad.collect_chunks + np.concatenate (one device gather instead of a host copy per chunk).
 * out_dev f32 [sum(ends - starts)] (capacity out_cap samples); *out_len = samples written.  Stream-ordered. */
ARIES_API int aries_collect_chunks(aries_ctx* ctx, const float* pcm_dev, int64_t n_samples, const int64_t* starts_host,
                                   const int64_t* ends_host, int n_chunks, float* out_dev, int64_t out_cap,
                                   int64_t* out_len, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ARIES_B200_H_ */
