// Host-side interface of the Whisper text decoder + greedy generate assembly (decoder.cu) -- row f1.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "encoder.h"   // WeightView

namespace aries {

struct DecoderShapeC {
    int vocab, d_model, n_heads, n_layers, d_ffn, n_text_ctx, n_audio_ctx;
};

struct GenerateOptsC {
    int max_length;                     // total positions (prompt included), <= n_text_ctx
    int suppress_blank, blank_id;
    int eot, sot, no_speech, no_timestamps, timestamp_begin;
    int max_initial_timestamp_index;
    const int* suppress_tokens;         // host
    int n_suppress;
    const int* forced;                  // tests: host [batch, n_forced] continuation to force (NULL in production)
    int n_forced;
    int* argmax_out;                    // tests: host [batch, max_length] raw argmax per sampled position (or NULL)
    float* logits_out;                  // tests: host [max_length - 1, batch, vocab] logits of every step (or NULL)
};

struct DecoderPlan;

cudaError_t decoder_plan_create(int device, int sm_count, const DecoderShapeC& cfg, const WeightView* weights,
                                int n_weights, int max_batch, DecoderPlan** out, std::string* why);
void decoder_plan_destroy(DecoderPlan* pl);
const char* decoder_plan_error(const DecoderPlan* pl);
const DecoderShapeC* decoder_plan_cfg(const DecoderPlan* pl);
int decoder_plan_max_batch(const DecoderPlan* pl);

// enc_out: device bf16 [batch, n_audio_ctx, d_model] (what aries_encoder_run wrote).  prompts: host [batch, prompt_len].
// tokens_out: host [batch, max_length] (prompt, then the sampled ids, EOT-filled once a sequence has ended);
// lengths: host [batch] sampled ids before EOT; scores / no_speech_prob: host [batch].  Synchronises `stream`.
cudaError_t decoder_generate(DecoderPlan* pl, const void* enc_out, int batch, const int* prompts, int prompt_len,
                             const GenerateOptsC& opts, int* tokens_out, int* lengths, float* scores,
                             float* no_speech_prob, cudaStream_t stream);

// ctranslate2 Whisper.detect_language: one decoder step on <|startoftranscript|>, softmax over the language-token logits.
// probs: host [batch, n_lang] (order of lang_ids).  Synchronises.
cudaError_t decoder_detect_language(DecoderPlan* pl, const void* enc_out, int batch, const GenerateOptsC& ids,
                                    const int* lang_ids, int n_lang, float* probs, cudaStream_t stream);

// Timing / launch counts of the last decoder_generate: [0] cross-KV projection ms, [1] decode loop ms, [2] steps run,
// [3] kernels per step, [4] kernels of the cross-KV phase.
void decoder_plan_last_stats(const DecoderPlan* pl, float out[5]);

}  // namespace aries
