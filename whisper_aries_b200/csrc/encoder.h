// Host-side interface of the encoder assembly (encoder.cu).
#pragma once
#include <cuda_runtime.h>

#include <string>

namespace aries {

struct EncoderShapeC {
    int n_mels, d_model, n_heads, n_layers, d_ffn, n_ctx;
};

struct WeightView {
    const char* name;
    const float* data;
    int ndim;
    long long shape[4];
};

struct EncoderPlan;
class Profiler;

cudaError_t encoder_plan_create(int device, int sm_count, const EncoderShapeC& cfg, const WeightView* weights,
                                int n_weights, EncoderPlan** out, std::string* why);
void encoder_plan_destroy(EncoderPlan* pl);
size_t encoder_workspace_bytes(const EncoderPlan* pl, int batch);
// mel: device f32 [batch, n_mels, frames], or nullptr when the conv1 operand in the workspace was written directly
// (fused PCM path); out: device bf16 [batch, 1500, d_model].  Stream-ordered.
cudaError_t encoder_run(EncoderPlan* pl, const float* mel, int batch, int frames, void* out_bf16, void* workspace,
                        size_t ws_bytes, cudaStream_t stream);
// The conv1 operand inside the workspace: bf16 time-major [batch, 3002, c_pad]; the fused PCM path has the log-mel
// kernel write rows 1..3000 of it (rows 0 and 3001 are zeroed by encoder_run).
void* encoder_workspace_conv1_operand(const EncoderPlan* pl, void* workspace, int batch, int* c_pad);
const char* encoder_plan_error(const EncoderPlan* pl);
int encoder_plan_last_launches(const EncoderPlan* pl);
const EncoderShapeC* encoder_plan_cfg(const EncoderPlan* pl);
Profiler* encoder_plan_profiler(EncoderPlan* pl);

}  // namespace aries
