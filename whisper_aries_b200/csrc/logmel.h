// Host-side interface of the log-mel kernels (logmel.cu).
#pragma once
#include <cuda_runtime.h>

namespace aries {

namespace mel { struct Tables; }

struct LogmelPlan;
class Profiler;

// filters: host f32 [n_mels, 201] (Slaney bank as FeatureExtractor.get_mel_filters builds it).
cudaError_t logmel_plan_create(int device, int sm_count, int n_mels, const float* filters, LogmelPlan** out,
                               const char** why);
void logmel_plan_destroy(LogmelPlan* pl);
int logmel_plan_n_mels(const LogmelPlan* pl);

// pcm: device f32, `batch` signals of n_samples each, pcm_stride floats apart.  out: device f32
// [batch, n_mels, frames_out]; frames beyond (n_samples + padding) / 160 are zero-filled, frames beyond frames_out
// are computed (they count for the clamp maximum) but not stored.  Stream-ordered; no host sync in steady state.
cudaError_t logmel_run(LogmelPlan* pl, const float* pcm, int batch, long long n_samples, long long pcm_stride,
                       int padding, float* out, int frames_out, cudaStream_t stream, int* launches,
                       Profiler* prof = nullptr);

// Fused variant for aries_encode_pcm: ONE launch that writes the conv1 operand directly -- bf16, time-major
// [batch, frames_out + 2, c_pad] (rows 1 .. frames_out; channels >= n_mels zero; rows 0 and frames_out + 1 are the
// caller's) -- clamp included (the CTA that finishes a window's last tile fixes that window's flagged tiles in L2).
cudaError_t logmel_run_time_major(LogmelPlan* pl, const float* pcm, int batch, long long n_samples, long long pcm_stride,
                                  int padding, void* out_tm_bf16, int frames_out, int c_pad, cudaStream_t stream,
                                  int* launches, Profiler* prof = nullptr);

// [B, n_mels, frames] f32 -> [B, 3002, c_pad] bf16 rows 1..3000 (time-major conv1 operand); pad rows untouched.
cudaError_t mel_to_time_major(const float* mel, int batch, int n_mels, int frames, void* out_bf16, int c_pad,
                              cudaStream_t stream);

void logmel_host_tables(mel::Tables* tb);

// pcm.cu: s16 PCM -> f32 / 32768 on the device (row f3)
cudaError_t pcm_s16_to_f32(const short* in, float* out, long long n, int sm_count, cudaStream_t stream);

}  // namespace aries
