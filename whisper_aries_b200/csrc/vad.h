// Host-side interface of the VAD front end (vad.cu).
#pragma once
#include <cuda_runtime.h>

#include <vector>

namespace aries {

struct VadOpts {                       // faster-whisper 1.1.1 VadOptions
    double threshold = 0.5;
    double neg_threshold = -1.0;       // < 0: upstream's None -> max(threshold - 0.15, 0.01)
    int min_speech_duration_ms = 0;
    double max_speech_duration_s = 0;  // <= 0 or inf: unlimited
    int min_silence_duration_ms = 2000;
    int speech_pad_ms = 400;
};

long long vad_num_windows(long long n_samples);
// probs: host f32 [n_windows]; fills sample ranges [start, end) of the speech chunks, padded as upstream does.
void vad_speech_timestamps(const float* probs, long long n_windows, long long audio_len, const VadOpts& opts,
                           std::vector<long long>* starts, std::vector<long long>* ends);
// Stand-in probability model (log-energy logistic), device f32 [n_samples] -> device f32 [vad_num_windows(n_samples)].
cudaError_t vad_energy_probs(const float* pcm, long long n_samples, float center_db, float width_db, float* probs,
                             cudaStream_t stream);
// out[offs[k] + j] = pcm[starts[k] + j]; d_starts / d_offs: device i64 [n_chunks]; total = sum of chunk lengths.
cudaError_t collect_chunks(const float* pcm, const long long* d_starts, const long long* d_offs, int n_chunks,
                           long long total, float* out, int sm_count, cudaStream_t stream);

}  // namespace aries
