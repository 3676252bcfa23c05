// Row f4: what `vad_filter=True` does BEFORE the feature extractor.
//
// The reference always transcribes with vad_filter=True (ref: final_optimized_transcriber.py:440, whitelisted at :318),
// so faster-whisper 1.1.1 `transcribe` runs  speech_probs = SileroVAD(audio)  ->  get_speech_timestamps state machine
// -> collect_chunks  and only the concatenated speech reaches FeatureExtractor.  Built here:
//   * vad_speech_timestamps: the state machine + padding pass of `get_speech_timestamps` (host C++, sequential by
//     nature: 31 windows per second of audio), bit-exact on segment boundaries with the oracle (oracle/vad.py);
//   * collect_chunks_kernel: the concatenation as ONE device gather (HBM-bound: 4 B read + 4 B written per kept
//     sample), so filtered PCM goes straight into the log-mel kernel without visiting the host;
//   * vad_energy_probs_kernel: a stand-in speech-probability model (log-energy through a logistic) -- NOT upstream's
//     Silero network, whose trained weights ship inside the faster-whisper wheel and do not exist offline; any
//     callable producing per-window probabilities plugs in instead (whisper_aries_b200/vad.py).
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <limits>
#include <vector>

#include "vad.h"

namespace aries {

namespace {

constexpr int kWindow = 512;

// one warp per 512-sample window of the zero-padded signal: 16 samples per lane, shuffle reduction
__global__ void __launch_bounds__(256) vad_energy_probs_kernel(const float* __restrict__ pcm, long long n_samples,
                                                               long long n_windows, float center_db, float inv_width_db,
                                                               float* __restrict__ probs) {
    const int lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= n_windows) return;
    const long long base = w * kWindow;
    float acc = 0.0f;
#pragma unroll
    for (int k = 0; k < kWindow / 32; ++k) {
        const long long i = base + lane + 32 * k;
        const float v = i < n_samples ? __ldg(pcm + i) : 0.0f;
        acc = fmaf(v, v, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        const float rms = sqrtf(acc * (1.0f / kWindow));
        const float db = 20.0f * log10f(rms + 1e-10f);
        probs[w] = 1.0f / (1.0f + expf(-(db - center_db) * inv_width_db));
    }
}

// out[i] = pcm[start[k] + (i - off[k])] for the chunk k with off[k] <= i < off[k + 1]
__global__ void __launch_bounds__(256) collect_chunks_kernel(const float* __restrict__ pcm, const long long* __restrict__ starts,
                                                             const long long* __restrict__ offs, int n_chunks,
                                                             long long total, float* __restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        int lo = 0, hi = n_chunks - 1;               // last chunk whose offset is <= i
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (offs[mid] <= i) lo = mid; else hi = mid - 1;
        }
        out[i] = __ldg(pcm + starts[lo] + (i - offs[lo]));
    }
}

}  // namespace

long long vad_num_windows(long long n_samples) {
    // upstream pads with `window - len % window` samples: a FULL extra window when len is a multiple of 512
    return (n_samples + (kWindow - n_samples % kWindow)) / kWindow;
}

void vad_speech_timestamps(const float* probs, long long n_windows, long long audio_len, const VadOpts& o,
                           std::vector<long long>* starts, std::vector<long long>* ends) {
    // Python arithmetic: the thresholds and sample counts below are floats (doubles) upstream, window * i is an int
    const double sr = 16000.0;
    const double threshold = o.threshold;
    const double neg_threshold = o.neg_threshold >= 0.0 ? o.neg_threshold : std::fmax(threshold - 0.15, 0.01);
    const double min_speech = sr * o.min_speech_duration_ms / 1000.0;
    const double pad = sr * o.speech_pad_ms / 1000.0;
    const double max_speech = (std::isfinite(o.max_speech_duration_s) && o.max_speech_duration_s > 0.0)
                                  ? sr * o.max_speech_duration_s - kWindow - 2.0 * pad
                                  : std::numeric_limits<double>::infinity();
    const double min_silence = sr * o.min_silence_duration_ms / 1000.0;
    const double min_silence_at_max = sr * 98.0 / 1000.0;

    bool triggered = false, have_cur = false;
    long long cur_start = 0, temp_end = 0, prev_end = 0, next_start = 0;
    starts->clear();
    ends->clear();
    for (long long i = 0; i < n_windows; ++i) {
        const double p = (double)probs[i];               // numpy 1.26.4 compares the float32 scalar in float64
        const long long pos = (long long)kWindow * i;
        if (p >= threshold && temp_end) {
            temp_end = 0;
            if (next_start < prev_end) next_start = pos;
        }
        if (p >= threshold && !triggered) {
            triggered = true;
            cur_start = pos;
            have_cur = true;
            continue;
        }
        if (triggered && (double)(pos - cur_start) > max_speech) {
            if (prev_end) {
                starts->push_back(cur_start);
                ends->push_back(prev_end);
                have_cur = false;
                if (next_start < prev_end) {
                    triggered = false;
                } else {
                    cur_start = next_start;
                    have_cur = true;
                }
                prev_end = next_start = temp_end = 0;
            } else {
                starts->push_back(cur_start);
                ends->push_back(pos);
                have_cur = false;
                prev_end = next_start = temp_end = 0;
                triggered = false;
                continue;
            }
        }
        if (p < neg_threshold && triggered) {
            if (!temp_end) temp_end = pos;
            if ((double)(pos - temp_end) > min_silence_at_max) prev_end = temp_end;
            if ((double)(pos - temp_end) < min_silence) continue;
            if ((double)(temp_end - cur_start) > min_speech) {
                starts->push_back(cur_start);
                ends->push_back(temp_end);
            }
            have_cur = false;
            prev_end = next_start = temp_end = 0;
            triggered = false;
            continue;
        }
    }
    if (have_cur && (double)(audio_len - cur_start) > min_speech) {
        starts->push_back(cur_start);
        ends->push_back(audio_len);
    }
    // padding pass: int(max(0, start - pad)), silence // 2 (floor division of an int), int(min(len, end + pad))
    const size_t n = starts->size();
    auto& S = *starts;
    auto& E = *ends;
    for (size_t i = 0; i < n; ++i) {
        if (i == 0) S[0] = (long long)std::fmax(0.0, (double)S[0] - pad);
        if (i + 1 != n) {
            const long long silence = S[i + 1] - E[i];
            if ((double)silence < 2.0 * pad) {
                const long long half = silence >= 0 ? silence / 2 : -((-silence + 1) / 2);    // Python floor division
                E[i] += half;
                S[i + 1] = (long long)std::fmax(0.0, (double)(S[i + 1] - half));
            } else {
                E[i] = (long long)std::fmin((double)audio_len, (double)E[i] + pad);
                S[i + 1] = (long long)std::fmax(0.0, (double)S[i + 1] - pad);
            }
        } else {
            E[i] = (long long)std::fmin((double)audio_len, (double)E[i] + pad);
        }
    }
}

cudaError_t vad_energy_probs(const float* pcm, long long n_samples, float center_db, float width_db, float* probs,
                             cudaStream_t stream) {
    const long long n_win = vad_num_windows(n_samples);
    if (n_win <= 0 || !(width_db > 0.0f)) return cudaErrorInvalidValue;
    const unsigned blocks = (unsigned)((n_win + 7) / 8);
    vad_energy_probs_kernel<<<blocks, 256, 0, stream>>>(pcm, n_samples, n_win, center_db, 1.0f / width_db, probs);
    return cudaGetLastError();
}

cudaError_t collect_chunks(const float* pcm, const long long* d_starts, const long long* d_offs, int n_chunks,
                           long long total, float* out, int sm_count, cudaStream_t stream) {
    if (total <= 0 || n_chunks <= 0) return cudaSuccess;
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)sm_count * 16) blocks = (long long)sm_count * 16;
    collect_chunks_kernel<<<(unsigned)blocks, 256, 0, stream>>>(pcm, d_starts, d_offs, n_chunks, total, out);
    return cudaGetLastError();
}

}  // namespace aries
