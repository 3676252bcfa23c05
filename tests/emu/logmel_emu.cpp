// CPU walk-through of the log-mel kernel's per-thread code (whisper_aries_b200/csrc/logmel_core.cuh): the same
// functions the CUDA kernel calls, executed thread by thread with plain arrays standing in for shared memory.
// TEST INFRASTRUCTURE: checks the FFT factorisation / indexing against the numpy oracle without a GPU.  It is not a
// fallback and is never loaded by the product package.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../whisper_aries_b200/csrc/logmel_core.cuh"

using namespace aries::mel;

static void build_tables(Tables& tb) {
    const double pi = 3.14159265358979323846;
    for (int i = 0; i < kNfft; ++i) tb.window[i] = (float)(0.5 - 0.5 * std::cos(2.0 * pi * i / kNfft));
    for (int k1 = 0; k1 <= 10; ++k1)
        for (int n2 = 0; n2 < 20; ++n2) {
            const double a = 2.0 * pi * (double)(k1 * n2) / 400.0, sc = (k1 == 10) ? 2.0 : 1.0;
            tb.tw_re[k1][n2] = (float)(sc * std::cos(a));
            tb.tw_im[k1][n2] = (float)(-sc * std::sin(a));
        }
}

extern "C" int emu_fft20(const float* in_re, const float* in_im, float* out_re, float* out_im) {
    float xr[20], xi[20];
    for (int i = 0; i < 20; ++i) { xr[i] = in_re[i]; xi[i] = in_im[i]; }
    fft20(xr, xi);
    for (int i = 0; i < 20; ++i) { out_re[i] = xr[i]; out_im[i] = xi[i]; }
    return 0;
}

// pcm [n_samples] -> out [n_mels, frames_out] with the kernel's two-pass clamp.  filters: [n_mels, 201].
extern "C" int emu_logmel(const float* pcm, long long n_samples, int padding, const float* filters, int n_mels,
                          float* out, int frames_out) {
    Tables tb;
    build_tables(tb);
    static MelBank bank;
    if (n_mels > kMaxMels || !build_mel_bank(filters, n_mels, bank)) return -1;
    const long long padded = n_samples + padding;
    const int n_frames = (int)(padded / kHop);
    const int tiles = (n_frames + kTileFrames - 1) / kTileFrames;
    std::vector<float> pcm_s(kPcmWords), e_re(kExchangeFloat2), e_im(kExchangeFloat2), P(kPowerFloats);
    std::vector<float> tile_min(tiles > 0 ? tiles : 1);
    float gmax = -INFINITY;
    for (int t = 0; t < tiles; ++t) {
        const long long s0 = (long long)t * kTileFrames * kHop;
        for (int s = 0; s < kTileSamples; ++s) {
            const long long j = source_index(s0 + s, n_samples, padded);
            pcm_s[pcm_addr(s)] = j >= 0 ? pcm[j] : 0.0f;
        }
        for (int warp = 0; warp < 20; ++warp)
            for (int lane = 0; lane < 32; ++lane) stage1(pcm_s.data(), e_re.data(), e_im.data(), tb, warp, lane);
        for (int warp = 0; warp < 20; ++warp)
            for (int lane = 0; lane < 32; ++lane) {
                float yr[20], yi[20];
                stage2_load(e_re.data(), e_im.data(), warp, lane, yr, yi);
                stage2_power(P.data(), warp, lane, yr, yi);
            }
        float tmin = INFINITY;
        for (int m = 0; m < n_mels; ++m)
            for (int lane = 0; lane < 32; ++lane) {
                float acc[2];
                mel_dot2(P.data(), bank, m, lane, acc[0], acc[1]);
                for (int h = 0; h < 2; ++h) {
                    const int f = t * kTileFrames + lane + 32 * h;
                    if (f >= n_frames) continue;
                    const float v = std::fmaf(std::log2(std::fmax(acc[h], 1e-10f)), 0.07525749891599529f, 1.0f);
                    gmax = std::fmax(gmax, v);
                    tmin = std::fmin(tmin, v);
                    if (f < frames_out) out[(long long)m * frames_out + f] = v;
                }
            }
        tile_min[t] = tmin;
    }
    const float thr = gmax - 2.0f;
    for (int m = 0; m < n_mels; ++m)
        for (int f = 0; f < frames_out; ++f) {
            float& v = out[(long long)m * frames_out + f];
            if (f >= n_frames) v = 0.0f;
            else if (v < thr) v = thr;
        }
    return n_frames;
}
