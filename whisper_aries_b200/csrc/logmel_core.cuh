// Per-thread math of the log-mel kernel (logmel.cu), written so that the SAME code also compiles as plain C++:
// tests/emu/logmel_emu.cpp walks it thread by thread on the CPU to check the index algebra without a GPU.
//
// What it computes (upstream faster-whisper 1.1.1 FeatureExtractor.stft/__call__,
// reached from the reference through model.transcribe, ref: final_optimized_transcriber.py:326):
//   frames of 400 samples every 160, periodic Hann (f32), 400-point DFT, |.|^2 for bins 0..200.
//
// How: a tile is 64 frames = 32 frame PAIRS; lane p of every warp owns the pair (p, p+32) and transforms
// z = frame_a + i*frame_b with ONE 400-point complex FFT, 400 = 20 x 20 (Cooley-Tukey), each 20-point
// transform done entirely in registers as a 4 x 5 Good-Thomas prime-factor FFT (no inner twiddles).
//   stage 1  (item n2 = 0..19):  A[k1]  = sum_n1 z[20 n1 + n2] w20^(n1 k1); the two real frames are separated
//                                 here (Aa = A[k1] + conj A[20-k1], Ab = (A[k1] - conj A[20-k1]) / i), so only
//                                 k1 = 0..10 go on; twiddled by W400^(n2 k1); k1 = 0 and k1 = 10 are real
//                                 sequences and stay packed (a + i b) -> exactly 20 outputs per thread.
//   stage 2  (item q  = 0..19):  20-point FFT over n2 -> 2 X[k1 + 20 k2]; k2 >= 10 lands on the mirrored bin
//                                 400 - k (conjugate symmetry), so every output of q >= 2 is a needed bin.
// Everything carries a factor 2 (power: 4) that is folded, exactly, into the mel weights (x 0.25).
// Lanes always index frames, so every shared-memory access is [element][lane]: conflict-free, and window /
// twiddle factors are warp-uniform.
#pragma once

#ifdef __CUDACC__
#define ARIES_HD __host__ __device__ __forceinline__
#else
#define ARIES_HD inline
#endif

namespace aries {
namespace mel {

constexpr int kNfft = 400;
constexpr int kHop = 160;
constexpr int kBins = 201;
constexpr int kTileFrames = 64;                      // 32 lanes x 2 frames
constexpr int kTileSamples = (kTileFrames - 1) * kHop + kNfft;        // 10480 padded-signal samples per tile
constexpr int kPcmWords = kTileSamples + kTileSamples / kHop + 1;     // skewed by one word per 160 samples
constexpr int kFrameStride = kHop + 1;               // 161: consecutive frames start one bank apart
constexpr int kItems = 20;                            // 20-point sub-transforms per stage and pair
constexpr int kExchangeFloat2 = kItems * 20 * 32;     // E[q][n2][lane]
constexpr int kPowerFloats = kBins * kTileFrames;     // P[bin][frame], aliases E

// Skewed shared-memory address of padded-signal sample s (relative to the tile start).
ARIES_HD int pcm_addr(int s) { return s + s / kHop; }

// Index into the un-padded signal for padded coordinate s (whole signal): zero-pad `padding` samples at the end,
// THEN reflect 200 at both ends (numpy "reflect", also when the pad is longer than the signal).
// Returns -1 for a sample that is zero (inside the zero pad).
ARIES_HD long long source_index(long long s, long long n_samples, long long padded_len /* n_samples + padding */) {
    long long j = s - kNfft / 2;
    if (j < 0 || j >= padded_len) {
        if (padded_len <= 1) {
            j = 0;
        } else {
            const long long period = 2 * (padded_len - 1);
            j %= period;
            if (j < 0) j += period;
            if (j >= padded_len) j = period - j;
        }
    }
    return j < n_samples ? j : -1;
}

struct Tables {
    float window[kNfft];       // periodic Hann, f32 (np.hanning(401)[:-1].astype(f32))
    float tw_re[11][20];       // cos(2 pi k1 n2 / 400)          (row 10 pre-multiplied by 2, see stage1)
    float tw_im[11][20];       // -sin(2 pi k1 n2 / 400)
};

// ------------------------------------------------------------------------------------------- small DFTs
// Complex numbers as (re, im) PAIRS: on the device every complex add / real-scaled FMA below is ONE packed f32x2
// instruction (FADD2 / FFMA2 / FMUL2 of sm_100) instead of two scalar ones -- 125 instead of 224 FP instructions per
// 20-point transform, which is 44 % of this issue-bound kernel's instruction stream.  Each component is computed by
// exactly the same IEEE operations in the same order as the scalar formulation, so the values are bit-identical; the
// host build (tests/emu) spells the pairs out.
struct cpx {
    float re, im;
};
#ifdef __CUDA_ARCH__
ARIES_HD cpx cadd(cpx a, cpx b) { const float2 r = __fadd2_rn(make_float2(a.re, a.im), make_float2(b.re, b.im)); return {r.x, r.y}; }
ARIES_HD cpx csub(cpx a, cpx b) { const float2 r = __fadd2_rn(make_float2(a.re, a.im), make_float2(-b.re, -b.im)); return {r.x, r.y}; }
ARIES_HD cpx cfma(float s, cpx a, cpx c) {                       // s * a + c, both components
    const float2 r = __ffma2_rn(make_float2(s, s), make_float2(a.re, a.im), make_float2(c.re, c.im));
    return {r.x, r.y};
}
ARIES_HD cpx cscale(float s, cpx a) { const float2 r = __fmul2_rn(make_float2(s, s), make_float2(a.re, a.im)); return {r.x, r.y}; }
#else
ARIES_HD cpx cadd(cpx a, cpx b) { return {a.re + b.re, a.im + b.im}; }
ARIES_HD cpx csub(cpx a, cpx b) { return {a.re - b.re, a.im - b.im}; }
ARIES_HD cpx cfma(float s, cpx a, cpx c) { return {fmaf(s, a.re, c.re), fmaf(s, a.im, c.im)}; }
ARIES_HD cpx cscale(float s, cpx a) { return {s * a.re, s * a.im}; }
#endif
ARIES_HD cpx cmul_neg_i(cpx a) { return {a.im, -a.re}; }        // -i a  (a swap and one sign: no arithmetic pipe needed for the swap)

// 5-point DFT, forward (e^{-2 pi i / 5}).
ARIES_HD void dft5(cpx x0, cpx x1, cpx x2, cpx x3, cpx x4, cpx* y) {
    const float c1 = 0.30901699437494742f;    // cos(2pi/5)
    const float c2 = -0.80901699437494742f;   // cos(4pi/5)
    const float s1 = 0.95105651629515357f;    // sin(2pi/5)
    const float s2 = 0.58778525229247313f;    // sin(4pi/5)
    const cpx t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
    y[0] = cadd(x0, cadd(t1, t2));
    const cpx m1 = cfma(c2, t2, cfma(c1, t1, x0));
    const cpx m2 = cfma(c1, t2, cfma(c2, t1, x0));
    const cpx u1 = cfma(s2, t4, cscale(s1, t3));
    const cpx u2 = cfma(-s1, t4, cscale(s2, t3));
    // X1 = m1 - i u1, X4 = m1 + i u1, X2 = m2 - i u2, X3 = m2 + i u2
    const cpx w1 = cmul_neg_i(u1), w2 = cmul_neg_i(u2);
    y[1] = cadd(m1, w1);
    y[4] = csub(m1, w1);
    y[2] = cadd(m2, w2);
    y[3] = csub(m2, w2);
}

// 20-point DFT, forward, in place, natural order in and out.  Good-Thomas: n = (5a + 4b) mod 20,
// k = (5 ka + 16 kb) mod 20  =>  w20^(nk) = w4^(a ka) w5^(b kb).
ARIES_HD void fft20(float (&xr)[20], float (&xi)[20]) {
    cpx t[4][5];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int n0 = (5 * a) % 20, n1 = (5 * a + 4) % 20, n2 = (5 * a + 8) % 20, n3 = (5 * a + 12) % 20,
                  n4 = (5 * a + 16) % 20;
        dft5({xr[n0], xi[n0]}, {xr[n1], xi[n1]}, {xr[n2], xi[n2]}, {xr[n3], xi[n3]}, {xr[n4], xi[n4]}, t[a]);
    }
#pragma unroll
    for (int kb = 0; kb < 5; ++kb) {
        const cpx s02 = cadd(t[0][kb], t[2][kb]), d02 = csub(t[0][kb], t[2][kb]);
        const cpx s13 = cadd(t[1][kb], t[3][kb]), d13 = csub(t[1][kb], t[3][kb]);
        const int k0 = (16 * kb) % 20, k1 = (5 + 16 * kb) % 20, k2 = (10 + 16 * kb) % 20, k3 = (15 + 16 * kb) % 20;
        const cpx w = cmul_neg_i(d13);                                  // -i d13
        const cpx y0 = cadd(s02, s13), y2 = csub(s02, s13), y1 = cadd(d02, w), y3 = csub(d02, w);
        xr[k0] = y0.re; xi[k0] = y0.im;
        xr[k2] = y2.re; xi[k2] = y2.im;
        xr[k1] = y1.re; xi[k1] = y1.im;      // (d02) - i (d13)
        xr[k3] = y3.re; xi[k3] = y3.im;      // (d02) + i (d13)
    }
}

// ------------------------------------------------------------------------------------------- stage 1
// pcm: skewed tile of the padded signal; E: exchange buffer as separate re/im planes of [20][20][32] floats.
ARIES_HD void stage1(const float* pcm, float* e_re, float* e_im, const Tables& tb, int n2, int lane) {
    float xr[20], xi[20];
    const float* pa = pcm + kFrameStride * lane;
    const float* pb = pcm + kFrameStride * (lane + 32);
#pragma unroll
    for (int n1 = 0; n1 < 20; ++n1) {
        const int i = 20 * n1 + n2;
        const int off = i + i / kHop;
        const float w = tb.window[i];
        xr[n1] = pa[off] * w;
        xi[n1] = pb[off] * w;
    }
    fft20(xr, xi);
    const int col = n2 * 32 + lane;
    // q = 0: k1 = 0, both frames' DC sequences packed as 2 (ra + i rb)
    e_re[0 * 640 + col] = xr[0] + xr[0];
    e_im[0 * 640 + col] = xi[0] + xi[0];
    // q = 1: k1 = 10, packed 2 A[10] W400^(10 n2) (the factor 2 sits in the table row 10)
    {
        const float c = tb.tw_re[10][n2], s = tb.tw_im[10][n2];
        e_re[1 * 640 + col] = xr[10] * c - xi[10] * s;
        e_im[1 * 640 + col] = xr[10] * s + xi[10] * c;
    }
#pragma unroll
    for (int k1 = 1; k1 <= 9; ++k1) {
        const float ar = xr[k1] + xr[20 - k1], ai = xi[k1] - xi[20 - k1];     // 2 * A_a[k1]
        const float br = xi[k1] + xi[20 - k1], bi = xr[20 - k1] - xr[k1];     // 2 * A_b[k1]
        const float c = tb.tw_re[k1][n2], s = tb.tw_im[k1][n2];
        const int qa = 2 * k1, qb = 2 * k1 + 1;
        e_re[qa * 640 + col] = ar * c - ai * s;
        e_im[qa * 640 + col] = ar * s + ai * c;
        e_re[qb * 640 + col] = br * c - bi * s;
        e_im[qb * 640 + col] = br * s + bi * c;
    }
}

// ------------------------------------------------------------------------------------------- stage 2
ARIES_HD void stage2_load(const float* e_re, const float* e_im, int q, int lane, float (&yr)[20], float (&yi)[20]) {
#pragma unroll
    for (int n2 = 0; n2 < 20; ++n2) {
        yr[n2] = e_re[q * 640 + n2 * 32 + lane];
        yi[n2] = e_im[q * 640 + n2 * 32 + lane];
    }
}

// Writes 4 |X[bin]|^2 into P[bin][frame] (frame = lane or lane + 32).
ARIES_HD void stage2_power(float* P, int q, int lane, float (&yr)[20], float (&yi)[20]) {
    fft20(yr, yi);
    if (q >= 2) {
        const int k1 = q >> 1;
        const int col = lane + 32 * (q & 1);
#pragma unroll
        for (int k2 = 0; k2 < 20; ++k2) {
            const int bin = (k2 < 10) ? (k1 + 20 * k2) : ((20 - k1) + 20 * (19 - k2));
            P[bin * kTileFrames + col] = yr[k2] * yr[k2] + yi[k2] * yi[k2];
        }
    } else if (q == 0) {
#pragma unroll
        for (int k2 = 0; k2 <= 10; ++k2) {
            const int pt = (20 - k2) % 20;
            const float ur = yr[k2] + yr[pt], ui = yi[k2] - yi[pt];
            const float vr = yi[k2] + yi[pt], vi = yr[k2] - yr[pt];
            P[(20 * k2) * kTileFrames + lane] = 0.25f * (ur * ur + ui * ui);
            P[(20 * k2) * kTileFrames + lane + 32] = 0.25f * (vr * vr + vi * vi);
        }
    } else {
#pragma unroll
        for (int k2 = 0; k2 <= 9; ++k2) {
            const int pt = 19 - k2;
            const float ur = yr[k2] + yr[pt], ui = yi[k2] - yi[pt];
            const float vr = yi[k2] + yi[pt], vi = yr[k2] - yr[pt];
            P[(10 + 20 * k2) * kTileFrames + lane] = 0.25f * (ur * ur + ui * ui);
            P[(10 + 20 * k2) * kTileFrames + lane + 32] = 0.25f * (vr * vr + vi * vi);
        }
    }
}

// ------------------------------------------------------------------------------------------- mel projection
// Sparse triangular filters: filter m covers bins [start[m], start[m] + count[m]) with weights w[offset[m] + t]
// (already multiplied by 0.25).  One call does both halves of the tile for one lane: frames `lane` and `lane + 32`.
constexpr int kMaxMelWeights = 1024;
constexpr int kMaxMels = 256;

struct MelBank {
    float w[kMaxMelWeights];
    short start[kMaxMels];
    short count[kMaxMels];
    short offset[kMaxMels];
};

ARIES_HD void mel_dot2(const float* P, const MelBank& mb, int m, int lane, float& acc0, float& acc1) {
    const int n = mb.count[m];
    const float* w = mb.w + mb.offset[m];
    const float* pp = P + mb.start[m] * kTileFrames + lane;
    float a0 = 0.0f, a1 = 0.0f;
#pragma unroll 4
    for (int t = 0; t < n; ++t) {
        const float wt = w[t];
        a0 = fmaf(wt, pp[t * kTileFrames], a0);
        a1 = fmaf(wt, pp[t * kTileFrames + 32], a1);
    }
    acc0 = a0;
    acc1 = a1;
}

// Fills a MelBank from a dense [n_mels, 201] filter matrix; returns false if it has more than 1024 taps.
inline bool build_mel_bank(const float* filters, int n_mels, MelBank& mb) {
    int used = 0;
    for (int m = 0; m < kMaxMels; ++m) mb.start[m] = mb.count[m] = mb.offset[m] = 0;
    for (int i = 0; i < kMaxMelWeights; ++i) mb.w[i] = 0.0f;
    for (int m = 0; m < n_mels; ++m) {
        int lo = -1, hi = -1;
        for (int k = 0; k < kBins; ++k)
            if (filters[m * kBins + k] != 0.0f) {
                if (lo < 0) lo = k;
                hi = k;
            }
        const int cnt = lo < 0 ? 0 : hi - lo + 1;
        if (used + cnt > kMaxMelWeights) return false;
        mb.start[m] = (short)(lo < 0 ? 0 : lo);
        mb.count[m] = (short)cnt;
        mb.offset[m] = (short)used;
        for (int k = 0; k < cnt; ++k) mb.w[used + k] = 0.25f * filters[m * kBins + lo + k];   // power carries x4
        used += cnt;
    }
    return true;
}

}  // namespace mel
}  // namespace aries
