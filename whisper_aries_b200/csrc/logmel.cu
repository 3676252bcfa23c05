// K1: log-mel spectrogram for sm_100a.  One persistent CTA per SM walks 64-frame tiles of the batch.
//
// Replaces faster-whisper 1.1.1 FeatureExtractor.__call__ (numpy, host, single thread),
// which the reference reaches through model.transcribe (ref: final_optimized_transcriber.py:326, warm-up :189).
//
// Per tile (see logmel_core.cuh for the transform):
//   PCM (prefetched into registers during the previous tile's mel phase, coalesced float4, reflect / zero pad
//   resolved on the fly) -> skewed shared tile -> stage 1 (window, 20-pt FFTs, frame separation, twiddles)
//   -> exchange -> stage 2 (20-pt FFTs, |X|^2) -> sparse mel filters -> log10 / scale in registers
//   -> coalesced stores of (log10(mel) + 4) / 4, plus the running per-item maximum (warp shuffles -> shared
//   atomics -> one global atomic per tile) and the tile minimum.
// The "max(x, global_max - 8)" clamp needs the maximum over the WHOLE call.  FeatureExtractor.__call__ (f32 output in
// upstream's [n_mels, frames] layout) applies it with a second, tiny kernel that only rewrites tiles whose minimum is
// below the threshold (most tiles of real audio are not).
// The fused PCM -> encoder path (aries_encode_pcm) uses the <kTM = true> instance instead: it writes the conv1 operand
// directly -- bf16, time-major [B, 3002, c_pad], a tile being 64 consecutive rows = one contiguous, fully coalesced
// block staged through shared memory -- so the f32 mel tensor never exists in HBM and the transpose kernel is gone.
// The clamp is then applied to the flagged tiles of that bf16 tensor (logmel_clamp_tm_kernel, still in L2):
// bf16(max(v, thr)) == max(bf16(v), bf16(thr)) because rounding is monotonic, so clamping the stored bf16 values is
// bit-identical to clamping in f32 first.  (Round 2 also tried folding that pass into the same launch -- the CTA that
// finishes a window's last tile fixing the window's flagged tiles -- and measured 2.0 ms instead of 0.2: on the
// benchmark's signals 47 % of the tiles are flagged and one CTA per window walks them serially.)
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

#include "logmel.h"
#include "logmel_core.cuh"
#include "profiler.h"

namespace aries {

using namespace mel;

namespace {

constexpr int kThreads = 640;                                   // 20 warps = the 20 items of each stage
constexpr int kPcmWordsPadded = (kPcmWords + 3) & ~3;
constexpr int kBankSlots = 4;                                   // distinct filter banks resident per device
constexpr int kMaxDevices = 64;
constexpr int kFloat4PerTile = kTileSamples / 4;                // 2620
constexpr int kPrefetch = (kFloat4PerTile + kThreads - 1) / kThreads;   // 5 float4 per thread

__constant__ Tables c_tables;
__constant__ MelBank c_bank[kBankSlots];       // warp-uniform reads: the tap loop runs on the uniform datapath

struct Smem {
    float pcm[2][kPcmWordsPadded];
    float e_re[kExchangeFloat2];      // P (201 x 64 floats) aliases e_re/e_im
    float e_im[kExchangeFloat2];
    unsigned red_max;
    unsigned red_min;
};
static_assert(kPowerFloats <= 2 * kExchangeFloat2, "power tile must fit in the exchange buffer");
static_assert(sizeof(Smem) <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ unsigned order_f32(float v) {
    const unsigned u = __float_as_uint(v);
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float unorder_f32(unsigned u) {
    return __uint_as_float(u ^ ((u >> 31) ? 0x80000000u : 0xFFFFFFFFu));
}

struct Params {
    const float* pcm;
    long long pcm_stride;
    long long n_samples;
    long long padded_len;       // n_samples + padding
    int batch;
    int n_frames;               // (n_samples + padding) / 160
    int tiles_per_item;
    int n_mels;
    float* out;
    int frames_out;
    long long out_stride;       // n_mels * frames_out
    unsigned* gmax;             // [batch], order_f32 encoded
    float* tile_min;            // [batch][tiles_per_item]
    int bank_slot;
    // <kTM>: bf16 time-major output [batch][frames_out + 2][c_pad] (rows 1 .. frames_out written)
    unsigned short* out_tm;
    int c_pad;
};

__device__ __forceinline__ unsigned short bf16_rn(float v) {
    unsigned u = __float_as_uint(v);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (unsigned short)(u >> 16);
}

__device__ __forceinline__ float fast_log2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void prefetch_tile(const Params& p, int tile, float4 (&r)[kPrefetch]) {
    const int b = tile / p.tiles_per_item;
    const int t = tile - b * p.tiles_per_item;
    const long long s0 = (long long)t * (kTileFrames * kHop);
    const float* base = p.pcm + (long long)b * p.pcm_stride;
    const bool fast = (s0 >= kNfft / 2) && (s0 + kTileSamples - kNfft / 2 <= p.n_samples) &&
                      ((reinterpret_cast<uintptr_t>(base) & 15) == 0);
    if (fast) {
        const float4* src = reinterpret_cast<const float4*>(base + (s0 - kNfft / 2));
#pragma unroll
        for (int k = 0; k < kPrefetch; ++k) {
            const int i4 = threadIdx.x + k * kThreads;
            if (i4 < kFloat4PerTile) r[k] = __ldg(src + i4);
        }
    } else {
#pragma unroll
        for (int k = 0; k < kPrefetch; ++k) {
            const int i4 = threadIdx.x + k * kThreads;
            if (i4 < kFloat4PerTile) {
                float v[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const long long j = source_index(s0 + 4 * i4 + c, p.n_samples, p.padded_len);
                    v[c] = (j >= 0) ? __ldg(base + j) : 0.0f;
                }
                r[k] = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
    }
}

__device__ __forceinline__ void store_tile(float* pcm, const float4 (&r)[kPrefetch]) {
#pragma unroll
    for (int k = 0; k < kPrefetch; ++k) {
        const int i4 = threadIdx.x + k * kThreads;
        if (i4 < kFloat4PerTile) {
            float* d = pcm + pcm_addr(4 * i4);         // the 4 samples share one 160-block: consecutive words
            // lanes L and L + 8 start 32 words apart (the same bank): each group of 8 lanes writes its four words in
            // a different rotation, so one store instruction touches 32 distinct banks (r01: 2.8 M conflicts here)
            const float v[4] = {r[k].x, r[k].y, r[k].z, r[k].w};
            const int rot = (threadIdx.x >> 3) & 3;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int e = (c + rot) & 3;
                d[e] = (e == 0) ? v[0] : (e == 1) ? v[1] : (e == 2) ? v[2] : v[3];
            }
        }
    }
}

constexpr int kStagePitch = 130;                 // bf16 per staged row: 65 words, so lane = frame hits 32 banks

template <bool kTM>
__global__ void __launch_bounds__(kThreads, 1) logmel_tiles_kernel(const Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    float* P = sm.e_re;
    // <kTM> staging tile [64 frames][kStagePitch] bf16 behind the power tile (P ends 64 floats into e_im)
    unsigned short* staged = reinterpret_cast<unsigned short*>(sm.e_im + 64);
    static_assert(64 * 4 + kTileFrames * kStagePitch * 2 <= kExchangeFloat2 * 4, "staging tile must fit behind P");

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int total_tiles = p.batch * p.tiles_per_item;

    if (threadIdx.x == 0) {
        sm.red_max = 0u;
        sm.red_min = 0xFFFFFFFFu;
    }
    const MelBank& bank = c_bank[p.bank_slot];

    int tile = blockIdx.x;
    if (tile < total_tiles) {
        float4 r[kPrefetch];
        prefetch_tile(p, tile, r);
        store_tile(sm.pcm[0], r);
    }
    __syncthreads();

    for (int it = 0; tile < total_tiles; tile += gridDim.x, ++it) {
        const float* cur = sm.pcm[it & 1];
        float* nxt = sm.pcm[(it & 1) ^ 1];
        const int b = tile / p.tiles_per_item;
        const int t = tile - b * p.tiles_per_item;

        stage1(cur, sm.e_re, sm.e_im, c_tables, warp, lane);
        __syncthreads();
        {
            float yr[20], yi[20];
            stage2_load(sm.e_re, sm.e_im, warp, lane, yr, yi);
            __syncthreads();                      // every E read is done before P (aliasing E) is written
            stage2_power(P, warp, lane, yr, yi);
        }
        __syncthreads();

        const int next = tile + gridDim.x;
        const bool has_next = next < total_tiles;
        float4 r[kPrefetch];
        if (has_next) prefetch_tile(p, next, r);

        // ---- mel filters + log: a warp owns mel bins warp, warp + 20, ...; lane = frame (and frame + 32)
        float vmax = -INFINITY, vmin = INFINITY;
        {
            const int f0 = t * kTileFrames + lane;
            float* o0 = p.out + (long long)b * p.out_stride + f0;
            const int lim = p.n_frames < p.frames_out ? p.n_frames : p.frames_out;
            const bool full_tile = (t + 1) * kTileFrames <= lim;          // uniform: every frame is real and stored
            for (int m = warp; m < (kTM ? p.c_pad : p.n_mels); m += kThreads / 32) {
                if (kTM && m >= p.n_mels) {                    // channel padding of the conv1 operand
                    staged[lane * kStagePitch + m] = 0;
                    staged[(lane + 32) * kStagePitch + m] = 0;
                    continue;
                }
                float a0, a1;
                mel_dot2(P, bank, m, lane, a0, a1);
                const float v0 = fmaf(fast_log2(fmaxf(a0, 1e-10f)), 0.07525749891599529f, 1.0f);   // (log10 + 4) / 4
                const float v1 = fmaf(fast_log2(fmaxf(a1, 1e-10f)), 0.07525749891599529f, 1.0f);
                if (kTM) {
                    // frames past the audio are zero FEATURES (pad_or_trim); frames past frames_out are not stored below
                    const bool r0 = f0 < p.n_frames, r1 = f0 + 32 < p.n_frames;
                    if (r0) { vmax = fmaxf(vmax, v0); vmin = fminf(vmin, v0); }
                    if (r1) { vmax = fmaxf(vmax, v1); vmin = fminf(vmin, v1); }
                    staged[lane * kStagePitch + m] = r0 ? bf16_rn(v0) : (unsigned short)0;
                    staged[(lane + 32) * kStagePitch + m] = r1 ? bf16_rn(v1) : (unsigned short)0;
                    continue;
                }
                float* om = o0 + (size_t)m * p.frames_out;
                if (full_tile) {
                    vmax = fmaxf(vmax, fmaxf(v0, v1));
                    vmin = fminf(vmin, fminf(v0, v1));
                    om[0] = v0;
                    om[32] = v1;
                } else {
                    if (f0 < p.n_frames) {
                        vmax = fmaxf(vmax, v0);
                        vmin = fminf(vmin, v0);
                        if (f0 < p.frames_out) om[0] = v0;
                    }
                    if (f0 + 32 < p.n_frames) {
                        vmax = fmaxf(vmax, v1);
                        vmin = fminf(vmin, v1);
                        if (f0 + 32 < p.frames_out) om[32] = v1;
                    }
                }
            }
        }
        if (has_next) store_tile(nxt, r);

#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
            vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        }
        if (lane == 0) {
            atomicMax(&sm.red_max, order_f32(vmax));
            atomicMin(&sm.red_min, order_f32(vmin));
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            atomicMax(p.gmax + b, sm.red_max);
            p.tile_min[tile] = unorder_f32(sm.red_min);
            sm.red_max = 0u;
            sm.red_min = 0xFFFFFFFFu;
        }
        // the next write of red_* is three barriers away; the next read of P/E likewise
        if (kTM) {
            // the tile = 64 consecutive rows of the time-major tensor: one contiguous block, 4-byte coalesced stores
            const int wpr = p.c_pad >> 1;                                       // 32-bit words per row
            unsigned* dst = reinterpret_cast<unsigned*>(p.out_tm) +
                            ((long long)b * (p.frames_out + 2) + 1 + (long long)t * kTileFrames) * wpr;
            const unsigned* src = reinterpret_cast<const unsigned*>(staged);
            const int rows = min(kTileFrames, p.frames_out - t * kTileFrames);
            for (int i = threadIdx.x; i < rows * wpr; i += kThreads) {
                const int row = i / wpr, w = i - row * wpr;
                dst[i] = src[row * (kStagePitch / 2) + w];
            }
            __syncthreads();                  // staged (aliasing E) is free again before the next tile's stage 1
        }
    }
}

// Second pass: x = max(x, max - 2) where needed (== (max(log10, gmax - 8) + 4) / 4), zeros after the last frame
// when the caller asked for more frames than the audio has (pad_or_trim).
__global__ void __launch_bounds__(256) logmel_clamp_kernel(const Params p) {
    const int t = blockIdx.x;
    const int b = blockIdx.y;
    const float thr = unorder_f32(p.gmax[b]) - 2.0f;
    const int f0 = t * kTileFrames;
    const bool has_tail = f0 + kTileFrames > p.n_frames;               // tile reaches past the last real frame
    const bool needs_clamp = (t < p.tiles_per_item) && (p.tile_min[b * p.tiles_per_item + t] < thr);
    if (!has_tail && !needs_clamp) return;
    float* out_b = p.out + (long long)b * p.out_stride;
    for (int i = threadIdx.x; i < p.n_mels * kTileFrames; i += blockDim.x) {
        const int m = i / kTileFrames;
        const int f = f0 + (i % kTileFrames);
        if (f >= p.frames_out) continue;
        float* q = out_b + (long long)m * p.frames_out + f;
        if (f >= p.n_frames) {
            *q = 0.0f;
        } else if (needs_clamp) {
            const float v = *q;
            if (v < thr) *q = thr;
        }
    }
}

// Clamp pass of the fused path: one CTA per (tile, window); only tiles whose minimum is below the window's threshold do
// anything.  16 words (32 bf16) per thread, all loads issued before the first store.
__global__ void __launch_bounds__(256) logmel_clamp_tm_kernel(const Params p) {
    const int t = blockIdx.x;
    const int b = blockIdx.y;
    const float thr = unorder_f32(p.gmax[b]) - 2.0f;
    if (!(p.tile_min[b * p.tiles_per_item + t] < thr)) return;
    const unsigned short thr_bf = bf16_rn(thr);
    const float thr_r = __uint_as_float((unsigned)thr_bf << 16);
    const int wpr = p.c_pad >> 1;
    const int lim = min(p.n_frames, p.frames_out);
    const int rows = min(kTileFrames, lim - t * kTileFrames);              // real, stored frames of this tile
    if (rows <= 0) return;
    unsigned* tile_w = reinterpret_cast<unsigned*>(p.out_tm) +
                       ((long long)b * (p.frames_out + 2) + 1 + (long long)t * kTileFrames) * wpr;
    const int mel_words = (p.n_mels + 1) >> 1;
    const int total = rows * mel_words;
    constexpr int kPer = 16;
    unsigned u[kPer];
    int idx[kPer];
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        const int i = threadIdx.x + k * 256;
        const int row = i / mel_words, w = i - row * mel_words;
        idx[k] = (i < total) ? row * wpr + w : -1;
        u[k] = (i < total) ? tile_w[idx[k]] : 0u;
    }
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        if (idx[k] < 0) continue;
        const int w = idx[k] % wpr;
        const float lo = __uint_as_float(u[k] << 16), hi = __uint_as_float(u[k] & 0xFFFF0000u);
        const bool last_odd = (2 * w + 1 >= p.n_mels);                      // channel padding stays 0
        const unsigned nlo = lo < thr_r ? thr_bf : (u[k] & 0xFFFFu);
        const unsigned nhi = (!last_odd && hi < thr_r) ? thr_bf : (u[k] >> 16);
        const unsigned nu = nlo | (nhi << 16);
        if (nu != u[k]) tile_w[idx[k]] = nu;
    }
}

// [B, n_mels, frames] f32  ->  [B, frames + 2, c_pad] bf16, time-major with one zero row before and after and
// zero channels beyond n_mels: the layout the conv1 implicit GEMM reads (encoder.cu).
__global__ void __launch_bounds__(256) mel_to_time_major_kernel(const float* __restrict__ mel, int n_mels, int frames,
                                                                int frames_in_stride, unsigned short* __restrict__ out,
                                                                int c_pad) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int f0 = blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
    const float* src = mel + (long long)b * n_mels * frames_in_stride;
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r, f = f0 + tx;
        tile[r][tx] = (c < n_mels && f < frames) ? src[(long long)c * frames_in_stride + f] : 0.0f;
    }
    __syncthreads();
    unsigned short* dst = out + ((long long)b * (3000 + 2) + 1) * c_pad;
    for (int r = ty; r < 32; r += 8) {
        const int f = f0 + r, c = c0 + tx;
        if (f < 3000 && c < c_pad) {
            const float v = tile[tx][r];
            // round-to-nearest-even bf16
            unsigned u = __float_as_uint(v);
            u += 0x7FFFu + ((u >> 16) & 1u);
            dst[(long long)f * c_pad + c] = (unsigned short)(u >> 16);
        }
    }
}

void build_tables(Tables& tb) {
    const double pi = 3.14159265358979323846;
    for (int i = 0; i < kNfft; ++i) tb.window[i] = (float)(0.5 - 0.5 * std::cos(2.0 * pi * i / kNfft));
    for (int k1 = 0; k1 <= 10; ++k1) {
        for (int n2 = 0; n2 < 20; ++n2) {
            const double a = 2.0 * pi * (double)(k1 * n2) / 400.0;
            const double scale = (k1 == 10) ? 2.0 : 1.0;
            tb.tw_re[k1][n2] = (float)(scale * std::cos(a));
            tb.tw_im[k1][n2] = (float)(-scale * std::sin(a));
        }
    }
}

}  // namespace

void logmel_host_tables(mel::Tables* tb) { build_tables(*tb); }

struct LogmelPlan {
    int device = 0;
    int n_mels = 0;
    int sm_count = 0;
    int bank_slot = -1;
    // scratch, grown on demand
    unsigned* d_gmax = nullptr;
    float* d_tile_min = nullptr;
    int cap_batch = 0;
    long long cap_tiles = 0;
};

namespace {
// Filter banks live in __constant__ memory (warp-uniform reads).  A device holds up to kBankSlots distinct banks;
// handles with identical filters share a slot.
struct BankSlot {
    int refs = 0;
    MelBank bank;
};
std::mutex g_bank_mutex;
BankSlot* g_slots[kMaxDevices] = {};

int acquire_bank_slot(int device, const MelBank& mb, cudaError_t* err) {
    std::lock_guard<std::mutex> lock(g_bank_mutex);
    *err = cudaSuccess;
    if (device < 0 || device >= kMaxDevices) return -1;
    if (!g_slots[device]) g_slots[device] = new BankSlot[kBankSlots];
    BankSlot* slots = g_slots[device];
    for (int i = 0; i < kBankSlots; ++i)
        if (slots[i].refs > 0 && std::memcmp(&slots[i].bank, &mb, sizeof(MelBank)) == 0) {
            ++slots[i].refs;
            return i;
        }
    for (int i = 0; i < kBankSlots; ++i)
        if (slots[i].refs == 0) {
            *err = cudaMemcpyToSymbol(c_bank, &mb, sizeof(MelBank), (size_t)i * sizeof(MelBank));
            if (*err != cudaSuccess) return -1;
            slots[i].bank = mb;
            slots[i].refs = 1;
            return i;
        }
    return -1;
}

void release_bank_slot(int device, int slot) {
    std::lock_guard<std::mutex> lock(g_bank_mutex);
    if (device >= 0 && device < kMaxDevices && g_slots[device] && slot >= 0 && slot < kBankSlots &&
        g_slots[device][slot].refs > 0)
        --g_slots[device][slot].refs;
}
}  // namespace

cudaError_t logmel_plan_create(int device, int sm_count, int n_mels, const float* filters, LogmelPlan** out,
                               const char** why) {
    *out = nullptr;
    if (n_mels <= 0 || n_mels > kMaxMels) {
        *why = "n_mels out of range (1..256)";
        return cudaErrorInvalidValue;
    }
    static thread_local MelBank mb;
    if (!build_mel_bank(filters, n_mels, mb)) {
        *why = "mel filter bank too dense for the sparse kernel (more than 1024 taps)";
        return cudaErrorInvalidValue;
    }
    LogmelPlan* pl = new LogmelPlan();
    pl->device = device;
    pl->n_mels = n_mels;
    pl->sm_count = sm_count;
    cudaError_t e;
    Tables tb;
    build_tables(tb);
    if ((e = cudaMemcpyToSymbol(c_tables, &tb, sizeof(tb))) != cudaSuccess) goto fail;
    if ((e = cudaFuncSetAttribute(logmel_tiles_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)sizeof(Smem))) != cudaSuccess)
        goto fail;
    if ((e = cudaFuncSetAttribute(logmel_tiles_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)sizeof(Smem))) != cudaSuccess)
        goto fail;
    pl->bank_slot = acquire_bank_slot(device, mb, &e);
    if (e != cudaSuccess) goto fail;
    if (pl->bank_slot < 0) {
        *why = "too many distinct mel filter banks on this device (at most 4 at a time)";
        delete pl;
        return cudaErrorInvalidValue;
    }
    *out = pl;
    return cudaSuccess;
fail:
    *why = cudaGetErrorString(e);
    logmel_plan_destroy(pl);
    return e;
}

void logmel_plan_destroy(LogmelPlan* pl) {
    if (!pl) return;
    release_bank_slot(pl->device, pl->bank_slot);
    cudaFree(pl->d_gmax);
    cudaFree(pl->d_tile_min);
    delete pl;
}

int logmel_plan_n_mels(const LogmelPlan* pl) { return pl->n_mels; }

cudaError_t logmel_run(LogmelPlan* pl, const float* pcm, int batch, long long n_samples, long long pcm_stride,
                       int padding, float* out, int frames_out, cudaStream_t stream, int* launches, Profiler* prof) {
    const long long padded = n_samples + padding;
    const int n_frames = (int)(padded / kHop);
    const int tiles = (n_frames + kTileFrames - 1) / kTileFrames;
    cudaError_t e;
    if (batch > pl->cap_batch || (long long)batch * tiles > pl->cap_tiles) {
        // growing the scratch synchronises; steady-state calls with the same shape do not
        if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
        cudaFree(pl->d_gmax);
        cudaFree(pl->d_tile_min);
        pl->d_gmax = nullptr;
        pl->d_tile_min = nullptr;
        pl->cap_batch = 0;
        pl->cap_tiles = 0;
        if ((e = cudaMalloc(&pl->d_gmax, sizeof(unsigned) * batch)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&pl->d_tile_min, sizeof(float) * (size_t)batch * (tiles > 0 ? tiles : 1))) != cudaSuccess)
            return e;
        pl->cap_batch = batch;
        pl->cap_tiles = (long long)batch * tiles;
    }
    Params p{};
    p.pcm = pcm;
    p.pcm_stride = pcm_stride;
    p.n_samples = n_samples;
    p.padded_len = padded;
    p.batch = batch;
    p.n_frames = n_frames;
    p.tiles_per_item = tiles;
    p.n_mels = pl->n_mels;
    p.out = out;
    p.frames_out = frames_out;
    p.out_stride = (long long)pl->n_mels * frames_out;
    p.gmax = pl->d_gmax;
    p.tile_min = pl->d_tile_min;
    p.bank_slot = pl->bank_slot;
    int n_launch = 0;
    if ((e = cudaMemsetAsync(pl->d_gmax, 0, sizeof(unsigned) * batch, stream)) != cudaSuccess) return e;
    const long long total = (long long)batch * tiles;
    if (total > 0) {
        const int grid = (int)(total < pl->sm_count ? total : pl->sm_count);
        if (prof) prof->begin(KC_MEL, stream);
        logmel_tiles_kernel<false><<<grid, kThreads, sizeof(Smem), stream>>>(p);
        if (prof) prof->end(stream);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        ++n_launch;
    }
    const int out_tiles = (frames_out + kTileFrames - 1) / kTileFrames;
    const int clamp_tiles = out_tiles > tiles ? out_tiles : tiles;
    if (clamp_tiles > 0 && frames_out > 0) {
        if (prof) prof->begin(KC_MEL_CLAMP, stream);
        logmel_clamp_kernel<<<dim3(clamp_tiles, batch), 256, 0, stream>>>(p);
        if (prof) prof->end(stream);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        ++n_launch;
    }
    if (launches) *launches = n_launch;
    return cudaSuccess;
}

cudaError_t logmel_run_time_major(LogmelPlan* pl, const float* pcm, int batch, long long n_samples, long long pcm_stride,
                                  int padding, void* out_tm_bf16, int frames_out, int c_pad, cudaStream_t stream,
                                  int* launches, Profiler* prof) {
    const long long padded = n_samples + padding;
    const int n_frames = (int)(padded / kHop);
    const int tiles = (n_frames + kTileFrames - 1) / kTileFrames;
    if (c_pad < pl->n_mels || c_pad > kStagePitch - 2 || (c_pad & 1) || frames_out <= 0 || tiles <= 0)
        return cudaErrorInvalidValue;
    cudaError_t e;
    if (batch > pl->cap_batch || (long long)batch * tiles > pl->cap_tiles) {
        if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
        cudaFree(pl->d_gmax);
        cudaFree(pl->d_tile_min);
        pl->d_gmax = nullptr;
        pl->d_tile_min = nullptr;
        pl->cap_batch = 0;
        pl->cap_tiles = 0;
        if ((e = cudaMalloc(&pl->d_gmax, sizeof(unsigned) * batch)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&pl->d_tile_min, sizeof(float) * (size_t)batch * tiles)) != cudaSuccess) return e;
        pl->cap_batch = batch;
        pl->cap_tiles = (long long)batch * tiles;
    }
    Params p{};
    p.pcm = pcm;
    p.pcm_stride = pcm_stride;
    p.n_samples = n_samples;
    p.padded_len = padded;
    p.batch = batch;
    p.n_frames = n_frames;
    p.tiles_per_item = tiles;
    p.n_mels = pl->n_mels;
    p.frames_out = frames_out;
    p.gmax = pl->d_gmax;
    p.tile_min = pl->d_tile_min;
    p.bank_slot = pl->bank_slot;
    p.out_tm = static_cast<unsigned short*>(out_tm_bf16);
    p.c_pad = c_pad;
    if ((e = cudaMemsetAsync(pl->d_gmax, 0, sizeof(unsigned) * batch, stream)) != cudaSuccess) return e;
    if (tiles * kTileFrames < frames_out) {
        // audio shorter than the window: rows past its last tile are zero features (pad_or_trim); a rare, ragged-tail path
        const size_t row_bytes = (size_t)c_pad * 2, first = (size_t)1 + (size_t)tiles * kTileFrames;
        if ((e = cudaMemset2DAsync(static_cast<char*>(out_tm_bf16) + first * row_bytes, (size_t)(frames_out + 2) * row_bytes,
                                   0, (size_t)(frames_out - tiles * kTileFrames) * row_bytes, batch, stream)) != cudaSuccess)
            return e;
    }
    const long long total = (long long)batch * tiles;
    const int grid = (int)(total < pl->sm_count ? total : pl->sm_count);
    if (prof) prof->begin(KC_MEL, stream);
    logmel_tiles_kernel<true><<<grid, kThreads, sizeof(Smem), stream>>>(p);
    if (prof) prof->end(stream);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    const int real_tiles = ((n_frames < frames_out ? n_frames : frames_out) + kTileFrames - 1) / kTileFrames;
    static_assert(kTileFrames * ((kStagePitch - 2) / 2) <= 16 * 256, "clamp kernel covers a tile with 16 words per thread");
    if (prof) prof->begin(KC_MEL_CLAMP, stream);
    logmel_clamp_tm_kernel<<<dim3(real_tiles, batch), 256, 0, stream>>>(p);
    if (prof) prof->end(stream);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (launches) *launches = 2;
    return cudaSuccess;
}

cudaError_t mel_to_time_major(const float* mel, int batch, int n_mels, int frames, void* out_bf16, int c_pad,
                              cudaStream_t stream) {
    dim3 grid((3000 + 31) / 32, (c_pad + 31) / 32, batch);
    mel_to_time_major_kernel<<<grid, 256, 0, stream>>>(mel, n_mels, frames, frames,
                                                       reinterpret_cast<unsigned short*>(out_bf16), c_pad);
    return cudaGetLastError();
}

}  // namespace aries
