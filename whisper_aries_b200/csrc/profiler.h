// Optional per-kernel-class timing with CUDA events on the launching stream (bench.py's roofline numbers).
// Disabled by default: recording costs two event records per launch.
#pragma once
#include <cuda_runtime.h>

#include <vector>

namespace aries {

enum KernelClass {
    KC_MEL = 0,        // logmel_tiles_kernel
    KC_MEL_CLAMP,      // logmel_clamp_kernel
    KC_TRANSPOSE,      // mel_to_time_major_kernel
    KC_CONV1,          // gemm (implicit conv1)
    KC_CONV2,          // gemm (implicit conv2)
    KC_LAYERNORM,
    KC_QKV,
    KC_ATTENTION,
    KC_OPROJ,
    KC_FC1,
    KC_FC2,
    KC_COUNT
};

class Profiler {
public:
    ~Profiler() {
        for (auto& p : pool_) {
            cudaEventDestroy(p.a);
            cudaEventDestroy(p.b);
        }
    }
    void enable(bool on) { on_ = on; }
    bool enabled() const { return on_; }
    void begin(int cls, cudaStream_t s) {
        if (!on_) return;
        if (used_ == pool_.size()) {
            Pair p{};
            cudaEventCreate(&p.a);
            cudaEventCreate(&p.b);
            pool_.push_back(p);
        }
        pool_[used_].cls = cls;
        cudaEventRecord(pool_[used_].a, s);
    }
    void end(cudaStream_t s) {
        if (!on_) return;
        cudaEventRecord(pool_[used_].b, s);
        ++used_;
    }
    // Synchronises on the recorded events, ADDS per-class milliseconds / launch counts, forgets the records.
    cudaError_t collect(float* ms, int* counts) {
        for (size_t i = 0; i < used_; ++i) {
            cudaError_t e = cudaEventSynchronize(pool_[i].b);
            if (e != cudaSuccess) return e;
            float t = 0.f;
            if ((e = cudaEventElapsedTime(&t, pool_[i].a, pool_[i].b)) != cudaSuccess) return e;
            ms[pool_[i].cls] += t;
            counts[pool_[i].cls] += 1;
        }
        used_ = 0;
        return cudaSuccess;
    }

private:
    struct Pair {
        cudaEvent_t a, b;
        int cls;
    };
    std::vector<Pair> pool_;
    size_t used_ = 0;
    bool on_ = false;
};

}  // namespace aries
