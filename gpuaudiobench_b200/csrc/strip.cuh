// strip.cuh — per-track channel strip on the engine's output stage (strip.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace b200conv {

struct StripParams {
    const float* in;     // [T][B] track-major, or [B][ld] sample-major with track t in column col0 + t
    float* out;          // same layout; may alias `in`
    int T, B;
    int sample_major, ld, col0;
    uint32_t ops;        // B200CONV_STRIP_GAIN | _STATS | _BIQUAD
    float gain;          // used when gains == nullptr
    const float* gains;  // [T] or nullptr
    const float* coef;   // [T][5] b0 b1 b2 a1 a2, or [5] when shared_coef
    int shared_coef;
    float* state;        // [T][2] z1 z2 (biquad delay line), read and — unless peek — written back
    float* stats;        // [T][2] mean, max of the strip INPUT, or nullptr
    int peek;
};

cudaError_t launch_strip(const StripParams& p, cudaStream_t st);

}  // namespace b200conv
