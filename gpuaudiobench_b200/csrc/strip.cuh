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

#ifdef __CUDACC__
// The strip of ONE track run by ONE thread over `n` samples that already sit in shared memory (in place):
// what the fused UPOLS kernel does in its last-CTA epilogue, where the track's B output samples are in
// shared memory anyway, instead of a separate strip launch.  Same operations in the same order as the
// strip kernels (strip.cu), so the bits are the same; flags are tested per sample (uniform, no divergence).
__device__ __forceinline__ void strip_track_in_smem(const StripParams& p, int t, float* xs, int n) {
    const bool stats = p.ops & 1u, gain = p.ops & 2u, biquad = p.ops & 4u;  // B200CONV_STRIP_STATS / GAIN / BIQUAD
    const float g = gain ? (p.gains ? p.gains[t] : p.gain) : 1.0f;
    float b0 = 1.0f, b1 = 0.0f, b2 = 0.0f, a1 = 0.0f, a2 = 0.0f, z1 = 0.0f, z2 = 0.0f;
    if (biquad) {
        const float* c = p.coef + (p.shared_coef ? 0 : 5 * static_cast<size_t>(t));
        b0 = c[0]; b1 = c[1]; b2 = c[2]; a1 = c[3]; a2 = c[4];
        z1 = p.state[2 * t];
        z2 = p.state[2 * t + 1];
    }
    float mean = 0.0f, mx = -1e9f;
    for (int i = 0; i < n; i += 4) {  // n is a power of two >= 16 here
        float4 v = *reinterpret_cast<const float4*>(xs + i);
        float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float s = x[u];
            if (stats) {
                mean = __fadd_rn(mean, s);
                if (s > mx) mx = s;
            }
            if (gain) s = __fmul_rn(g, s);
            if (biquad) {
                const float w = __fsub_rn(__fsub_rn(s, __fmul_rn(a1, z1)), __fmul_rn(a2, z2));
                s = __fadd_rn(__fadd_rn(__fmul_rn(b0, w), __fmul_rn(b1, z1)), __fmul_rn(b2, z2));
                z2 = z1;
                z1 = w;
            }
            x[u] = s;
        }
        *reinterpret_cast<float4*>(xs + i) = make_float4(x[0], x[1], x[2], x[3]);
    }
    if (stats && p.stats) {
        p.stats[2 * t] = __fdiv_rn(mean, static_cast<float>(n));
        p.stats[2 * t + 1] = mx;
    }
    if (biquad && !p.peek) {
        p.state[2 * t] = z1;
        p.state[2 * t + 1] = z2;
    }
}
#endif

}  // namespace b200conv
