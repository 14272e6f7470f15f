// strip.cu — per-track channel strip: input statistics, gain and one Direct-Form-II biquad, applied to
// the engine's output before the stereo bus.  It replaces three reference kernels
// that all run ONE THREAD PER TRACK over the whole buffer with stride-B global accesses:
//     GainKernel        cuda/bench_gain.cu:6-24        y = gain * x
//     GainStatsKernel   cuda/bench_gainstats.cu:7-32   y = x * gain; mean, max of x per track
//     IIRFilterKernel   cuda/bench_iir.cu:10-44        w = x - a1 z1 - a2 z2; y = b0 w + b1 z1 + b2 z2
// Order inside the strip: statistics of the input, then gain, then the biquad.
//
// The recursion (and the reference's running sum for the mean) is sequential per track, so the work per
// track is a dependent chain of B steps; what the kernel can fix is memory access and the chain length:
//   track-major  [T][B]:  one WARP per track.  Lanes load 32 consecutive samples (one 128-B line), the
//       warp walks them with a shuffle broadcast and every lane runs the identical chain (no divergence,
//       the shuffles are off the critical path); lane i keeps sample i's result, stores are coalesced.
//   sample-major [B][Tg]: one THREAD per track, lanes = adjacent tracks, so a row load is coalesced;
//       rows are fetched 32 ahead (double buffered in registers) to cover the L2 latency.
// Arithmetic uses __fmul_rn/__fadd_rn/__fsub_rn in the reference's evaluation order, so outputs, filter
// state and statistics are BIT-IDENTICAL to the CPU loops (bench_gain.cu:90-92, bench_gainstats.cu:
// 121-142, bench_iir.cu:176-203; the host build does not contract a*b+c).
#include "strip.cuh"

#include "../../include/b200conv.h"
#include "common.cuh"

namespace b200conv {

namespace {

constexpr int kStripWarps = 4;    // warps (= tracks) per CTA, track-major kernel
constexpr int kStripThreads = 64; // threads (= tracks) per CTA, sample-major kernel
constexpr int kStripAhead = 32;   // rows fetched ahead, sample-major kernel

struct Biquad {
    float b0, b1, b2, a1, a2;
};

template <bool STATS, bool GAIN, bool BIQUAD>
struct Chain {
    float g, z1, z2, mean, mx;
    Biquad c;
    __device__ __forceinline__ float step(float x) {
        if (STATS) {
            mean = __fadd_rn(mean, x);
            if (x > mx) mx = x;
        }
        if (GAIN) x = __fmul_rn(g, x);
        if (BIQUAD) {
            const float w = __fsub_rn(__fsub_rn(x, __fmul_rn(c.a1, z1)), __fmul_rn(c.a2, z2));
            x = __fadd_rn(__fadd_rn(__fmul_rn(c.b0, w), __fmul_rn(c.b1, z1)), __fmul_rn(c.b2, z2));
            z2 = z1;
            z1 = w;
        }
        return x;
    }
};

template <bool STATS, bool GAIN, bool BIQUAD>
__device__ __forceinline__ Chain<STATS, GAIN, BIQUAD> chain_begin(const StripParams& p, int t) {
    Chain<STATS, GAIN, BIQUAD> ch;
    ch.g = GAIN ? (p.gains ? p.gains[t] : p.gain) : 1.0f;
    ch.mean = 0.0f;
    ch.mx = -1e9f;  // the reference's start value (bench_gainstats.cu:16)
    ch.z1 = ch.z2 = 0.0f;
    ch.c = Biquad{1.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    if (BIQUAD) {
        const float* c = p.coef + (p.shared_coef ? 0 : 5 * static_cast<size_t>(t));
        ch.c = Biquad{c[0], c[1], c[2], c[3], c[4]};
        ch.z1 = p.state[2 * t];
        ch.z2 = p.state[2 * t + 1];
    }
    return ch;
}

template <bool STATS, bool GAIN, bool BIQUAD>
__device__ __forceinline__ void chain_end(const StripParams& p, int t, const Chain<STATS, GAIN, BIQUAD>& ch) {
    if (STATS && p.stats) {
        p.stats[2 * t] = __fdiv_rn(ch.mean, static_cast<float>(p.B));
        p.stats[2 * t + 1] = ch.mx;
    }
    if (BIQUAD && !p.peek) {
        p.state[2 * t] = ch.z1;
        p.state[2 * t + 1] = ch.z2;
    }
}

template <bool STATS, bool GAIN, bool BIQUAD>
__global__ void __launch_bounds__(kStripWarps * 32) strip_rows_kernel(StripParams p) {
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * kStripWarps + (threadIdx.x >> 5);
    pdl_launch_dependents();
    pdl_wait_primary();  // the input is the output of the kernel launched just before us
    if (t >= p.T) return;
    auto ch = chain_begin<STATS, GAIN, BIQUAD>(p, t);
    const float* row = p.in + static_cast<size_t>(t) * p.B;
    float* orow = p.out + static_cast<size_t>(t) * p.B;
    float s = lane < p.B ? row[lane] : 0.0f;
    for (int n0 = 0; n0 < p.B; n0 += 32) {
        const int cnt = min(32, p.B - n0);
        const float cur = s;
        if (n0 + 32 + lane < p.B) s = row[n0 + 32 + lane];  // next line in flight during the chain
        float mine = 0.0f;
#pragma unroll 8
        for (int i = 0; i < cnt; ++i) {
            const float y = ch.step(__shfl_sync(0xffffffffu, cur, i));
            if (i == lane) mine = y;
        }
        if (lane < cnt) orow[n0 + lane] = mine;
    }
    if (lane == 0) chain_end(p, t, ch);
}

template <bool STATS, bool GAIN, bool BIQUAD>
__global__ void __launch_bounds__(kStripThreads) strip_cols_kernel(StripParams p) {
    const int t = blockIdx.x * kStripThreads + threadIdx.x;
    pdl_launch_dependents();
    pdl_wait_primary();
    if (t >= p.T) return;
    auto ch = chain_begin<STATS, GAIN, BIQUAD>(p, t);
    const float* col = p.in + p.col0 + t;
    float* ocol = p.out + p.col0 + t;
    const size_t ld = static_cast<size_t>(p.ld);
    float cur[kStripAhead], nxt[kStripAhead];
#pragma unroll
    for (int i = 0; i < kStripAhead; ++i) nxt[i] = i < p.B ? col[i * ld] : 0.0f;
    for (int n0 = 0; n0 < p.B; n0 += kStripAhead) {
#pragma unroll
        for (int i = 0; i < kStripAhead; ++i) cur[i] = nxt[i];
#pragma unroll
        for (int i = 0; i < kStripAhead; ++i)
            if (n0 + kStripAhead + i < p.B) nxt[i] = col[(n0 + kStripAhead + i) * ld];
#pragma unroll
        for (int i = 0; i < kStripAhead; ++i)
            if (n0 + i < p.B) ocol[(n0 + i) * ld] = ch.step(cur[i]);
    }
    chain_end(p, t, ch);
}

template <bool STATS, bool GAIN, bool BIQUAD>
cudaError_t launch_t(const StripParams& p, cudaStream_t st) {
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cfg.stream = st;
    if (p.sample_major) {
        cfg.gridDim = dim3((p.T + kStripThreads - 1) / kStripThreads);
        cfg.blockDim = dim3(kStripThreads);
        return cudaLaunchKernelEx(&cfg, strip_cols_kernel<STATS, GAIN, BIQUAD>, p);
    }
    cfg.gridDim = dim3((p.T + kStripWarps - 1) / kStripWarps);
    cfg.blockDim = dim3(kStripWarps * 32);
    return cudaLaunchKernelEx(&cfg, strip_rows_kernel<STATS, GAIN, BIQUAD>, p);
}

}  // namespace

cudaError_t launch_strip(const StripParams& p, cudaStream_t st) {
    const bool s = p.ops & B200CONV_STRIP_STATS, g = p.ops & B200CONV_STRIP_GAIN, q = p.ops & B200CONV_STRIP_BIQUAD;
    switch ((s ? 1 : 0) | (g ? 2 : 0) | (q ? 4 : 0)) {
        case 1: return launch_t<true, false, false>(p, st);
        case 2: return launch_t<false, true, false>(p, st);
        case 3: return launch_t<true, true, false>(p, st);
        case 4: return launch_t<false, false, true>(p, st);
        case 5: return launch_t<true, false, true>(p, st);
        case 6: return launch_t<false, true, true>(p, st);
        case 7: return launch_t<true, true, true>(p, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace b200conv
