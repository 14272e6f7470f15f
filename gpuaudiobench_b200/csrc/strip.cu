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
//   track-major  [T][B]:  one WARP per track.  The row is staged in shared memory with 16-byte cp.async
//       (1024 samples per chunk: one exposed memory latency per chunk), read back as broadcast float4s,
//       8 samples per loop trip; every lane runs the identical chain (no divergence), lane 0 records the
//       results in shared memory and the warp writes them out coalesced.  (The first version walked the
//       row with one shuffle per sample: in-order issue put the shuffle latency on every sample, 16 us
//       per call whatever the operation.)
//   sample-major [B][Tg]: one THREAD per track, lanes = adjacent tracks, so a row access is coalesced;
//       128-row chunks are staged in shared memory with cp.async (LDGSTS), double buffered.
//   gain alone has no chain at all: a plain float4 elementwise pass.
// Arithmetic uses __fmul_rn/__fadd_rn/__fsub_rn in the reference's evaluation order, so outputs, filter
// state and statistics are BIT-IDENTICAL to the CPU loops (bench_gain.cu:90-92, bench_gainstats.cu:
// 121-142, bench_iir.cu:176-203; the host build does not contract a*b+c).
#include "strip.cuh"

#include <algorithm>
#include <cstdint>

#include "../../include/b200conv.h"
#include "common.cuh"

namespace b200conv {

namespace {

constexpr int kStripWarps = 4;    // warps (= tracks) per CTA, track-major kernel
constexpr int kStripThreads = 32; // threads (= tracks) per CTA, sample-major kernel (one warp: 32 KB of staging)
constexpr int kStripRows = 128;   // rows per shared-memory chunk (16 KB, double buffered), sample-major kernel
constexpr int kStripChunk = 1024; // samples per shared-memory chunk and warp (4 KB in + 4 KB out), track-major kernel

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

// Asynchronous global->shared copies (LDGSTS): the data never passes through registers, and completion
// is tracked per commit group, not by the load scoreboards a register prefetch ring shares with every
// younger load (measured: a rolling 64-row register ring cost one full L2 round trip per SAMPLE).
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// 8 consecutive samples through the chain.  The loop bodies below are kept THIS small on purpose: a
// single warp running fully unrolled straight-line code (the second version: 64 KB of SASS per kernel) is
// bound by instruction fetch, not by the chain.
template <typename ChainT>
__device__ __forceinline__ void chain8(ChainT& ch, const float (&x)[8], float (&y)[8]) {
#pragma unroll
    for (int u = 0; u < 8; ++u) y[u] = ch.step(x[u]);
}

template <bool STATS, bool GAIN, bool BIQUAD>
__global__ void __launch_bounds__(kStripWarps * 32) strip_rows_kernel(StripParams p, int vec) {
    __shared__ __align__(16) float xin[kStripWarps][kStripChunk];
    __shared__ __align__(16) float yout[kStripWarps][kStripChunk];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = blockIdx.x * kStripWarps + warp;
    pdl_launch_dependents();
    pdl_wait_primary();  // the input is the output of the kernel launched just before us
    if (t >= p.T) return;  // warp-uniform; no block-wide barrier below
    auto ch = chain_begin<STATS, GAIN, BIQUAD>(p, t);
    const float* row = p.in + static_cast<size_t>(t) * p.B;
    float* orow = p.out + static_cast<size_t>(t) * p.B;
    float* xs = xin[warp];
    float* ys = yout[warp];
    for (int c0 = 0; c0 < p.B; c0 += kStripChunk) {
        const int len = min(kStripChunk, p.B - c0);
        // stage the chunk: one exposed memory latency per 1024 samples
        if (vec) {
            for (int i = lane * 4; i < len; i += 128) cp_async16(xs + i, row + c0 + i);
        } else {
            for (int i = lane; i < len; i += 32) cp_async4(xs + i, row + c0 + i);
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
        // every lane runs the identical chain on broadcast reads (no divergence); lane 0 records it
        int n = 0;
        for (; n + 8 <= len; n += 8) {
            const float4 a = *reinterpret_cast<const float4*>(xs + n);
            const float4 b = *reinterpret_cast<const float4*>(xs + n + 4);
            const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            float y[8];
            chain8(ch, x, y);
            if (lane == 0) {
                *reinterpret_cast<float4*>(ys + n) = make_float4(y[0], y[1], y[2], y[3]);
                *reinterpret_cast<float4*>(ys + n + 4) = make_float4(y[4], y[5], y[6], y[7]);
            }
        }
        for (; n < len; ++n) {
            const float y = ch.step(xs[n]);
            if (lane == 0) ys[n] = y;
        }
        __syncwarp();
        if (vec) {
            for (int i = lane * 4; i < len; i += 128)
                *reinterpret_cast<float4*>(orow + c0 + i) = *reinterpret_cast<const float4*>(ys + i);
        } else {
            for (int i = lane; i < len; i += 32) orow[c0 + i] = ys[i];
        }
        __syncwarp();
    }
    if (lane == 0) chain_end(p, t, ch);
}

template <bool STATS, bool GAIN, bool BIQUAD>
__global__ void __launch_bounds__(kStripThreads) strip_cols_kernel(StripParams p) {
    // [2 buffers][kStripRows][32 tracks]; lane t reads word n*32 + t: conflict free
    __shared__ float tile[2][kStripRows * 32];
    const int lane = threadIdx.x;
    const int t = blockIdx.x * kStripThreads + threadIdx.x;
    pdl_launch_dependents();
    pdl_wait_primary();
    if (t >= p.T) return;  // a lane without a track neither copies nor computes (no barrier below)
    auto ch = chain_begin<STATS, GAIN, BIQUAD>(p, t);
    const float* col = p.in + p.col0 + t;
    float* ocol = p.out + p.col0 + t;
    const size_t ld = static_cast<size_t>(p.ld);
    auto fetch = [&](int buf, int n0) {
        float* dst = &tile[buf][lane];
        const int rows = min(kStripRows, p.B - n0);
        const float* src = col + static_cast<size_t>(n0) * ld;
        for (int r = 0; r < rows; ++r, dst += 32, src += ld) cp_async4(dst, src);
        cp_async_commit();
    };
    fetch(0, 0);
    int buf = 0;
    for (int n0 = 0; n0 < p.B; n0 += kStripRows, buf ^= 1) {
        if (n0 + kStripRows < p.B) {
            fetch(buf ^ 1, n0 + kStripRows);  // next chunk in flight during this chunk's chain
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        const float* src = &tile[buf][lane];
        const int rows = min(kStripRows, p.B - n0);
        float* dst = ocol + static_cast<size_t>(n0) * ld;
        int r = 0;
        for (; r + 8 <= rows; r += 8, src += 8 * 32, dst += 8 * ld) {
            float x[8], y[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) x[u] = src[u * 32];
            chain8(ch, x, y);
#pragma unroll
            for (int u = 0; u < 8; ++u) dst[u * ld] = y[u];
        }
        for (; r < rows; ++r, src += 32, dst += ld) *dst = ch.step(*src);
    }
    chain_end(p, t, ch);
}

// gain alone: elementwise.  Track-major: float4 over the flat [T*B] array when everything is 16-B aligned.
__global__ void __launch_bounds__(256) gain_flat_kernel(StripParams p, size_t n_total, int vec) {
    pdl_launch_dependents();
    pdl_wait_primary();
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (vec) {
        const float4* in4 = reinterpret_cast<const float4*>(p.in);
        float4* out4 = reinterpret_cast<float4*>(p.out);
        const int b4 = p.B / 4;
        for (; i < n_total / 4; i += stride) {
            const float g = p.gains ? p.gains[i / b4] : p.gain;
            float4 v = in4[i];
            v.x = __fmul_rn(g, v.x);
            v.y = __fmul_rn(g, v.y);
            v.z = __fmul_rn(g, v.z);
            v.w = __fmul_rn(g, v.w);
            out4[i] = v;
        }
    } else {
        for (; i < n_total; i += stride) p.out[i] = __fmul_rn(p.gains ? p.gains[i / p.B] : p.gain, p.in[i]);
    }
}

// sample-major: element (n, t) at n*ld + col0 + t; threads run along t (coalesced)
__global__ void __launch_bounds__(256) gain_cols_kernel(StripParams p) {
    pdl_launch_dependents();
    pdl_wait_primary();
    const size_t total = static_cast<size_t>(p.B) * p.T;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t n = i / p.T;
        const int t = static_cast<int>(i - n * p.T);
        const size_t at = n * p.ld + p.col0 + t;
        p.out[at] = __fmul_rn(p.gains ? p.gains[t] : p.gain, p.in[at]);
    }
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
    const int vec = (p.B % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.in) | reinterpret_cast<uintptr_t>(p.out)) % 16 == 0);
    return cudaLaunchKernelEx(&cfg, strip_rows_kernel<STATS, GAIN, BIQUAD>, p, vec);
}

}  // namespace

static cudaError_t launch_gain_only(const StripParams& p, cudaStream_t st) {
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cfg.stream = st;
    cfg.blockDim = dim3(256);
    const size_t total = static_cast<size_t>(p.T) * p.B;
    if (p.sample_major) {
        cfg.gridDim = dim3(static_cast<unsigned>(std::min<size_t>((total + 255) / 256, 148 * 8)));
        return cudaLaunchKernelEx(&cfg, gain_cols_kernel, p);
    }
    const bool vec = (p.B % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.in) | reinterpret_cast<uintptr_t>(p.out)) % 16 == 0);
    const size_t items = vec ? total / 4 : total;
    cfg.gridDim = dim3(static_cast<unsigned>(std::max<size_t>(1, std::min<size_t>((items + 255) / 256, 148 * 8))));
    return cudaLaunchKernelEx(&cfg, gain_flat_kernel, p, total, vec ? 1 : 0);
}

cudaError_t launch_strip(const StripParams& p, cudaStream_t st) {
    const bool s = p.ops & B200CONV_STRIP_STATS, g = p.ops & B200CONV_STRIP_GAIN, q = p.ops & B200CONV_STRIP_BIQUAD;
    if (g && !s && !q) return launch_gain_only(p, st);
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
