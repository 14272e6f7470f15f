// conv_oracle.cpp — CPU restatement of the reference's conv1d validation path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product path (gpuaudiobench_b200/, include/,
// the gpubench host binary) links, loads or calls this file.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, and there
// only as the checker / the timed CPU baseline.
//
// Parity pin: the reference ships no golden vectors or tests for this path, so
// this restatement is pinned two ways:
//   (1) bit-for-bit against oracle/_ref/libgpuab_ref.so, which is the reference's OWN compiled
//       functions (built by oracle/Makefile from /root/reference/cuda/*.cu where they lie);
//   (2) against the golden fixtures in tests/golden/ generated from (1) by
//       tests/golden/make_golden.py (they travel to the GPU box; /root/reference does not).
//
// Build: g++ -std=c++17 -O2 -ffp-contract=off (no -march=native, no -ffast-math): -O0 and -O2
// give identical bits, -O3 -march=native does not.
//
// Every function cites the reference lines it follows (paths relative to /root/reference/).

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <random>
#include <thread>
#include <vector>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

extern "C" {

// ---------------------------------------------------------------------------------------------
// Input signal.  cuda/bench_utils.cu:238-245 (generateRandomAudioData), seeded 42 from
// cuda/bench_base.cu:44-49.  One sequential draw over the flat [T][B] array, libstdc++
// mt19937 + uniform_real_distribution<float>(-1, 1).  The std classes are the normative
// generator; they are used directly rather than re-derived.
// ---------------------------------------------------------------------------------------------
void oracle_generate_input(float* buf, size_t count, unsigned seed) {
    std::mt19937 engine(seed);
    std::uniform_real_distribution<float> uni(-1.0f, 1.0f);
    for (size_t n = 0; n < count; ++n) buf[n] = uni(engine);
}

// ---------------------------------------------------------------------------------------------
// Impulse responses, direct-form variant.  cuda/bench_conv1d.cu:159-178: Hamming-windowed sinc,
// per-track cutoff 0.1 + 0.05 t/T, centre tap L/2, scaled by 1/L, everything in float with a
// float PI constant.  Tracks [t_begin, t_end) of a T_total-track job are written to
// h[(t - t_begin) * L + k]: the cutoff uses the GLOBAL track index.
// ---------------------------------------------------------------------------------------------
void oracle_generate_ir_direct(float* h, int t_begin, int t_end, int T_total, int L) {
    const float PI = 3.14159265358979323846f;
    for (int t = t_begin; t < t_end; ++t) {
        float* row = h + static_cast<size_t>(t - t_begin) * L;
        for (int k = 0; k < L; ++k) {
            float freq = 0.1f + 0.05f * static_cast<float>(t) / static_cast<float>(T_total);
            float tt = static_cast<float>(k) - static_cast<float>(L) / 2.0f;
            float window =
                0.54f - 0.46f * cosf(2.0f * PI * static_cast<float>(k) / static_cast<float>(L - 1));
            float sinc = (tt == 0.0f) ? 1.0f : sinf(2.0f * PI * freq * tt) / (2.0f * PI * freq * tt);
            row[k] = window * sinc / static_cast<float>(L);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Impulse responses, FFT-variant generator.  cuda/bench_conv1d_accel.cu:152-165: the same text
// with the double constant M_PI, so 2.0f*M_PI*... sub-expressions are evaluated in double and
// narrowed at the cosf/sinf argument and at the float assignments.  ~49 % of taps differ from
// the direct variant by <= 1.2e-10.
// ---------------------------------------------------------------------------------------------
void oracle_generate_ir_accel(float* h, int t_begin, int t_end, int T_total, int L) {
    for (int t = t_begin; t < t_end; ++t) {
        float* row = h + static_cast<size_t>(t - t_begin) * L;
        for (int k = 0; k < L; ++k) {
            float freq = 0.1f + 0.05f * (float)t / (float)T_total;
            float tt = (float)k - (float)L / 2.0f;
            float window = 0.54f - 0.46f * cosf(2.0f * M_PI * (float)k / (float)(L - 1));
            float sinc = (tt == 0.0f) ? 1.0f : sinf(2.0f * M_PI * freq * tt) / (2.0f * M_PI * freq * tt);
            row[k] = window * sinc / (float)L;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Oracle R1.  cuda/bench_conv1d.cu:188-208 (Conv1DBenchmark::conv1DCPUReference; GPU twin :7-27).
//   y[t*B+i] = sum_{j<L, 0 <= t*B+i-j < T*B} h[t*L+j] * x[t*B+i-j]
// fp32, accumulator starts at 0.0f, `samp += h*x`, j ascending, output track-major.  The bound
// is on the FLAT index, so track t's history is the tail of tracks < t.
// Computes tracks [t_begin, t_end) with the full-job semantics (x and h are the full arrays).
// ---------------------------------------------------------------------------------------------
void oracle_conv1d_r1_range(const float* x, const float* h, float* y, int L, int B, int T,
                            int t_begin, int t_end) {
    const int total = T * B;
    for (int t = t_begin; t < t_end; ++t) {
        for (int i = 0; i < B; ++i) {
            float samp = 0.0f;
            for (int j = 0; j < L; ++j) {
                int xi = t * B + i - j;
                if (xi >= 0 && xi < total) {
                    samp += h[t * L + j] * x[xi];
                }
            }
            y[t * B + i] = samp;
        }
    }
}

void oracle_conv1d_r1(const float* x, const float* h, float* y, int L, int B, int T) {
    oracle_conv1d_r1_range(x, h, y, L, B, T, 0, T);
}

// ---------------------------------------------------------------------------------------------
// Oracle R2.  cuda/bench_conv1d_accel.cu:234-252 (== metal-swift Convolution1DBaseBenchmark.swift
// :94-115 == webgpu Convolution1DBenchmark.js:137-159).
//   y[T*n+t] = sum_{k<L, 0 <= n-k < B} x[t*B+n-k] * h[t*L+k]
// fp32, `+= x*h`, k ascending, per-track zero history, output SAMPLE-MAJOR (interleaved).
// 64-bit indexing so the streaming form (T=1, B = whole stream) can be long.
// ---------------------------------------------------------------------------------------------
void oracle_conv1d_r2_range(const float* x, const float* h, float* y, int L, int64_t B, int T,
                            int t_begin, int t_end) {
    for (int t = t_begin; t < t_end; ++t) {
        for (int64_t n = 0; n < B; ++n) {
            float acc = 0.0f;
            for (int k = 0; k < L; ++k) {
                int64_t xi = n - k;
                if (xi >= 0 && xi < B) {
                    float xv = x[static_cast<int64_t>(t) * B + xi];
                    float hv = h[static_cast<int64_t>(t) * L + k];
                    acc += xv * hv;
                }
            }
            y[static_cast<int64_t>(T) * n + t] = acc;
        }
    }
}

void oracle_conv1d_r2(const float* x, const float* h, float* y, int L, int B, int T) {
    oracle_conv1d_r2_range(x, h, y, L, B, T, 0, T);
}

// ---------------------------------------------------------------------------------------------
// Streaming oracle: R2 with track_count = 1 and buffer_size = the whole
// stream gives y[n] = sum_k h[k] x[n-k] for every n of M consecutive blocks with true history.
// Same loop, same order; the k-loop is clipped to k <= n (the skipped iterations are exactly
// the ones R2's bounds test rejects), so values are bit-identical to oracle_conv1d_r2(T=1).
// ---------------------------------------------------------------------------------------------
void oracle_stream(const float* x, const float* h, float* y, int L, int64_t nsamples) {
    for (int64_t n = 0; n < nsamples; ++n) {
        float acc = 0.0f;
        int kmax = static_cast<int>(std::min<int64_t>(L - 1, n));
        for (int k = 0; k <= kmax; ++k) acc += x[n - k] * h[k];
        y[n] = acc;
    }
}

// ---------------------------------------------------------------------------------------------
// FFT1D oracle.  cuda/bench_fft.cu:149-168 (FFTBenchmark::cpuFFTReference):
// naive O(N^2) DFT, k = 0..N/2, everything in float INCLUDING the angle -2*PI*k*n/size (so the
// oracle itself loses ~2.4e-4 rad at k*n ~ 5e5: it is a coarse reference, the fp64 DFT in the tests
// is the truth), `sum += input[n] * cosf/sinf(angle)`.
// ---------------------------------------------------------------------------------------------
void oracle_fft_reference(const float* input, float* real_output, float* imag_output, int size) {
    const float PI = 3.14159265358979323846f;
    for (int k = 0; k < size / 2 + 1; k++) {
        float sum_real = 0.0f, sum_imag = 0.0f;
        for (int n = 0; n < size; n++) {
            float angle = -2.0f * PI * k * n / size;
            sum_real += input[n] * cosf(angle);
            sum_imag += input[n] * sinf(angle);
        }
        real_output[k] = sum_real;
        imag_output[k] = sum_imag;
    }
}

// ---------------------------------------------------------------------------------------------
// Validation metrics.
// Absolute: cuda/bench_base.cu:193-222 (compareWithReference) — float running sum of |diff|,
// max |diff|, count of elements over tolerance.  Conv1D uses tol 1e-3 (bench_conv1d.cu:108).
// Relative: cuda/bench_conv1d_accel.cu:312-336 — |g-c|/|c| (absolute when c == 0), float sums,
// pass iff max < 1e-3.
// ---------------------------------------------------------------------------------------------
void oracle_compare_abs(const float* gpu, const float* cpu, size_t n, float tol, float* max_err,
                        float* mean_err, int* n_over) {
    float sum = 0.0f, mx = 0.0f;
    int over = 0;
    for (size_t i = 0; i < n; ++i) {
        float d = std::abs(gpu[i] - cpu[i]);
        sum += d;
        mx = std::max(mx, d);
        if (d > tol) ++over;
    }
    *max_err = mx;
    *mean_err = sum / static_cast<float>(n);
    *n_over = over;
}

void oracle_compare_rel(const float* gpu, const float* cpu, size_t n, float* max_err,
                        float* mean_err) {
    float mx = 0.0f, total = 0.0f;
    for (size_t i = 0; i < n; ++i) {
        float err = fabsf(gpu[i] - cpu[i]);
        float rel = cpu[i] != 0 ? err / fabsf(cpu[i]) : err;
        mx = fmaxf(mx, rel);
        total += rel;
    }
    *max_err = mx;
    *mean_err = total / n;
}

// ---------------------------------------------------------------------------------------------
// Latency statistics.  cuda/bench_utils.cu:358-414 (calculateStatistics: mean, median, sample
// std-dev, min, max, linear-interpolated p95/p99) and cuda/globals.cu:83-89 (CSV/JSON writers:
// nearest-rank sorted[n*q], deadline 1000*B/fs, meets iff p99 <= threshold).
// out8 = {mean, median, std, min, max, p95, p99, count}.
// ---------------------------------------------------------------------------------------------
void oracle_statistics(const float* lat, size_t n, float* out8) {
    std::fill(out8, out8 + 8, 0.0f);
    if (n == 0) return;
    std::vector<float> s(lat, lat + n);
    std::sort(s.begin(), s.end());
    float sum = std::accumulate(lat, lat + n, 0.0f);
    float mean = sum / static_cast<float>(n);
    size_t mid = n / 2;
    float median = (n % 2 == 0) ? (s[mid - 1] + s[mid]) / 2.0f : s[mid];
    float var = 0.0f;
    for (size_t i = 0; i < n; ++i) {
        float d = lat[i] - mean;
        var += d * d;
    }
    var /= static_cast<float>(n - 1);
    auto pct = [&](float p) {
        float idx = p / 100.0f * static_cast<float>(n - 1);
        size_t lo = static_cast<size_t>(std::floor(idx));
        size_t hi = static_cast<size_t>(std::ceil(idx));
        if (lo == hi) return s[lo];
        float w = idx - static_cast<float>(lo);
        return s[lo] * (1.0f - w) + s[hi] * w;
    };
    out8[0] = mean;
    out8[1] = median;
    out8[2] = std::sqrt(var);
    out8[3] = s.front();
    out8[4] = s.back();
    out8[5] = pct(95.0f);
    out8[6] = pct(99.0f);
    out8[7] = static_cast<float>(n);
}

// out5 = {p50, p95, p99, threshold_ms, meets_deadline(0/1)}
void oracle_nearest_rank(const float* lat, size_t n, int bufsize, int fs, float* out5) {
    std::vector<float> s(lat, lat + n);
    std::sort(s.begin(), s.end());
    out5[0] = s[n * 0.50];
    out5[1] = s[n * 0.95];
    out5[2] = s[n * 0.99];
    out5[3] = 1000.0f * bufsize / fs;
    out5[4] = (out5[2] <= out5[3]) ? 1.0f : 0.0f;
}

// FNV-1a-64 over raw bytes: the tripwire hash of SURVEY.md App. A.3.
uint64_t oracle_fnv1a64(const void* data, size_t nbytes) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    uint64_t hsh = 1469598103934665603ull;
    for (size_t i = 0; i < nbytes; ++i) {
        hsh ^= p[i];
        hsh *= 1099511628211ull;
    }
    return hsh;
}

// ---------------------------------------------------------------------------------------------
// Channel-strip stages: the CPU references of the Gain, GainStats and IIRFilter
// plugins.  Track-major [T][B] in and out.
// ---------------------------------------------------------------------------------------------

// cuda/bench_gain.cu:90-92 (GainBenchmark::calculateCPUReference; GAIN_VALUE = 2.0f,
// cuda/benchmark_constants.cuh:6).
void oracle_gain(const float* x, float* y, size_t n, float gain) {
    for (size_t i = 0; i < n; ++i) y[i] = gain * x[i];
}

// cuda/bench_gainstats.cu:121-142 (GAINSTATS_GAIN = 0.5f, benchmark_constants.cuh:7): output
// gain * x; per track the running float sum / B and the maximum of the INPUT, stats [T][2].
void oracle_gainstats(const float* x, float* y, float* stats, size_t T, size_t B, float gain) {
    for (size_t i = 0; i < T * B; ++i) y[i] = gain * x[i];
    for (size_t t = 0; t < T; ++t) {
        float mean = 0.0f;
        float maxVal = -1e9f;
        for (size_t i = 0; i < B; ++i) {
            float samp = x[t * B + i];
            mean += samp;
            if (samp > maxVal) maxVal = samp;
        }
        mean /= B;
        stats[2 * t + 0] = mean;
        stats[2 * t + 1] = maxVal;
    }
}

// cuda/bench_iir.cu:205-228 (calculateButterworthCoefficients; the plugin passes 0.25f, :162).
// out5 = b0, b1, b2, a1, a2 after normalisation by a0.
void oracle_butterworth(float normalized_frequency, float* out5) {
    const float PI = 3.14159265358979323846f;
    float omega = 2.0f * PI * normalized_frequency;
    float cos_omega = cosf(omega);
    float sin_omega = sinf(omega);
    float alpha = sin_omega / (2.0f * 0.707f);
    float b0 = (1.0f - cos_omega) / 2.0f;
    float b1 = 1.0f - cos_omega;
    float b2 = (1.0f - cos_omega) / 2.0f;
    float a0 = 1.0f + alpha;
    float a1 = -2.0f * cos_omega;
    float a2 = 1.0f - alpha;
    out5[0] = b0 / a0;
    out5[1] = b1 / a0;
    out5[2] = b2 / a0;
    out5[3] = a1 / a0;
    out5[4] = a2 / a0;
}

// cuda/bench_iir.cu:176-203 (iirFilterCPUReference): Direct Form II biquad per track, state
// [T][2] = (z1, z2) read at entry and written back.  coeffs_stride = 0: one coefficient set for all
// tracks (the reference's case); 5: coeffs is [T][5].
void oracle_iir(const float* x, float* y, const float* coeffs, int coeffs_stride, float* state, int T, int B) {
    for (int track = 0; track < T; ++track) {
        const float* c = coeffs + static_cast<size_t>(track) * coeffs_stride;
        const float b0 = c[0], b1 = c[1], b2 = c[2], a1 = c[3], a2 = c[4];
        float z1 = state[track * 2];
        float z2 = state[track * 2 + 1];
        const size_t start = static_cast<size_t>(track) * B;
        for (int i = 0; i < B; ++i) {
            float xin = x[start + i];
            float w = xin - a1 * z1 - a2 * z2;
            float out = b0 * w + b1 * z1 + b2 * z2;
            z2 = z1;
            z1 = w;
            y[start + i] = out;
        }
        state[track * 2] = z1;
        state[track * 2 + 1] = z2;
    }
}

// The engine's strip = the three stages chained in the order include/b200conv.h states: statistics of
// the input, gain, biquad.  ops bits: 1 STATS, 2 GAIN, 4 BIQUAD.  gains: [T] or null (then `gain`).
void oracle_strip(const float* x, float* y, int T, int B, unsigned ops, float gain, const float* gains,
                  const float* coeffs, int coeffs_stride, float* state, float* stats) {
    std::vector<float> tmp(static_cast<size_t>(B)), sink(static_cast<size_t>(B));
    for (int t = 0; t < T; ++t) {
        const float* xt = x + static_cast<size_t>(t) * B;
        float* yt = y + static_cast<size_t>(t) * B;
        std::copy(xt, xt + B, tmp.begin());
        if (ops & 1u) oracle_gainstats(xt, sink.data(), stats + 2 * t, 1, static_cast<size_t>(B), 1.0f);
        if (ops & 2u) oracle_gain(xt, tmp.data(), static_cast<size_t>(B), gains ? gains[t] : gain);
        if (ops & 4u)
            oracle_iir(tmp.data(), yt, coeffs + static_cast<size_t>(t) * coeffs_stride, 0, state + 2 * t, 1, B);
        else
            std::copy(tmp.begin(), tmp.end(), yt);
    }
}

// ---------------------------------------------------------------------------------------------
// CPU-baseline timing legs (bench.py cpu_baseline / --impl reference when oracle/_ref is absent).
// The reference runs its oracle single-threaded inside setupBenchmark (bench_conv1d.cu:42-55);
// tracks are independent, so the threaded leg splits [0, T) into contiguous ranges.  Returns
// seconds of wall time (steady_clock), work = T*B*L loop iterations.
// ---------------------------------------------------------------------------------------------
double oracle_time_r1(const float* x, const float* h, float* y, int L, int B, int T, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    nthreads = std::min(nthreads, T);
    auto t0 = std::chrono::steady_clock::now();
    if (nthreads == 1) {
        oracle_conv1d_r1_range(x, h, y, L, B, T, 0, T);
    } else {
        std::vector<std::thread> pool;
        for (int w = 0; w < nthreads; ++w) {
            int a = static_cast<int>(static_cast<int64_t>(T) * w / nthreads);
            int b = static_cast<int>(static_cast<int64_t>(T) * (w + 1) / nthreads);
            pool.emplace_back([=] { oracle_conv1d_r1_range(x, h, y, L, B, T, a, b); });
        }
        for (auto& th : pool) th.join();
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

double oracle_time_r2(const float* x, const float* h, float* y, int L, int B, int T, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    nthreads = std::min(nthreads, T);
    auto t0 = std::chrono::steady_clock::now();
    if (nthreads == 1) {
        oracle_conv1d_r2_range(x, h, y, L, B, T, 0, T);
    } else {
        std::vector<std::thread> pool;
        for (int w = 0; w < nthreads; ++w) {
            int a = static_cast<int>(static_cast<int64_t>(T) * w / nthreads);
            int b = static_cast<int>(static_cast<int64_t>(T) * (w + 1) / nthreads);
            pool.emplace_back([=] { oracle_conv1d_r2_range(x, h, y, L, B, T, a, b); });
        }
        for (auto& th : pool) th.join();
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

int oracle_hardware_threads() {
    unsigned n = std::thread::hardware_concurrency();
    return n ? static_cast<int>(n) : 1;
}

}  // extern "C"
