#pragma once
// Shared pieces of the two convolution plugins: RAII handle over the b200conv C ABI, the
// reference-shaped IR generators, the plugins' CPU validation loops, and the stated-tolerance
// accuracy metrics (SNR in dB, max-abs relative to the reference peak).
#include <cstddef>
#include <string>
#include <vector>

#include "b200conv.h"

namespace ConvCommon {

// Owns one b200conv engine; throws std::runtime_error carrying b200conv_last_error() on failure.
class Engine {
public:
    Engine() = default;
    ~Engine();
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;
    void create(b200conv_algo algo, b200conv_layout layout, size_t tracks, size_t block, int ir_len);
    void loadIR(const float* host_ir);
    void primeHistory(const float* host_hist);  // nullptr -> zero history
    void reset();
    void process(const float* d_in, float* d_out, float* d_mix, bool advance_state, cudaStream_t stream);
    b200conv_info info();
    bool valid() const { return handle_ != nullptr; }

private:
    b200conv_engine* handle_ = nullptr;
    static void check(int rc, const char* what);
};

// The same over b200conv_group_*: one engine per GPU in this process, host buffers in and out.
class Group {
public:
    Group() = default;
    ~Group();
    Group(const Group&) = delete;
    Group& operator=(const Group&) = delete;
    void create(b200conv_algo algo, b200conv_layout layout, size_t total_tracks, size_t block, int ir_len, int n_gpus);
    void loadIR(const float* host_ir);
    void primeHistory(const float* host_hist);
    void reset();
    void processHost(const float* h_in, float* h_out, float* h_mix, bool advance_state);
    bool valid() const { return handle_ != nullptr; }

private:
    b200conv_group* handle_ = nullptr;
    static void check(int rc, const char* what);
};

enum class IRVariant { DIRECT_FLOAT_PI, ACCEL_DOUBLE_PI };

// Hamming-windowed sinc, cutoff BASE + RANGE * t / T, centre L/2, scaled 1/L.
// DIRECT_FLOAT_PI follows cuda/bench_conv1d.cu:159-178 (all-float, float PI constant);
// ACCEL_DOUBLE_PI follows cuda/bench_conv1d_accel.cu:152-165 (double M_PI in the products).
void generateImpulseResponses(float* h, size_t track_count, int ir_len, IRVariant variant);

// The plugins' CPU validation references (run once in setupBenchmark, like the reference does):
//   flatHistory : y[t*B+i] = sum_j h[t*L+j] x[t*B+i-j] over the FLAT input index (track-major out)
//                 — cuda/bench_conv1d.cu:188-208
//   zeroHistory : y[T*n+t] = sum_{k<=n} x[t*B+n-k] h[t*L+k]              (sample-major out)
//                 — cuda/bench_conv1d_accel.cu:234-252
void cpuConvFlatHistory(const float* x, const float* h, float* y, int L, int B, int T);
void cpuConvZeroHistory(const float* x, const float* h, float* y, int L, int B, int T);

struct Accuracy {
    double snr_db;        // 10 log10(sum ref^2 / sum (got-ref)^2)
    double max_abs_err;
    double ref_peak;      // max |ref|
};
Accuracy measureAccuracy(const float* got, const float* ref, size_t n);
std::string describeAccuracy(const Accuracy& a, double min_snr_db, double max_rel_to_peak);

}  // namespace ConvCommon
