#include "conv_common.cuh"

#include <cmath>
#include <cstdio>
#include <stdexcept>

#include "benchmark_constants.cuh"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace ConvCommon {

void Engine::check(int rc, const char* what) {
    if (rc != B200CONV_OK)
        throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + b200conv_last_error());
}

Engine::~Engine() { b200conv_destroy(handle_); }

void Engine::create(b200conv_algo algo, b200conv_layout layout, size_t tracks, size_t block, int ir_len) {
    b200conv_destroy(handle_);
    handle_ = nullptr;
    b200conv_config cfg{};
    cfg.abi_version = B200CONV_ABI_VERSION;
    cfg.device = 0;  // the reference never selects a device (main.cu:309): device 0
    cfg.tracks = static_cast<uint32_t>(tracks);
    cfg.track_offset = 0;
    cfg.total_tracks = static_cast<uint32_t>(tracks);
    cfg.block = static_cast<uint32_t>(block);
    cfg.ir_len = static_cast<uint32_t>(ir_len);
    cfg.algo = algo;
    cfg.out_layout = layout;
    check(b200conv_create(&cfg, &handle_), "b200conv_create");
}

void Engine::loadIR(const float* host_ir) { check(b200conv_load_ir(handle_, host_ir), "b200conv_load_ir"); }
void Engine::primeHistory(const float* host_hist) { check(b200conv_prime_history(handle_, host_hist), "b200conv_prime_history"); }
void Engine::reset() { check(b200conv_reset(handle_), "b200conv_reset"); }

void Engine::process(const float* d_in, float* d_out, float* d_mix, bool advance_state, cudaStream_t stream) {
    check(b200conv_process(handle_, d_in, d_out, d_mix, advance_state ? 0u : B200CONV_PEEK, stream), "b200conv_process");
}

b200conv_info Engine::info() {
    b200conv_info i{};
    check(b200conv_query(handle_, &i), "b200conv_query");
    return i;
}

void Group::check(int rc, const char* what) {
    if (rc != B200CONV_OK)
        throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + b200conv_group_last_error());
}

Group::~Group() { b200conv_group_destroy(handle_); }

void Group::create(b200conv_algo algo, b200conv_layout layout, size_t total_tracks, size_t block, int ir_len, int n_gpus) {
    b200conv_group_destroy(handle_);
    handle_ = nullptr;
    b200conv_config cfg{};
    cfg.abi_version = B200CONV_ABI_VERSION;
    cfg.tracks = static_cast<uint32_t>(total_tracks);
    cfg.block = static_cast<uint32_t>(block);
    cfg.ir_len = static_cast<uint32_t>(ir_len);
    cfg.algo = algo;
    cfg.out_layout = layout;
    check(b200conv_group_create(&cfg, n_gpus, &handle_), "b200conv_group_create");
}

void Group::loadIR(const float* host_ir) { check(b200conv_group_load_ir(handle_, host_ir), "b200conv_group_load_ir"); }
void Group::primeHistory(const float* host_hist) { check(b200conv_group_prime_history(handle_, host_hist), "b200conv_group_prime_history"); }
void Group::reset() { check(b200conv_group_reset(handle_), "b200conv_group_reset"); }
void Group::processHost(const float* h_in, float* h_out, float* h_mix, bool advance_state) {
    check(b200conv_group_process_host(handle_, h_in, h_out, h_mix, advance_state ? 0u : B200CONV_PEEK), "b200conv_group_process_host");
}

void generateImpulseResponses(float* h, size_t track_count, int ir_len, IRVariant variant) {
    using namespace BenchmarkConstants;
    const float PI = 3.14159265358979323846f;
    const float T = static_cast<float>(track_count);
    const float Lf = static_cast<float>(ir_len);
    for (size_t track = 0; track < track_count; ++track) {
        float* row = h + track * static_cast<size_t>(ir_len);
        const float freq = CONV1D_IR_BASE_FREQ + CONV1D_IR_FREQ_RANGE * static_cast<float>(track) / T;
        for (int k = 0; k < ir_len; ++k) {
            const float kf = static_cast<float>(k);
            const float tt = kf - Lf / 2.0f;
            float window, sinc;
            if (variant == IRVariant::DIRECT_FLOAT_PI) {
                window = HAMMING_WINDOW_A0 - HAMMING_WINDOW_A1 * cosf(2.0f * PI * kf / static_cast<float>(ir_len - 1));
                sinc = (tt == 0.0f) ? 1.0f : sinf(2.0f * PI * freq * tt) / (2.0f * PI * freq * tt);
            } else {  // products with the double constant are evaluated in double, narrowed at cosf/sinf and on assignment
                window = HAMMING_WINDOW_A0 - HAMMING_WINDOW_A1 * cosf(2.0f * M_PI * kf / static_cast<float>(ir_len - 1));
                sinc = (tt == 0.0f) ? 1.0f : sinf(2.0f * M_PI * freq * tt) / (2.0f * M_PI * freq * tt);
            }
            row[k] = window * sinc / Lf;
        }
    }
}

void cpuConvFlatHistory(const float* x, const float* h, float* y, int L, int B, int T) {
    const int total = T * B;
    for (int t = 0; t < T; ++t) {
        const float* taps = h + static_cast<size_t>(t) * L;
        for (int i = 0; i < B; ++i) {
            const int newest = t * B + i;  // flat index of x[n]; older samples may belong to earlier tracks
            float acc = 0.0f;
            for (int j = 0; j < L; ++j) {
                const int idx = newest - j;
                if (idx >= 0 && idx < total) acc += taps[j] * x[idx];
            }
            y[newest] = acc;
        }
    }
}

void cpuConvZeroHistory(const float* x, const float* h, float* y, int L, int B, int T) {
    for (int t = 0; t < T; ++t) {
        const float* sig = x + static_cast<size_t>(t) * B;
        const float* taps = h + static_cast<size_t>(t) * L;
        for (int n = 0; n < B; ++n) {
            float acc = 0.0f;
            for (int k = 0; k < L; ++k) {
                const int idx = n - k;
                if (idx >= 0 && idx < B) acc += sig[idx] * taps[k];
            }
            y[static_cast<size_t>(T) * n + t] = acc;
        }
    }
}

Accuracy measureAccuracy(const float* got, const float* ref, size_t n) {
    double sig = 0.0, noise = 0.0, worst = 0.0, peak = 0.0;
    for (size_t i = 0; i < n; ++i) {
        const double r = ref[i], d = static_cast<double>(got[i]) - r;
        sig += r * r;
        noise += d * d;
        worst = std::fmax(worst, std::fabs(d));
        peak = std::fmax(peak, std::fabs(r));
    }
    Accuracy a;
    a.snr_db = noise > 0.0 ? 10.0 * std::log10(sig / noise) : 300.0;
    a.max_abs_err = worst;
    a.ref_peak = peak;
    return a;
}

std::string describeAccuracy(const Accuracy& a, double min_snr_db, double max_rel_to_peak) {
    char buf[256];
    std::snprintf(buf, sizeof(buf), "SNR %.1f dB (need >= %.0f), max|err| %.3e = %.2e x peak %.3e (need <= %.0e)", a.snr_db,
                  min_snr_db, a.max_abs_err, a.ref_peak > 0 ? a.max_abs_err / a.ref_peak : 0.0, a.ref_peak, max_rel_to_peak);
    return buf;
}

}  // namespace ConvCommon
