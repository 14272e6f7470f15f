#pragma once
// Channel-strip plugins: "gain", "GainStats" and "IIRFilter" on the engine's strip kernels
// (b200conv_strip_process) — SURVEY.md §8(f) #4.  Drop-ins for the reference's GainBenchmark
// (cuda/bench_gain.cuh:8-35), GainStatsBenchmark (cuda/bench_gainstats.cuh:7-48) and IIRBenchmark
// (cuda/bench_iir.cuh:7-72): same class names, constructors, overrides, registry names, constants and
// validation tolerances.  The three reference classes repeat one pattern (upload, one-thread-per-track
// kernel, download, CPU loop); here they share one base that owns the strip description and buffers.
#include <vector>

#include "b200conv.h"
#include "bench_base.cuh"

// biquad coefficients, a0 normalised to 1 (cuda/bench_iir.cuh:8-11) — same field order as b200conv_strip::biquad
struct IIRCoefficients {
    float b0, b1, b2;
    float a1, a2;
};

class ChannelStripBenchmark : public GPUABenchmark {
public:
    ~ChannelStripBenchmark() override;
    void setupBenchmark() override;
    void runKernel() override { performBenchmarkIteration(); }
    void performBenchmarkIteration() override;
    void validate(ValidationData& validation_data) override;

    // views for the C binding / tests (valid after an iteration)
    const float* cpuReference() const { return cpu_output_.data(); }
    const float* hostStats() const { return h_stats_; }
    const float* hostState() const { return h_state_; }
    const float* cpuStats() const { return cpu_stats_.data(); }
    const float* cpuState() const { return cpu_state_.data(); }
    const IIRCoefficients& coefficients() const { return coeffs_; }
    bool lastRunBitExact() const { return bit_exact_; }

protected:
    ChannelStripBenchmark(const std::string& name, size_t buffer_size, size_t track_count, uint32_t ops, float gain,
                          float output_tolerance, float aux_tolerance);
    // CPU loops of the reference, one pass over the plugin's input from the given state
    void cpuPass(std::vector<float>& out, std::vector<float>& stats, std::vector<float>& state) const;

    uint32_t ops_;
    float gain_;
    float output_tolerance_, aux_tolerance_;
    IIRCoefficients coeffs_{1.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    bool validation_enabled_ = true;

private:
    float* h_stats_ = nullptr;  // pinned [T][2] mean, max
    float* h_state_ = nullptr;  // pinned [T][2] z1, z2
    float* d_stats_ = nullptr;
    float* d_state_ = nullptr;
    float* d_coeffs_ = nullptr;
    std::vector<float> cpu_output_, cpu_stats_, cpu_state_;
    int iterations_done_ = 0;
    bool bit_exact_ = false;
};

class GainBenchmark : public ChannelStripBenchmark {
public:
    explicit GainBenchmark(size_t buffer_size = BUFSIZE, size_t track_count = NTRACKS, bool enable_validation = true);
};

class GainStatsBenchmark : public ChannelStripBenchmark {
public:
    static const int NSTATS = 2;  // mean, max
    GainStatsBenchmark(size_t buffer_size = BUFSIZE, size_t track_count = NTRACKS);
};

class IIRBenchmark : public ChannelStripBenchmark {
public:
    static const int STATES_PER_TRACK = 2;  // z1, z2
    IIRBenchmark(size_t buffer_size = BUFSIZE, size_t track_count = NTRACKS);
    // 2nd-order Butterworth low-pass, fc = normalized_frequency * fs (cuda/bench_iir.cu:205-228)
    static IIRCoefficients calculateButterworthCoefficients(float normalized_frequency);
};
