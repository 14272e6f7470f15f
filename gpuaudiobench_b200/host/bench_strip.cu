#include "bench_strip.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>

#include "benchmark_constants.cuh"

ChannelStripBenchmark::ChannelStripBenchmark(const std::string& name, size_t buffer_size, size_t track_count, uint32_t ops,
                                             float gain, float output_tolerance, float aux_tolerance)
    : GPUABenchmark(name, buffer_size, track_count),
      ops_(ops),
      gain_(gain),
      output_tolerance_(output_tolerance),
      aux_tolerance_(aux_tolerance) {}

ChannelStripBenchmark::~ChannelStripBenchmark() {
    BenchmarkUtils::freeHostBuffers({h_stats_, h_state_});
    BenchmarkUtils::freeDeviceBuffers({d_stats_, d_state_, d_coeffs_});
}

void ChannelStripBenchmark::setupBenchmark() {
    const size_t T = getTrackCount();
    allocateBuffers(getTotalElements());
    h_stats_ = BenchmarkUtils::allocateHostBuffer<float>(2 * T, benchmark_name_ + " host stats buffer");
    h_state_ = BenchmarkUtils::allocateHostBuffer<float>(2 * T, benchmark_name_ + " host state buffer");
    d_stats_ = BenchmarkUtils::allocateDeviceBuffer<float>(2 * T, benchmark_name_ + " device stats buffer");
    d_state_ = BenchmarkUtils::allocateDeviceBuffer<float>(2 * T, benchmark_name_ + " device state buffer");
    d_coeffs_ = BenchmarkUtils::allocateDeviceBuffer<float>(5, benchmark_name_ + " device coefficients buffer");
    std::memset(h_stats_, 0, 2 * T * sizeof(float));
    std::memset(h_state_, 0, 2 * T * sizeof(float));
    CUDA_CHECK(cudaMemset(d_stats_, 0, 2 * T * sizeof(float)));
    CUDA_CHECK(cudaMemset(d_state_, 0, 2 * T * sizeof(float)));
    CUDA_CHECK(cudaMemcpy(d_coeffs_, &coeffs_, sizeof(IIRCoefficients), cudaMemcpyHostToDevice));
    generateTestData(42);
    iterations_done_ = 0;
    if (ops_ & B200CONV_STRIP_BIQUAD)
        std::printf("IIR coefficients: b0=%.6f, b1=%.6f, b2=%.6f, a1=%.6f, a2=%.6f\n", coeffs_.b0, coeffs_.b1, coeffs_.b2,
                    coeffs_.a1, coeffs_.a2);
    std::printf("%s benchmark setup complete (B200 channel strip:%s%s%s)\n", benchmark_name_.c_str(),
                (ops_ & B200CONV_STRIP_STATS) ? " mean+max" : "", (ops_ & B200CONV_STRIP_GAIN) ? " gain" : "",
                (ops_ & B200CONV_STRIP_BIQUAD) ? " biquad" : "");
}

void ChannelStripBenchmark::performBenchmarkIteration() {
    if (!d_state_) throw std::runtime_error(benchmark_name_ + ": performBenchmarkIteration called before setupBenchmark");
    transferToDevice();
    b200conv_strip strip{};
    strip.ops = ops_ | B200CONV_STRIP_SHARED_COEFFS;
    strip.gain = gain_;
    strip.gains = nullptr;
    strip.biquad = d_coeffs_;
    BenchmarkUtils::CudaEventTimer gpu;
    gpu.start();
    if (b200conv_strip_process(getDeviceInput(), getDeviceOutput(), static_cast<uint32_t>(getTrackCount()),
                               static_cast<uint32_t>(getBufferSize()), B200CONV_OUT_TRACK_MAJOR, 0, 0, &strip, d_state_,
                               d_stats_, 0, nullptr) != B200CONV_OK)
        throw std::runtime_error(std::string("b200conv_strip_process failed: ") + b200conv_last_error());
    recordGpuDuration(gpu.stop());
    synchronizeAndCheck();
    transferToHost();
    const size_t aux = 2 * getTrackCount() * sizeof(float);
    if (ops_ & B200CONV_STRIP_STATS) CUDA_CHECK(cudaMemcpy(h_stats_, d_stats_, aux, cudaMemcpyDeviceToHost));
    if (ops_ & B200CONV_STRIP_BIQUAD) CUDA_CHECK(cudaMemcpy(h_state_, d_state_, aux, cudaMemcpyDeviceToHost));
    ++iterations_done_;
}

void ChannelStripBenchmark::cpuPass(std::vector<float>& out, std::vector<float>& stats, std::vector<float>& state) const {
    const size_t T = getTrackCount(), B = getBufferSize();
    const float* in = buffers.h_input;
    for (size_t t = 0; t < T; ++t) {
        const float* x = in + t * B;
        float* y = out.data() + t * B;
        if (ops_ & B200CONV_STRIP_STATS) {  // running float sum / B and maximum of the input (bench_gainstats.cu:127-141)
            float mean = 0.0f, peak = -1e9f;
            for (size_t i = 0; i < B; ++i) {
                mean += x[i];
                if (x[i] > peak) peak = x[i];
            }
            mean /= B;
            stats[2 * t] = mean;
            stats[2 * t + 1] = peak;
        }
        float z1 = state[2 * t], z2 = state[2 * t + 1];
        for (size_t i = 0; i < B; ++i) {
            float v = x[i];
            if (ops_ & B200CONV_STRIP_GAIN) v = gain_ * v;  // bench_gain.cu:91, bench_gainstats.cu:124
            if (ops_ & B200CONV_STRIP_BIQUAD) {              // Direct Form II, bench_iir.cu:190-197
                const float w = v - coeffs_.a1 * z1 - coeffs_.a2 * z2;
                v = coeffs_.b0 * w + coeffs_.b1 * z1 + coeffs_.b2 * z2;
                z2 = z1;
                z1 = w;
            }
            y[i] = v;
        }
        state[2 * t] = z1;
        state[2 * t + 1] = z2;
    }
}

void ChannelStripBenchmark::validate(ValidationData& validation_data) {
    if (!validation_enabled_) {
        validation_data.status = ValidationStatus::SUCCESS;
        validation_data.messages.push_back("Validation skipped (disabled)");
        return;
    }
    // The filter state lives on the device across iterations (as in the reference, bench_iir.cu:42-43), so
    // the CPU loop is advanced by the same number of passes before the last outputs are compared.  (The
    // reference compares the LAST GPU iteration with the FIRST CPU pass, which can only agree after one
    // iteration.)
    const size_t T = getTrackCount();
    cpu_output_.assign(getTotalElements(), 0.0f);
    cpu_stats_.assign(2 * T, 0.0f);
    cpu_state_.assign(2 * T, 0.0f);
    const int passes = (ops_ & B200CONV_STRIP_BIQUAD) ? std::max(1, iterations_done_) : 1;
    for (int p = 0; p < passes; ++p) cpuPass(cpu_output_, cpu_stats_, cpu_state_);

    validation_data = compareWithReference(cpu_output_.data(), output_tolerance_);
    bit_exact_ = std::memcmp(buffers.h_output, cpu_output_.data(), getTotalElements() * sizeof(float)) == 0;
    float aux_error = 0.0f;
    if (ops_ & B200CONV_STRIP_STATS) {
        for (size_t i = 0; i < 2 * T; ++i) aux_error = std::max(aux_error, std::abs(h_stats_[i] - cpu_stats_[i]));
        bit_exact_ = bit_exact_ && std::memcmp(h_stats_, cpu_stats_.data(), 2 * T * sizeof(float)) == 0;
    }
    if (ops_ & B200CONV_STRIP_BIQUAD) {
        for (size_t i = 0; i < 2 * T; ++i) aux_error = std::max(aux_error, std::abs(h_state_[i] - cpu_state_[i]));
        bit_exact_ = bit_exact_ && std::memcmp(h_state_, cpu_state_.data(), 2 * T * sizeof(float)) == 0;
    }
    const char* aux_name = (ops_ & B200CONV_STRIP_BIQUAD) ? "IIR state" : "Statistics";
    if (aux_error > aux_tolerance_) {
        validation_data.status = ValidationStatus::FAILURE;
        validation_data.messages.push_back(std::string(aux_name) + " validation failed");
        validation_data.max_error = std::max(validation_data.max_error, aux_error);
    } else if (validation_data.status == ValidationStatus::SUCCESS) {
        validation_data.messages.push_back(benchmark_name_ + " validation passed" +
                                           (bit_exact_ ? " (bit-identical to the CPU loop)" : ""));
    }
}

GainBenchmark::GainBenchmark(size_t buffer_size, size_t track_count, bool enable_validation)
    : ChannelStripBenchmark("Gain", buffer_size, track_count, B200CONV_STRIP_GAIN, BenchmarkConstants::GAIN_VALUE, 1e-5f, 0.0f) {
    validation_enabled_ = enable_validation;
}

GainStatsBenchmark::GainStatsBenchmark(size_t buffer_size, size_t track_count)
    : ChannelStripBenchmark("GainStats", buffer_size, track_count, B200CONV_STRIP_GAIN | B200CONV_STRIP_STATS,
                            BenchmarkConstants::GAINSTATS_GAIN, 1e-5f, 1e-4f) {}

IIRBenchmark::IIRBenchmark(size_t buffer_size, size_t track_count)
    : ChannelStripBenchmark("IIRFilter", buffer_size, track_count, B200CONV_STRIP_BIQUAD, 1.0f, 1e-4f, 1e-3f) {
    coeffs_ = calculateButterworthCoefficients(BenchmarkConstants::IIR_NORMALIZED_CUTOFF);
}

IIRCoefficients IIRBenchmark::calculateButterworthCoefficients(float normalized_frequency) {
    const float PI = 3.14159265358979323846f;
    const float omega = 2.0f * PI * normalized_frequency;
    const float cos_omega = cosf(omega), sin_omega = sinf(omega);
    const float alpha = sin_omega / (2.0f * BenchmarkConstants::IIR_BUTTERWORTH_Q);
    const float a0 = 1.0f + alpha;
    IIRCoefficients c;
    c.b0 = ((1.0f - cos_omega) / 2.0f) / a0;
    c.b1 = (1.0f - cos_omega) / a0;
    c.b2 = ((1.0f - cos_omega) / 2.0f) / a0;
    c.a1 = (-2.0f * cos_omega) / a0;
    c.a2 = (1.0f - alpha) / a0;
    return c;
}
