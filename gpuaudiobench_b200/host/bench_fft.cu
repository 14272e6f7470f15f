#include "bench_fft.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

#include "b200conv.h"

FFTBenchmark::FFTBenchmark(size_t buffer_size, size_t track_count)
    : GPUABenchmark("FFT1D", buffer_size, track_count),
      input_fft_size(track_count * FFT_SIZE),
      output_fft_size(track_count * (FFT_SIZE / 2 + 1)) {}

FFTBenchmark::~FFTBenchmark() {
    BenchmarkUtils::freeHostBuffers({h_input_fft, h_output_fft});
    BenchmarkUtils::freeDeviceBuffers({d_input_fft, d_output_fft});
}

void FFTBenchmark::allocateFFTBuffers() {
    h_input_fft = BenchmarkUtils::allocateHostBuffer<float>(input_fft_size, benchmark_name_ + " host input FFT buffer");
    h_output_fft = reinterpret_cast<float2*>(
        BenchmarkUtils::allocateHostBuffer<float>(2 * output_fft_size, benchmark_name_ + " host output FFT buffer"));
    d_input_fft = BenchmarkUtils::allocateDeviceBuffer<float>(input_fft_size, benchmark_name_ + " device input FFT buffer");
    d_output_fft = reinterpret_cast<float2*>(
        BenchmarkUtils::allocateDeviceBuffer<float>(2 * output_fft_size, benchmark_name_ + " device output FFT buffer"));
    std::fill(h_output_fft, h_output_fft + output_fft_size, make_float2(0.0f, 0.0f));
    CUDA_CHECK(cudaMemset(d_output_fft, 0, output_fft_size * sizeof(float2)));
}

void FFTBenchmark::setupBenchmark() {
    allocateFFTBuffers();
    // input as the reference draws it (bench_fft.cu:37-46): C rand() mapped to [-1, 1], zero padded
    const size_t live = std::min(getBufferSize(), static_cast<size_t>(FFT_SIZE));
    for (size_t t = 0; t < getTrackCount(); ++t) {
        float* row = h_input_fft + t * FFT_SIZE;
        for (size_t i = 0; i < live; ++i) row[i] = ((float)rand() / (float)RAND_MAX) * 2.0f - 1.0f;
        std::fill(row + live, row + FFT_SIZE, 0.0f);
    }
    calculateCPUReference();
    std::printf("FFT benchmark setup complete (FFT size = %d, %zu tracks, B200 Stockham R2C)\n", FFT_SIZE, getTrackCount());
}

void FFTBenchmark::runKernel() { performBenchmarkIteration(); }

void FFTBenchmark::performBenchmarkIteration() {
    if (!d_input_fft) throw std::runtime_error("FFTBenchmark::performBenchmarkIteration called before setupBenchmark");
    CUDA_CHECK(cudaMemcpy(d_input_fft, h_input_fft, input_fft_size * sizeof(float), cudaMemcpyHostToDevice));
    BenchmarkUtils::CudaEventTimer gpu;
    gpu.start();
    if (b200conv_rfft(d_input_fft, d_output_fft, static_cast<int>(getTrackCount()), FFT_SIZE, nullptr) != B200CONV_OK)
        throw std::runtime_error(std::string("b200conv_rfft failed: ") + b200conv_last_error());
    recordGpuDuration(gpu.stop());
    synchronizeAndCheck();
    CUDA_CHECK(cudaMemcpy(h_output_fft, d_output_fft, output_fft_size * sizeof(float2), cudaMemcpyDeviceToHost));
}

void FFTBenchmark::cpuFFTReference(const float* input, float2* output, int size) {
    const float PI = 3.14159265358979323846f;
    for (int k = 0; k <= size / 2; ++k) {
        float re = 0.0f, im = 0.0f;
        for (int n = 0; n < size; ++n) {
            const float angle = -2.0f * PI * k * n / size;  // float angle, as the reference
            re += input[n] * cosf(angle);
            im += input[n] * sinf(angle);
        }
        output[k] = make_float2(re, im);
    }
}

void FFTBenchmark::cpuFFTTruth(const float* input, double* re, double* im, int size) {
    std::vector<double> c(size), s(size);
    for (int q = 0; q < size; ++q) {
        const double a = -2.0 * 3.14159265358979323846 * q / size;
        c[q] = std::cos(a);
        s[q] = std::sin(a);
    }
    for (int k = 0; k <= size / 2; ++k) {
        double sr = 0.0, si = 0.0;
        for (int n = 0; n < size; ++n) {
            const int q = (k * n) % size;  // exact angle reduction
            sr += input[n] * c[q];
            si += input[n] * s[q];
        }
        re[k] = sr;
        im[k] = si;
    }
}

void FFTBenchmark::calculateCPUReference() {
    const size_t bins = binsPerTrack();
    cpu_reference.resize(output_fft_size);
    truth_re.resize(output_fft_size);
    truth_im.resize(output_fft_size);
    for (size_t t = 0; t < getTrackCount(); ++t) {
        cpuFFTReference(h_input_fft + t * FFT_SIZE, cpu_reference.data() + t * bins, FFT_SIZE);
        cpuFFTTruth(h_input_fft + t * FFT_SIZE, truth_re.data() + t * bins, truth_im.data() + t * bins, FFT_SIZE);
    }
}

void FFTBenchmark::validate(ValidationData& validation_data) {
    // the reference's metric (bench_fft.cu:74-97): |d re| + |d im| against the float DFT, tolerance 1e-3
    float max_error = 0.0f, mean_error = 0.0f;
    double sig = 0.0, noise = 0.0, oracle_noise = 0.0;
    for (size_t i = 0; i < output_fft_size; ++i) {
        const float total = std::abs(h_output_fft[i].x - cpu_reference[i].x) + std::abs(h_output_fft[i].y - cpu_reference[i].y);
        max_error = std::max(max_error, total);
        mean_error += total;
        const double dr = h_output_fft[i].x - truth_re[i], di = h_output_fft[i].y - truth_im[i];
        const double orr = cpu_reference[i].x - truth_re[i], oi = cpu_reference[i].y - truth_im[i];
        sig += truth_re[i] * truth_re[i] + truth_im[i] * truth_im[i];
        noise += dr * dr + di * di;
        oracle_noise += orr * orr + oi * oi;
    }
    mean_error /= (2.0f * output_fft_size);
    validation_data.max_error = max_error;
    validation_data.mean_error = mean_error;
    last_snr_db_ = noise > 0 ? 10.0 * std::log10(sig / noise) : 300.0;
    const double oracle_snr = oracle_noise > 0 ? 10.0 * std::log10(sig / oracle_noise) : 300.0;
    // stated tolerance: SNR >= 110 dB against the fp64 DFT.  The reference's float-angle DFT is itself
    // only ~60-70 dB accurate (angle -2*pi*k*n/N evaluated in float), so its 1e-3 check measures the
    // oracle's own error, not the transform's; it is reported, not enforced.
    char buf[256];
    std::snprintf(buf, sizeof(buf), "SNR vs fp64 DFT %.1f dB (need >= 110); reference float DFT oracle itself %.1f dB; "
                  "reference metric max(|dre|+|dim|) %.3e (reference tolerance 1e-3 %s)", last_snr_db_, oracle_snr,
                  max_error, max_error > 1e-3f ? "not met" : "met");
    validation_data.messages.push_back(buf);
    if (last_snr_db_ >= 110.0) {
        validation_data.status = ValidationStatus::SUCCESS;
        validation_data.messages.push_back("FFT validation passed");
    } else {
        validation_data.status = ValidationStatus::FAILURE;
        validation_data.messages.push_back("FFT validation failed");
    }
}
