#pragma once
// FFT1D plugin: batched 1024-point real-to-complex FFT on the engine's shared-memory Stockham
// transform (b200conv_rfft) instead of cuFFT.  Drop-in for the reference's FFTBenchmark
// (cuda/bench_fft.cuh:9-66): same constructor, overrides and registry name "FFT1D"; input
// [T][1024] real (buffers shorter than 1024 are zero padded), output [T][513] complex.
// the first widening step beyond the convolution path; it also gives the FFT
// stage of the UPOLS engine an oracle of its own.
#include <cuda_runtime.h>

#include <vector>

#include "bench_base.cuh"

class FFTBenchmark : public GPUABenchmark {
public:
    static constexpr int FFT_SIZE = 1024;

    FFTBenchmark(size_t buffer_size = BUFSIZE, size_t track_count = NTRACKS);
    ~FFTBenchmark() override;

    void setupBenchmark() override;
    void runKernel() override;
    void performBenchmarkIteration() override;
    void validate(ValidationData& validation_data) override;

    size_t binsPerTrack() const { return FFT_SIZE / 2 + 1; }
    const float* hostInputFFT() const { return h_input_fft; }
    const float2* hostOutputFFT() const { return h_output_fft; }
    const float2* cpuReferenceFFT() const { return cpu_reference.data(); }
    double lastSnrDb() const { return last_snr_db_; }

private:
    void allocateFFTBuffers();
    void calculateCPUReference();
    // naive float DFT exactly as the reference computes it (bench_fft.cu:149-168)
    static void cpuFFTReference(const float* input, float2* output, int size);
    // the same DFT in double with exact angle reduction: the truth the stated tolerance refers to
    static void cpuFFTTruth(const float* input, double* re, double* im, int size);

    float* h_input_fft = nullptr;    // pinned [T][FFT_SIZE]
    float2* h_output_fft = nullptr;  // pinned [T][FFT_SIZE/2+1]
    float* d_input_fft = nullptr;
    float2* d_output_fft = nullptr;
    std::vector<float2> cpu_reference;
    std::vector<double> truth_re, truth_im;
    size_t input_fft_size, output_fft_size;
    double last_snr_db_ = 0.0;
};
