#pragma once
// Conv1D_accel plugin: frequency-domain per-track FIR on the B200 engine (b200conv, UPOLS).
// Drop-in for the reference's Conv1DAccelBenchmark (cuda/bench_conv1d_accel.cuh:8-79): same
// constructor and overrides, registry name "Conv1D_accel", input track-major [T][B], output
// SAMPLE-major [B][T] (bench_conv1d_accel.cu:44,249).  cuFFT plans, the three T*N complex buffers
// and the per-track memcpy loops (bench_conv1d_accel.cu:88-150,258-304) are replaced by the engine.
#include "bench_base.cuh"
#include "conv_common.cuh"

class Conv1DAccelBenchmark : public GPUABenchmark {
public:
    static constexpr int DEFAULT_IR_LEN = 512;

    Conv1DAccelBenchmark(int ir_length = DEFAULT_IR_LEN, size_t buffer_size = BUFSIZE, size_t track_count = NTRACKS);
    ~Conv1DAccelBenchmark() override;

    void setupBenchmark() override;
    void runKernel() override;
    void performBenchmarkIteration() override;
    void validate(ValidationData& validation_data) override;

    int getIRLength() const { return ir_length_; }
    int getFFTSize() const { return fft_size_; }
    int getPartitionCount() const { return partitions_; }
    const float* hostIR() const { return h_ir_buf; }
    const float* cpuReference() const { return cpu_reference; }
    b200conv_info engineInfo() { return engine_.info(); }

private:
    void generateImpulseResponses();
    void calculateCPUReference();

    int ir_length_;
    int fft_size_;    // N = 2B (the reference used nextpow2(L+B-1) per buffer, bench_conv1d_accel.cu:52)
    int partitions_;  // P = ceil(L/B)
    float* h_ir_buf = nullptr;
    float* cpu_reference = nullptr;
    ConvCommon::Engine engine_;
    ConvCommon::Group group_;  // used instead of engine_ when NGPUS > 1
    bool ready_ = false;
};
