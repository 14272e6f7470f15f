#pragma once
// Conv1D plugin: direct-form per-track FIR on the B200 engine (b200conv, DIRECT algorithm).
// Drop-in for the reference's Conv1DBenchmark (cuda/bench_conv1d.cuh:8-66): same constructor,
// same four overrides, same registry name "Conv1D", input and output both track-major [T][B].
// The texture-memory kernel and its cudaArray (bench_conv1d.cu:7-27,123-157) are replaced by the
// engine; the CPU reference and the abs-1e-3 check are kept, with the stated SNR tolerance added.
#include <memory>
#include <vector>

#include "bench_base.cuh"
#include "conv_common.cuh"

class Conv1DBenchmark : public GPUABenchmark {
public:
    static constexpr int DEFAULT_IR_LEN = 1024;

    Conv1DBenchmark(int ir_length = DEFAULT_IR_LEN, size_t buffer_size = BUFSIZE, size_t track_count = NTRACKS);
    ~Conv1DBenchmark() override;

    void setupBenchmark() override;
    void runKernel() override;
    void performBenchmarkIteration() override;
    void validate(ValidationData& validation_data) override;

    int getIRLength() const { return ir_length_; }
    const float* hostIR() const { return h_ir_buf; }
    const float* cpuReference() const { return cpu_reference; }
    b200conv_info engineInfo() { return engine_.info(); }

private:
    void allocateConvBuffers();
    void generateImpulseResponses();
    void primeReferenceHistory();
    void calculateCPUReference();
    void oneIteration(const char* caller);

    int ir_length_;
    float* h_ir_buf = nullptr;       // pinned, [T][L]
    float* cpu_reference = nullptr;  // pinned, [T][B]
    ConvCommon::Engine engine_;
    ConvCommon::Group group_;  // used instead of engine_ when NGPUS > 1
    std::vector<float> history_;  // R1-compatible priming history [T][L-1]
    bool ready_ = false;
};
