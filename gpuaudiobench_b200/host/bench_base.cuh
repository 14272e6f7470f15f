#pragma once
// GPUABenchmark — the plugin base class of the gpubench CLI.
//
// Same lifecycle and public surface as the reference's cuda/bench_base.cuh:18-139 so that a plugin
// written against the reference compiles against this header: constructor (name, buffer_size,
// track_count); four pure virtuals setupBenchmark / runKernel / performBenchmarkIteration /
// validate; services allocateBuffers, transferToDevice/Host, runBenchmark(iterations, warmup),
// generateTestData(seed), writeResults, printResults; BenchmarkResult, ValidationStatus,
// ValidationData.  Behaviour kept on purpose: synchronous pinned copies on the default stream
// (bench_base.cu:30-42), warm-up exceptions printed and swallowed (:73-79), wall + GPU latency
// per iteration, throughput = T*B*4 bytes / mean latency (:107-113).
#include <functional>
#include <string>
#include <utility>
#include <vector>

#include "bench_utils.cuh"
#include "globals.cuh"

class GPUABenchmark {
public:
    struct BenchmarkResult {
        std::vector<float> latencies;      // wall ms per iteration
        std::vector<float> gpu_latencies;  // device ms per iteration (empty if the plugin records none)
        BenchmarkUtils::Statistics statistics{};
        BenchmarkUtils::Statistics gpu_statistics{};
        std::string benchmark_name;
        size_t buffer_size = 0;
        size_t track_count = 0;
        int iterations = 0;
        double throughput_gbps = 0.0;  // GiB/s of input samples, as the reference computes it
        double samples_per_sec = 0.0;
        size_t bytes_processed = 0;
        float mean_latency_ms = 0.0f;
    };

    enum class ValidationStatus { SUCCESS = 0, FAILURE = 1, FATAL = -1 };

    struct ValidationData {
        ValidationStatus status = ValidationStatus::SUCCESS;
        std::vector<std::string> messages;
        float max_error = 0.0f;
        float mean_error = 0.0f;
    };

    GPUABenchmark(const std::string& name, size_t buffer_size = BUFSIZE, size_t track_count = NTRACKS);
    virtual ~GPUABenchmark();
    GPUABenchmark(const GPUABenchmark&) = delete;
    GPUABenchmark& operator=(const GPUABenchmark&) = delete;

    // ---- what a plugin implements -------------------------------------------------------------
    virtual void setupBenchmark() = 0;
    virtual void runKernel() = 0;
    virtual void performBenchmarkIteration() = 0;
    virtual void validate(ValidationData& validation_data) = 0;

    // ---- what the base provides ---------------------------------------------------------------
    void allocateBuffers(size_t element_count);
    void transferToDevice();
    void transferToHost();
    BenchmarkResult runKernelBenchmark(int iterations = NRUNS, int warmupIterations = 3);
    BenchmarkResult runBenchmark(int iterations = NRUNS, int warmupIterations = 3);
    void generateTestData(unsigned int seed = 42);
    void writeResults(const BenchmarkResult& result, const std::string& filename = "");
    void printResults(const BenchmarkResult& result);

    const std::string& getName() const { return benchmark_name_; }
    size_t getBufferSize() const { return buffer_size_; }
    size_t getTrackCount() const { return track_count_; }
    size_t getTotalElements() const { return buffer_size_ * track_count_; }
    // read-only view of the last iteration's output (used by the C binding and the tests)
    const float* hostOutput() const { return buffers.h_output; }
    const float* hostInput() const { return buffers.h_input; }

protected:
    struct BufferSet {  // pinned host + device in/out, released together
        float* h_input = nullptr;
        float* h_output = nullptr;
        float* d_input = nullptr;
        float* d_output = nullptr;
        size_t element_count = 0;
        size_t size_bytes = 0;
        void cleanup();
        ~BufferSet() { cleanup(); }
    };

    BufferSet buffers;
    BenchmarkUtils::BenchmarkTimer timer;
    std::string benchmark_name_;
    size_t buffer_size_;
    size_t track_count_;
    float current_iteration_gpu_ms_ = 0.0f;

    float* getHostInput() { return buffers.h_input; }
    float* getHostOutput() { return buffers.h_output; }
    float* getDeviceInput() { return buffers.d_input; }
    float* getDeviceOutput() { return buffers.d_output; }
    BenchmarkUtils::BenchmarkParams makeBenchmarkParams(float gainValue = 0.0f) const;
    std::pair<int, int> calculateGridDimensions(int desired_threads_per_block = 256) const;
    void synchronizeAndCheck();
    ValidationData compareWithReference(const float* cpu_reference, float tolerance = 1e-5f);
    void resetGpuIterationMetrics();
    void recordGpuDuration(float milliseconds);
    BenchmarkResult runWithIteration(int iterations, int warmupIterations, const std::function<void()>& iterationBody);
};
