#include "bench_base.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <iomanip>
#include <iostream>
#include <memory>
#include <stdexcept>

GPUABenchmark::GPUABenchmark(const std::string& name, size_t buffer_size, size_t track_count)
    : benchmark_name_(name), buffer_size_(buffer_size), track_count_(track_count) {}

GPUABenchmark::~GPUABenchmark() = default;

void GPUABenchmark::BufferSet::cleanup() {
    if (!h_input && !h_output && !d_input && !d_output) return;
    const cudaError_t sync = cudaDeviceSynchronize();  // nothing may still be reading the buffers
    if (sync != cudaSuccess)
        std::fprintf(stderr, "Warning: cudaDeviceSynchronize before cleanup failed: %s\n", cudaGetErrorString(sync));
    BenchmarkUtils::freeHostBuffers({h_input, h_output});
    BenchmarkUtils::freeDeviceBuffers({d_input, d_output});
    h_input = h_output = d_input = d_output = nullptr;
}

void GPUABenchmark::allocateBuffers(size_t element_count) {
    if (element_count == 0) throw std::invalid_argument("allocateBuffers requires element_count > 0");
    buffers.element_count = element_count;
    buffers.size_bytes = element_count * sizeof(float);
    buffers.h_input = BenchmarkUtils::allocateHostBuffer<float>(element_count, benchmark_name_ + " host input buffer");
    buffers.h_output = BenchmarkUtils::allocateHostBuffer<float>(element_count, benchmark_name_ + " host output buffer");
    buffers.d_input = BenchmarkUtils::allocateDeviceBuffer<float>(element_count, benchmark_name_ + " device input buffer");
    buffers.d_output = BenchmarkUtils::allocateDeviceBuffer<float>(element_count, benchmark_name_ + " device output buffer");
}

void GPUABenchmark::transferToDevice() {
    if (!buffers.d_input || !buffers.h_input)
        throw std::runtime_error("transferToDevice called before input buffers were allocated");
    BenchmarkUtils::copyToDevice(buffers.d_input, buffers.h_input, buffers.element_count);
}

void GPUABenchmark::transferToHost() {
    if (!buffers.d_output || !buffers.h_output)
        throw std::runtime_error("transferToHost called before output buffers were allocated");
    BenchmarkUtils::copyToHost(buffers.h_output, buffers.d_output, buffers.element_count);
}

void GPUABenchmark::generateTestData(unsigned int seed) {
    if (!buffers.h_input) throw std::runtime_error("generateTestData called before host input buffer allocation");
    BenchmarkUtils::generateRandomAudioData(buffers.h_input, buffers.element_count, seed);
}

GPUABenchmark::BenchmarkResult GPUABenchmark::runKernelBenchmark(int iterations, int warmupIterations) {
    return runWithIteration(iterations, warmupIterations, [this] { runKernel(); });
}

GPUABenchmark::BenchmarkResult GPUABenchmark::runBenchmark(int iterations, int warmupIterations) {
    return runWithIteration(iterations, warmupIterations, [this] { performBenchmarkIteration(); });
}

GPUABenchmark::BenchmarkResult GPUABenchmark::runWithIteration(int iterations, int warmupIterations,
                                                               const std::function<void()>& iterationBody) {
    BenchmarkResult res;
    res.benchmark_name = benchmark_name_;
    res.buffer_size = buffer_size_;
    res.track_count = track_count_;
    res.iterations = iterations;

    // DAW-style pacing: after every iteration (warm-up included) wait for the next buffer period,
    // as the Metal port does (GPUABenchmark.swift:371-390)
    std::unique_ptr<BenchmarkUtils::DAWSimulator> daw;
    if (DAWSIM)
        daw = std::make_unique<BenchmarkUtils::DAWSimulator>(
            static_cast<double>(buffer_size_) / FS,
            DAWSIM_SLEEP ? BenchmarkUtils::DAWSimulator::Mode::SLEEP : BenchmarkUtils::DAWSimulator::Mode::SPIN,
            DAWSIM_JITTER_US * 1e-6);

    if (warmupIterations > 0) {
        std::printf("Running %d warmup iterations...\n", warmupIterations);
        for (int w = 1; w <= warmupIterations; ++w) {
            try {
                resetGpuIterationMetrics();
                iterationBody();
                std::printf("  Warmup %d/%d completed\n", w, warmupIterations);
            } catch (const std::exception& ex) {  // the reference reports and carries on
                std::printf("  Warmup iteration %d failed: %s\n", w, ex.what());
            }
            if (daw) daw->wait();
        }
        std::printf("Warmup complete, starting timed iterations...\n");
    }

    res.latencies.reserve(iterations);
    std::vector<float> device_ms;
    device_ms.reserve(iterations);
    for (int i = 0; i < iterations; ++i) {
        resetGpuIterationMetrics();
        res.latencies.push_back(static_cast<float>(BenchmarkUtils::BenchmarkTimer::measureKernel(iterationBody)));
        device_ms.push_back(current_iteration_gpu_ms_);
        if (daw) daw->wait();
    }
    res.statistics = BenchmarkUtils::calculateStatistics(res.latencies);
    if (std::any_of(device_ms.begin(), device_ms.end(), [](float v) { return v > 0.0f; })) {
        res.gpu_latencies = std::move(device_ms);
        res.gpu_statistics = BenchmarkUtils::calculateStatistics(res.gpu_latencies);
    }

    const size_t samples = buffer_size_ * track_count_;
    res.bytes_processed = samples * sizeof(float);
    res.mean_latency_ms = res.statistics.mean;
    const double mean_s = res.mean_latency_ms / 1000.0;
    res.throughput_gbps = (res.bytes_processed / (1024.0 * 1024.0 * 1024.0)) / mean_s;
    res.samples_per_sec = samples / mean_s;
    return res;
}

void GPUABenchmark::writeResults(const BenchmarkResult& result, const std::string& filename) {
    const std::string path = filename.empty() ? "/tmp/" + result.benchmark_name + "_latencies.txt" : filename;
    BenchmarkUtils::writeLatenciesToFile(result.latencies, path);
}

void GPUABenchmark::printResults(const BenchmarkResult& result) {
    BenchmarkUtils::printStatistics(result.latencies, result.benchmark_name);
    if (!result.gpu_latencies.empty()) {
        const BenchmarkUtils::Statistics g = result.gpu_statistics.count ? result.gpu_statistics
                                                                         : BenchmarkUtils::calculateStatistics(result.gpu_latencies);
        std::cout << std::fixed << std::setprecision(3) << "GPU Median:  " << g.median << " ms\n"
                  << "GPU P95:     " << g.p95 << " ms\n"
                  << "GPU Mean:    " << g.mean << " ms" << std::endl;
    }
    std::cout << "\nPerformance Metrics:\n"
              << std::fixed << std::setprecision(3) << "Throughput:        " << result.throughput_gbps << " GB/s\n"
              << "Samples/sec:       " << std::setprecision(0) << result.samples_per_sec << "\n"
              << "Bytes processed:   " << result.bytes_processed << std::endl;
}

BenchmarkUtils::BenchmarkParams GPUABenchmark::makeBenchmarkParams(float gainValue) const {
    return BenchmarkUtils::makeBenchmarkParams(buffer_size_, track_count_, gainValue);
}

void GPUABenchmark::resetGpuIterationMetrics() { current_iteration_gpu_ms_ = 0.0f; }

void GPUABenchmark::recordGpuDuration(float milliseconds) {
    if (milliseconds > 0.0f) current_iteration_gpu_ms_ += milliseconds;
}

std::pair<int, int> GPUABenchmark::calculateGridDimensions(int desired_threads_per_block) const {
    const int threads = std::max(32, std::min(desired_threads_per_block, 512));
    const int blocks = (static_cast<int>(track_count_) + threads - 1) / threads;
    return {blocks, threads};
}

void GPUABenchmark::synchronizeAndCheck() { CUDA_CHECK(cudaDeviceSynchronize()); }

GPUABenchmark::ValidationData GPUABenchmark::compareWithReference(const float* cpu_reference, float tolerance) {
    ValidationData v;
    if (!buffers.h_output || !cpu_reference) {
        v.status = ValidationStatus::FATAL;
        v.messages.push_back("Null pointer in validation comparison");
        return v;
    }
    float total = 0.0f, worst = 0.0f;
    int over = 0;
    for (size_t i = 0; i < buffers.element_count; ++i) {
        const float diff = std::abs(buffers.h_output[i] - cpu_reference[i]);
        total += diff;
        worst = std::max(worst, diff);
        if (diff > tolerance) {
            ++over;
            if (v.messages.size() < 10)
                v.messages.push_back("Error at index " + std::to_string(i) + ": expected " + std::to_string(cpu_reference[i]) +
                                     ", got " + std::to_string(buffers.h_output[i]) + ", diff " + std::to_string(diff));
        }
    }
    v.mean_error = total / static_cast<float>(buffers.element_count);
    v.max_error = worst;
    if (over > 0) {
        v.status = ValidationStatus::FAILURE;
        v.messages.insert(v.messages.begin(), "Validation failed: " + std::to_string(over) + " out of " +
                                                  std::to_string(buffers.element_count) + " elements exceeded tolerance");
    }
    return v;
}
