#pragma once
// Host utilities of the benchmark plugins: buffers, timers, the input generator, statistics.
// API-compatible with the subset of the reference's cuda/bench_utils.cuh that the convolution
// path uses; the kernels themselves live behind include/b200conv.h,
// so there are no launchKernel* templates here — `timeStream` is the CudaEventTimer use-case.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdint>
#include <functional>
#include <initializer_list>
#include <string>
#include <vector>

namespace BenchmarkUtils {

// POD the reference passes by value to its kernels (bench_utils.cuh:22-27); kept because the
// plugin API exposes makeBenchmarkParams().
struct BenchmarkParams {
    uint32_t bufferSize = 0;
    uint32_t trackCount = 0;
    uint32_t totalSamples = 0;
    float gainValue = 0.0f;
};
BenchmarkParams makeBenchmarkParams(size_t bufferSize, size_t trackCount, float gainValue = 0.0f);

// ---- memory (throwing, like bench_utils.cu:99-171) ------------------------------------------
template <typename T> T* allocateDeviceBuffer(size_t count, const std::string& name = "device buffer");
template <typename T> T* allocateHostBuffer(size_t count, const std::string& name = "host buffer");  // pinned
template <typename T> void copyToDevice(T* dst, const T* src, size_t count);
template <typename T> void copyToHost(T* dst, const T* src, size_t count);
void freeDeviceBuffers(std::initializer_list<void*> buffers);
void freeHostBuffers(std::initializer_list<void*> buffers);

// ---- timing ------------------------------------------------------------------------------------
class BenchmarkTimer {  // wall clock, microsecond resolution reported in ms (bench_utils.cu:187-216)
public:
    void start();
    void stop();
    double elapsed_ms() const;
    void reset();
    static double measureKernel(std::function<void()> body);

private:
    std::chrono::steady_clock::time_point begin_{}, end_{};
    bool running_ = false;
};

class CudaEventTimer {  // device time between two events on a stream (bench_utils.cu:28-95)
public:
    CudaEventTimer();
    ~CudaEventTimer();
    CudaEventTimer(const CudaEventTimer&) = delete;
    CudaEventTimer& operator=(const CudaEventTimer&) = delete;
    void start(cudaStream_t stream = 0);
    float stop(cudaStream_t stream = 0);  // records, synchronises, returns ms (0 if not started)
    bool isRunning() const { return running_; }

private:
    cudaEvent_t first_ = nullptr, second_ = nullptr;
    bool running_ = false;
};

void collectLatencies(std::vector<float>& latencies, std::function<void()> benchmark, int iterations);

// Paces iterations like an audio callback: wait() returns at the next multiple of the buffer period
// (optionally jittered), spinning or sleeping.  Semantics of the reference's Metal DAWSimulator
// (BenchmarkUtilities.swift:151-178): the first call arms nextStart = now + period; every call waits
// until nextStart (+ uniform jitter in [-j, +j]) if that is still in the future, then advances
// nextStart by one period — an iteration that overruns its period is not waited for.
class DAWSimulator {
public:
    enum class Mode { SPIN, SLEEP };
    DAWSimulator(double buffer_duration_s, Mode mode, double jitter_s, unsigned seed = 1);
    void wait();
    double bufferDuration() const { return period_; }

private:
    double period_, jitter_;
    Mode mode_;
    bool armed_ = false;
    std::chrono::steady_clock::time_point next_start_{};
    uint64_t rng_;
};

// ---- data ---------------------------------------------------------------------------------------
// std::mt19937(seed) + uniform_real_distribution<float>(-1, 1), one sequential draw (bench_utils.cu:238-245)
void generateRandomAudioData(float* buffer, size_t samples, unsigned int seed = 42);

// ---- errors -------------------------------------------------------------------------------------
void checkCudaError(cudaError_t error, const std::string& message);
#define CUDA_CHECK(call)                                                           \
    do {                                                                           \
        cudaError_t cuda_check_status_ = (call);                                   \
        if (cuda_check_status_ != cudaSuccess) BenchmarkUtils::checkCudaError(cuda_check_status_, #call); \
    } while (0)

// ---- statistics (bench_utils.cu:358-458) ---------------------------------------------------------
struct Statistics {
    float mean, median, std_dev, min_val, max_val, p95, p99;
    size_t count;
};
Statistics calculateStatistics(const std::vector<float>& latencies);  // linear-interpolated p95/p99
void writeLatenciesToFile(const std::vector<float>& latencies, const std::string& filename);
void printStatistics(const std::vector<float>& latencies, const std::string& benchmark_name);

}  // namespace BenchmarkUtils
