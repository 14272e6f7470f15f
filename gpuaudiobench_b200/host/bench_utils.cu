#include "bench_utils.cuh"

#include <algorithm>
#include <cmath>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <numeric>
#include <random>
#include <stdexcept>
#include <thread>

namespace BenchmarkUtils {

BenchmarkParams makeBenchmarkParams(size_t bufferSize, size_t trackCount, float gainValue) {
    BenchmarkParams p;
    p.bufferSize = static_cast<uint32_t>(bufferSize);
    p.trackCount = static_cast<uint32_t>(trackCount);
    p.totalSamples = static_cast<uint32_t>(bufferSize * trackCount);
    p.gainValue = gainValue;
    return p;
}

// ---- memory -------------------------------------------------------------------------------------
namespace {
std::string describe(const char* what, const std::string& name, size_t bytes, cudaError_t err) {
    return std::string("Failed to allocate ") + what + name + " (" + std::to_string(bytes) + " bytes): " +
           cudaGetErrorString(err);
}
}  // namespace

template <typename T> T* allocateDeviceBuffer(size_t count, const std::string& name) {
    void* p = nullptr;
    const size_t bytes = count * sizeof(T);
    const cudaError_t err = cudaMalloc(&p, bytes);
    if (err != cudaSuccess) throw std::runtime_error(describe("", name, bytes, err));
    return static_cast<T*>(p);
}

template <typename T> T* allocateHostBuffer(size_t count, const std::string& name) {
    void* p = nullptr;
    const size_t bytes = count * sizeof(T);
    const cudaError_t err = cudaMallocHost(&p, bytes);
    if (err != cudaSuccess) throw std::runtime_error(describe("pinned ", name, bytes, err));
    return static_cast<T*>(p);
}

template <typename T> void copyToDevice(T* dst, const T* src, size_t count) {
    if (!dst || !src) throw std::invalid_argument("copyToDevice received null pointer");
    const size_t bytes = count * sizeof(T);
    const cudaError_t err = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
    if (err != cudaSuccess)
        throw std::runtime_error("Failed to copy " + std::to_string(bytes) + " bytes to device: " + cudaGetErrorString(err));
}

template <typename T> void copyToHost(T* dst, const T* src, size_t count) {
    if (!dst || !src) throw std::invalid_argument("copyToHost received null pointer");
    const size_t bytes = count * sizeof(T);
    const cudaError_t err = cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost);
    if (err != cudaSuccess)
        throw std::runtime_error("Failed to copy " + std::to_string(bytes) + " bytes to host: " + cudaGetErrorString(err));
}

void freeDeviceBuffers(std::initializer_list<void*> buffers) {
    for (void* b : buffers)
        if (b) cudaFree(b);
}

void freeHostBuffers(std::initializer_list<void*> buffers) {
    for (void* b : buffers)
        if (b) cudaFreeHost(b);
}

template float* allocateDeviceBuffer<float>(size_t, const std::string&);
template int* allocateDeviceBuffer<int>(size_t, const std::string&);
template float* allocateHostBuffer<float>(size_t, const std::string&);
template int* allocateHostBuffer<int>(size_t, const std::string&);
template void copyToDevice<float>(float*, const float*, size_t);
template void copyToHost<float>(float*, const float*, size_t);

// ---- timing -------------------------------------------------------------------------------------
void BenchmarkTimer::start() {
    begin_ = std::chrono::steady_clock::now();
    running_ = true;
}
void BenchmarkTimer::stop() {
    end_ = std::chrono::steady_clock::now();
    running_ = false;
}
double BenchmarkTimer::elapsed_ms() const {
    const auto until = running_ ? std::chrono::steady_clock::now() : end_;
    return std::chrono::duration_cast<std::chrono::microseconds>(until - begin_).count() / 1000.0;
}
void BenchmarkTimer::reset() { running_ = false; }
double BenchmarkTimer::measureKernel(std::function<void()> body) {
    BenchmarkTimer t;
    t.start();
    body();
    t.stop();
    return t.elapsed_ms();
}

CudaEventTimer::CudaEventTimer() {
    CUDA_CHECK(cudaEventCreate(&first_));
    CUDA_CHECK(cudaEventCreate(&second_));
}
CudaEventTimer::~CudaEventTimer() {
    if (first_) cudaEventDestroy(first_);
    if (second_) cudaEventDestroy(second_);
}
void CudaEventTimer::start(cudaStream_t stream) {
    CUDA_CHECK(cudaEventRecord(first_, stream));
    running_ = true;
}
float CudaEventTimer::stop(cudaStream_t stream) {
    if (!running_) return 0.0f;
    CUDA_CHECK(cudaEventRecord(second_, stream));
    CUDA_CHECK(cudaEventSynchronize(second_));
    float ms = 0.0f;
    CUDA_CHECK(cudaEventElapsedTime(&ms, first_, second_));
    running_ = false;
    return ms;
}

void collectLatencies(std::vector<float>& latencies, std::function<void()> benchmark, int iterations) {
    latencies.assign(0, 0.0f);
    latencies.reserve(iterations);
    for (int i = 0; i < iterations; ++i) latencies.push_back(static_cast<float>(BenchmarkTimer::measureKernel(benchmark)));
}

DAWSimulator::DAWSimulator(double buffer_duration_s, Mode mode, double jitter_s, unsigned seed)
    : period_(buffer_duration_s), jitter_(jitter_s), mode_(mode), rng_(0x9E3779B97F4A7C15ull ^ seed) {}

void DAWSimulator::wait() {
    using clock = std::chrono::steady_clock;
    const auto period = std::chrono::duration_cast<clock::duration>(std::chrono::duration<double>(period_));
    const auto now = clock::now();
    if (!armed_) {
        next_start_ = now + period;
        armed_ = true;
    }
    double jitter = 0.0;
    if (jitter_ > 0.0) {  // xorshift64*: uniform in [-jitter, +jitter]
        rng_ ^= rng_ >> 12;
        rng_ ^= rng_ << 25;
        rng_ ^= rng_ >> 27;
        const double u = static_cast<double>((rng_ * 0x2545F4914F6CDD1Dull) >> 11) / 9007199254740992.0;
        jitter = (2.0 * u - 1.0) * jitter_;
    }
    const auto target = next_start_ + std::chrono::duration_cast<clock::duration>(std::chrono::duration<double>(jitter));
    if (target > now) {
        if (mode_ == Mode::SLEEP)
            std::this_thread::sleep_until(target);
        else
            while (clock::now() < target) {
            }
    }
    next_start_ += period;
}

// ---- data ---------------------------------------------------------------------------------------
void generateRandomAudioData(float* buffer, size_t samples, unsigned int seed) {
    std::mt19937 engine(seed);
    std::uniform_real_distribution<float> uniform(-1.0f, 1.0f);
    for (size_t n = 0; n < samples; ++n) buffer[n] = uniform(engine);
}

void checkCudaError(cudaError_t error, const std::string& message) {
    if (error != cudaSuccess) throw std::runtime_error(message + ": " + cudaGetErrorString(error));
}

// ---- statistics ---------------------------------------------------------------------------------
Statistics calculateStatistics(const std::vector<float>& lat) {
    Statistics s{0, 0, 0, 0, 0, 0, 0, 0};
    const size_t n = lat.size();
    if (n == 0) return s;
    s.count = n;
    std::vector<float> sorted(lat);
    std::sort(sorted.begin(), sorted.end());
    s.min_val = sorted.front();
    s.max_val = sorted.back();
    s.mean = std::accumulate(lat.begin(), lat.end(), 0.0f) / static_cast<float>(n);
    s.median = (n % 2) ? sorted[n / 2] : (sorted[n / 2 - 1] + sorted[n / 2]) / 2.0f;
    float ss = 0.0f;
    for (float v : lat) ss += (v - s.mean) * (v - s.mean);
    s.std_dev = std::sqrt(ss / static_cast<float>(n - 1));
    auto interpolated = [&](float percent) {
        const float pos = percent / 100.0f * static_cast<float>(n - 1);
        const size_t lo = static_cast<size_t>(std::floor(pos)), hi = static_cast<size_t>(std::ceil(pos));
        if (lo == hi) return sorted[lo];
        const float w = pos - static_cast<float>(lo);
        return sorted[lo] * (1.0f - w) + sorted[hi] * w;
    };
    s.p95 = interpolated(95.0f);
    s.p99 = interpolated(99.0f);
    return s;
}

void writeLatenciesToFile(const std::vector<float>& latencies, const std::string& filename) {
    const Statistics s = calculateStatistics(latencies);
    std::ofstream out(filename);
    if (!out.is_open()) throw std::runtime_error("Failed to open file for writing: " + filename);
    out << std::fixed << std::setprecision(3) << "# Latency Statistics (ms)\n"
        << "# Count: " << s.count << "\n# Mean: " << s.mean << "\n# Median: " << s.median << "\n# Std Dev: " << s.std_dev
        << "\n# Min: " << s.min_val << "\n# Max: " << s.max_val << "\n# P95: " << s.p95 << "\n# P99: " << s.p99
        << "\n#\n# Raw latencies:\n";
    for (float v : latencies) out << v << "\n";
}

void printStatistics(const std::vector<float>& latencies, const std::string& benchmark_name) {
    const Statistics s = calculateStatistics(latencies);
    std::cout << "\n=== " << benchmark_name << " Benchmark Results ===\n"
              << std::fixed << std::setprecision(3) << "Iterations: " << s.count << "\nMean:       " << s.mean
              << " ms\nMedian:     " << s.median << " ms\nStd Dev:    " << s.std_dev << " ms\nMin:        " << s.min_val
              << " ms\nMax:        " << s.max_val << " ms\nP95:        " << s.p95 << " ms\nP99:        " << s.p99
              << " ms\n==========================================\n\n";
}

}  // namespace BenchmarkUtils
