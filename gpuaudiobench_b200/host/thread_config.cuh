#pragma once
// Launch presets.  The reference's cuda/thread_config.cuh:6-9 sizes its one-thread-per-track
// kernels with these; the B200 engine sizes its own grids from the SM count (b200conv_plan), so
// here they only parameterise GPUABenchmark::calculateGridDimensions, kept for API parity.
#include <cstddef>

namespace ThreadConfig {

constexpr int SMALL_BLOCK_SIZE_1D = 128;
constexpr int DEFAULT_BLOCK_SIZE_1D = 256;
constexpr int LARGE_BLOCK_SIZE_1D = 512;
constexpr int MAX_BLOCK_SIZE_1D = 1024;

// ceil(totalThreads / blockSize), as cuda/thread_config.cuh:22-25
inline int calculateGridSize1D(size_t totalThreads, int blockSize = DEFAULT_BLOCK_SIZE_1D) {
    return static_cast<int>((totalThreads + static_cast<size_t>(blockSize) - 1) / static_cast<size_t>(blockSize));
}

}  // namespace ThreadConfig
