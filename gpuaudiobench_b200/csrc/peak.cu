// peak.cu — FP32 FMA peak microbenchmark: the roofline denominator of the direct FIR.
// MEASURED_PEAKS.json carries HBM GB/s and bf16 tensor TFLOP/s only; the direct form runs on the
// CUDA-core FMA pipe, so its peak is measured here the same way (a kernel that does nothing but
// independent FFMA chains on every SM, timed with CUDA events).
#include "../../include/b200conv.h"

#include <cuda_runtime.h>

#include <string>

namespace {

constexpr int kChains = 16;
constexpr int kIters = 8192;

__global__ void __launch_bounds__(256) fma_peak_kernel(float* sink, float a, float b) {
    float acc[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) acc[i] = static_cast<float>(threadIdx.x + i);
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += acc[i];
    if (s == 12345.678f) sink[0] = s;  // keeps the chains alive; practically never true
}

}  // namespace

extern "C" int b200conv_measure_fp32_peak(int device, double* tflops, double* elapsed_ms) {
    if (!tflops) return B200CONV_ERR_INVALID;
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return B200CONV_ERR_NO_DEVICE;
    if (cudaSetDevice(device) != cudaSuccess) return B200CONV_ERR_CUDA;
    float* sink = nullptr;
    if (cudaMalloc(&sink, 16) != cudaSuccess) return B200CONV_ERR_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = prop.multiProcessorCount * 8 * 4;  // 8 resident CTAs of 256 threads per SM, 4 waves
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        fma_peak_kernel<<<blocks, 256>>>(sink, 1.0000001f, 1e-9f);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) {
            cudaFree(sink);
            return B200CONV_ERR_CUDA;
        }
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    const double flop = 2.0 * kChains * static_cast<double>(kIters) * 256.0 * blocks;
    *tflops = flop / (best * 1e-3) / 1e12;
    if (elapsed_ms) *elapsed_ms = best;
    return B200CONV_OK;
}
