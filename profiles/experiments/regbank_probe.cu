// regbank_probe.cu — which FFMA operand patterns cost more than one issue cycle on sm_100a?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o regbank_probe regbank_probe.cu
// Every variant runs the same 256 independent-chain FFMAs per loop trip on fixed registers
// (x, h loaded once as float4 quads; acc stored as float4 quads, so all three arrays sit in
// 4-aligned register quads and index parity == register parity — check with cuobjdump -sass).
//   SHIFT variants: acc[i] += h[j] * x[(i + SHIFT) & 15]   (h[j] in the reuse cache for 16 FFMAs)
//       SHIFT 0: x and acc same index (same parity, same mod 4)   1: opposite parity
//       SHIFT 2: same parity, different mod 4                     4: same mod 4, different mod 8
//   NOREUSE: acc[i] += h[(i + j) & 15] * x[(i + 1) & 15]          (three fresh register reads)
//   PEAK:    acc[i] = acc[i] * h[0] + x[0]                        (two operands in the reuse cache)
#include <cuda_runtime.h>
#include <cstdio>

template <int SHIFT, int MODE>
__global__ void __launch_bounds__(256) probe(float4* sink, const float4* src, int iters) {
    float acc[16], h[16], x[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 a = src[threadIdx.x * 4 + q], b = src[4096 + threadIdx.x * 4 + q];
        h[4 * q] = a.x; h[4 * q + 1] = a.y; h[4 * q + 2] = a.z; h[4 * q + 3] = a.w;
        x[4 * q] = b.x; x[4 * q + 1] = b.y; x[4 * q + 2] = b.z; x[4 * q + 3] = b.w;
        acc[4 * q] = acc[4 * q + 1] = acc[4 * q + 2] = acc[4 * q + 3] = 0.0f;
    }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (MODE == 0) acc[i] = fmaf(h[j], x[(i + SHIFT) & 15], acc[i]);
                if (MODE == 1) acc[i] = fmaf(h[(i + j) & 15], x[(i + 1) & 15], acc[i]);
                if (MODE == 2) acc[i] = fmaf(acc[i], h[0], x[0]);
            }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
        sink[(blockIdx.x * 256 + threadIdx.x) * 4 + q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
}

template <int SHIFT, int MODE>
static void run(const char* name, float4* sink, float4* src, int sms) {
    const int iters = 4096, blocks = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<SHIFT, MODE><<<blocks, 256>>>(sink, src, 64);
    cudaDeviceSynchronize();
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        probe<SHIFT, MODE><<<blocks, 256>>>(sink, src, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double flop = 2.0 * 256.0 * iters * 256.0 * blocks;
    printf("%-28s %8.3f ms  %7.2f TFLOP/s\n", name, best, flop / best * 1e-9);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float4 *sink, *src;
    cudaMalloc(&sink, sizeof(float4) * 4 * 256 * sms * 8);
    cudaMalloc(&src, sizeof(float4) * 8192);
    cudaMemset(src, 0, sizeof(float4) * 8192);
    run<0, 2>("peak (2 reuse operands)", sink, src, sms);
    run<0, 0>("shift 0 (same reg idx)", sink, src, sms);
    run<1, 0>("shift 1 (opposite parity)", sink, src, sms);
    run<2, 0>("shift 2 (same par, !=mod4)", sink, src, sms);
    run<3, 0>("shift 3 (opposite parity)", sink, src, sms);
    run<4, 0>("shift 4 (same mod4)", sink, src, sms);
    run<8, 0>("shift 8 (same mod8)", sink, src, sms);
    run<0, 1>("no reuse (3 fresh reads)", sink, src, sms);
    return cudaDeviceSynchronize() != cudaSuccess;
}
