// ffma2_probe.cu — is packed fma.rn.f32x2 (SASS FFMA2) full-rate on B200, and does it free issue
// slots for other instructions?  nvcc -gencode arch=compute_100a,code=sm_100a -O3 ffma2_probe.cu
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ void ffma2(float2& d, float2 a, float2 b) {
    unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(*reinterpret_cast<unsigned long long*>(&a)),
                 "l"(*reinterpret_cast<unsigned long long*>(&b)));
    d = *reinterpret_cast<float2*>(&dd);
}

// MODE 0: scalar FFMA, acc = acc*a+b (2 operands in the reuse cache: the optimistic peak)
// MODE 1: scalar FFMA with three live register operands per instruction (acc[i] += h[j]*x[(i+j)&15])
// MODE 2: FFMA2 with three live register-pair operands
// MODE 3: MODE 2 plus one shared-memory load per 8 FFMA2 (issue slots shared with LDS)
template <int MODE>
__global__ void __launch_bounds__(256) probe(float* sink, const float* src, int iters) {
    __shared__ float sm[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) sm[i] = src[i];
    __syncthreads();
    float acc[16], h[16], x[16];
    float2 acc2[8], h2[8], x2[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) { acc[i] = i; h[i] = src[i + threadIdx.x]; x[i] = src[64 + i + threadIdx.x]; }
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc2[i] = make_float2(i, -i); h2[i] = make_float2(h[2*i], h[2*i+1]); x2[i] = make_float2(x[2*i], x[2*i+1]); }
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], h[0], x[0]);
        } else if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = fmaf(h[j], x[(i + j) & 15], acc[i]);
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
#pragma unroll
                for (int i = 0; i < 8; ++i) ffma2(acc2[i], h2[j & 7], x2[(i + j) & 7]);
                if (MODE == 3) {
                    float4 v = *reinterpret_cast<const float4*>(&sm[((threadIdx.x + j * 8 + it) & 255) * 4]);
                    x2[j & 7].x += v.x; 
                }
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc2[i].x + acc2[i].y;
    if (s == 1234.5f) sink[0] = s;
}

template <int MODE> void run(const char* name, float* sink, float* src, int sms) {
    const int iters = 2048, blocks = sms * 8 * 2;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a); probe<MODE><<<blocks, 256>>>(sink, src, iters); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (rep && ms < best) best = ms;
    }
    double fma = (double)blocks * 256 * iters * 256;  // 256 scalar FMAs per iteration in every mode
    printf("%-44s %8.3f ms  %7.2f TFLOP/s\n", name, best, 2 * fma / (best * 1e-3) / 1e12);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    float *sink, *src; cudaMalloc(&sink, 16); cudaMalloc(&src, 1 << 20); cudaMemset(src, 0, 1 << 20);
    run<0>("FFMA  acc=acc*a+b (reuse-cache operands)", sink, src, p.multiProcessorCount);
    run<1>("FFMA  acc+=h[j]*x[k] (3 live operands)", sink, src, p.multiProcessorCount);
    run<2>("FFMA2 acc2+=h2*x2 (3 live pair operands)", sink, src, p.multiProcessorCount);
    run<3>("FFMA2 + 1 LDS.128 per 8 FFMA2", sink, src, p.multiProcessorCount);
    return 0;
}
