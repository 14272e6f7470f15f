// direct_fir.cuh — launch interface of the direct-form FIR kernels (direct_fir.cu).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>

namespace b200conv {

constexpr int kFirWarps = 8;                        // consumer warps per CTA
constexpr int kFirThreads = (kFirWarps + 1) * 32;   // + one TMA producer warp
constexpr int kFirMaxStages = 8;                    // mbarrier slots reserved in shared memory
constexpr size_t kFirMaxSmem = 112 * 1024;          // per CTA; two CTAs per SM fit in 227 KB

// All "block" quantities are in units of 16 floats (64 B).
struct FirParams {
    const float* h;     // [T][Lc*16]  taps, zero padded, chunk-swizzled
    const float* ring;  // [T][capb*16] input history ring, chunk-swizzled
    float* partial;     // [S][T][B]   per-tap-split partial outputs
    int T, B;
    int capb;           // ring capacity
    int posb;           // ring block index where the current buffer starts
    int Lc;             // padded tap blocks per track = S * nst * JSb
    int JSb;            // tap blocks per pipeline stage = kFirWarps * (32/A) * SPS
    int nst;            // stages per CTA
    int SPS;            // 16-tap steps per lane per stage (even)
    int nbuf;           // pipeline depth (<= kFirMaxStages)
    int xtile_blocks;   // shared-memory blocks reserved for the input window of one stage
};

cudaError_t launch_ring_append(const float* d_in, float* ring, int T, int B, int cap, int pos, cudaStream_t st);
cudaError_t launch_fir(const FirParams& p, int A, int S, int ntiles, size_t smem, cudaStream_t st);
cudaError_t launch_fir_finish(const float* partial, float* out, int S, int T, int B, int sample_major, int Tg,
                              int toff, cudaStream_t st);

}  // namespace b200conv
