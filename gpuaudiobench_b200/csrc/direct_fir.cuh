// direct_fir.cuh — launch interface of the direct-form FIR kernels (direct_fir.cu).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>

#include "bus_tree.cuh"

namespace b200conv {

#ifndef B200CONV_FIR_CTAS_PER_SM
#define B200CONV_FIR_CTAS_PER_SM 2
#endif
#ifndef B200CONV_FIR_WARPS
#define B200CONV_FIR_WARPS 8
#endif
constexpr int kFirWarps = B200CONV_FIR_WARPS;             // consumer warps per CTA
constexpr int kFirThreads = (kFirWarps + 1) * 32;         // + one TMA producer warp
constexpr int kFirMaxStages = 8;                          // mbarrier slots reserved in shared memory
constexpr int kFirCtasPerSm = B200CONV_FIR_CTAS_PER_SM;   // persistent grid = kFirCtasPerSm * SM count
constexpr size_t kFirMaxSmem = (kFirCtasPerSm >= 3 ? 74 : 112) * 1024;  // per CTA; kFirCtasPerSm CTAs fit in 227 KB
constexpr int kMixChunk = 8;                              // tracks a warp handles per step of the bus/finish kernels
constexpr int kBusWarps = 4;                              // warps per CTA of the bus/finish kernels

// All "block" quantities are in units of 16 floats (64 B).
struct FirParams {
    const float* h;     // [T][Lc*16]  taps, zero padded; chunk-swizzled iff 32/A > 1 (lanes differ in taps)
    const float* ring;  // [T][capb*16] input history ring, chunk-swizzled (holds samples BEFORE the current buffer)
    const float* d_in;  // [T][B]      the current buffer (read directly; appended to the ring by the finish kernel)
    float* partial;     // [MS][T][B]  one row per (CTA, track-tile) segment
    int T, B;
    int capb;           // ring capacity
    int posb;           // ring block index where the current buffer starts
    int Lc;             // padded tap blocks per track = NS * JSb
    int JSb;            // tap blocks per pipeline stage = kFirWarps * (32/A) * SPS
    int NS;             // tap stages per track-tile
    int SPS;            // 16-tap steps per lane per stage (even)
    int nbuf;           // pipeline depth (<= kFirMaxStages)
    int xtile_blocks;   // shared-memory blocks reserved for the input window of one stage
    int ntiles;         // 512-output tiles per track
    int U;              // units = T * ntiles * NS
    int G;              // CTAs (persistent grid)
};

cudaError_t launch_ring_append(const float* d_in, float* ring, int T, int B, int cap, int pos, cudaStream_t st);
cudaError_t launch_fir(const FirParams& p, int A, size_t smem, cudaStream_t st);
// Most partial rows any track-tile receives when U = n_tiles_total * NS units are split over G CTAs.
int fir_max_segments(int n_tiles_total, int NS, int G);
struct FinishParams {
    const float* partial;  // [MS][T][B]
    float* out;            // [T][B] or [B][Tg]
    int MS, T, B, sample_major, Tg, toff;
    const float* gains;    // [T][2]
    float* mix;            // [2][B] or null (no bus requested)
    const float* d_in;     // [T][B]
    float* ring;           // [T][cap] or null (PEEK: do not append)
    int cap, pos;
    int chunk;             // set by the launcher: tracks per warp step (<= kMixChunk)
    BusExchange x;         // x.world > 1: the cluster leaders exchange the bus over NVLink and sum in rank order
};
cudaError_t launch_fir_finish_mix(const FinishParams& p, cudaStream_t st);
// Deterministic stereo bus of an output already in memory: mix[c][n] = sum_t gains[t][c] * y_t[n].
// x.world > 1: the bus is exchanged over NVLink and summed in rank order inside this kernel.
cudaError_t launch_mix_cluster(const float* y, int sample_major, int Tg, int toff, const float* gains, float* mix, int T,
                               int B, const BusExchange& x, cudaStream_t st);

}  // namespace b200conv
