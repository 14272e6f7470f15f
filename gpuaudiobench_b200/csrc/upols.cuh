// upols.cuh — launch interface of the uniformly-partitioned overlap-save kernels (upols.cu).
#pragma once

#include <cuda_runtime.h>

#include "bus_tree.cuh"
#include "strip.cuh"
#include <stddef.h>

namespace b200conv {

// Forward real FFT of `count` windows of N = 2M samples each; window w is
//   [ first[w*first_stride .. +M) | second[w*second_stride .. +M) ]   (a null pointer reads zeros)
// and its M packed bins (bin 0 = {DC, Nyquist}) go to out[w*out_stride .. +M), scaled by `scale`.
// When prev_out != null the second half is also copied to prev_out[w*M ..] (the engine's
// "previous buffer" for the next call).
struct RfftParams {
    const float* first;
    size_t first_stride;
    const float* second;
    size_t second_stride;
    float2* out;
    size_t out_stride;  // in float2
    float* prev_out;
    int count;
    int M;      // complex FFT size = bins per partition = B
    int logM;
    float scale;
    int unpacked;  // 0: M packed bins (bin 0 = {DC, Nyquist}); 1: M+1 plain complex bins (cuFFT R2C layout)
};

// Y[s][t][k] = sum_{p in split s} H[t][p][k] * X[t][(slot0 + p) mod P][k]
struct MacParams {
    const float2* H;   // [T][P][M] partition spectra (packed bins, pre-scaled by 1/N)
    const float2* X;   // [T][P][M] frequency-domain delay line (ring of packed spectra)
    float2* Ypart;     // [S][T][M]
    int T, P, M;
    int slot0;         // ring slot of the newest block (p = 0)
    int S;             // partition splits (grid.y)
};

struct IrfftParams {
    const float2* Ypart;  // [S][T][M]
    int S;
    float* out;           // [T][B] or [B][Tg]
    int T, M, logM;
    int sample_major, Tg, toff;
};

// One launch per block for M = B <= 2048: forward FFT (split 0) + FDL-MAC + inverse/overlap-save (last CTA).
struct FusedParams {
    const float* d_in;    // [T][B]
    const float* prev;    // [T][B] previous buffer
    float* prev_w;        // [T][B] where this buffer is kept for the next block when commit != 0: the other half of a
                          // ping-pong (several CTAs of a track read `prev` while tile 0 of split 0 writes)
    const float2* H;      // [T][P][M]
    float2* X;            // [T][P][M] ring; slot0 receives X_m
    float2* Ypart;        // [S][T][M]
    unsigned* counters;   // [T], zero between launches
    float* out;           // [T][B] or [B][Tg]
    float* out2;          // optional second copy of the output, same layout (pinned host memory: the
                          // last CTA of each track posts its PCIe writes while other CTAs still stream)
    int T, P, M, logM, S, slot0, commit;
    int KT;               // bin tiles of 512 bins = max(1, M / 512); grid (S * KT, T)
    unsigned* xready;     // [T] KT > 1: tile 0 stores `seq` here when X_m is in the ring
    unsigned seq;         // launch sequence number, never 0, different for every launch
    int sample_major, Tg, toff;
    StripParams strip;    // strip.ops != 0: the last CTA of a track runs the channel strip on its B output samples
                          // in shared memory before writing them (in/out/T/B/layout fields unused here)
    BusTreeParams bus;    // bus.mix != null: the stereo bus (and its multi-GPU sum) as an epilogue of this launch
};
cudaError_t launch_upols_fused(const FusedParams& p, cudaStream_t st);
constexpr int kFusedMaxM = 2048;  // measured at 512 tracks x 96000 taps: fused 129.8 / 138.5 us at B = 1024 / 2048 (three-kernel
                                  // path: 149 / 158); at B = 4096 the fused kernel (64 KB of FFT ping-pong -> 3 CTAs per SM, 24
                                  // partitions per CTA, two 4096-point transforms per track on single CTAs) took 236 us against
                                  // 192 us for the three-kernel path, which therefore stays for B >= 4096

cudaError_t launch_rfft_fwd(const RfftParams& p, cudaStream_t st);
cudaError_t launch_fdl_mac(const MacParams& p, cudaStream_t st);
cudaError_t launch_irfft_ols(const IrfftParams& p, cudaStream_t st);

}  // namespace b200conv
