// tc_toeplitz.cuh — launch interface of the tensor-core direct-form FIR (tc_toeplitz.cu).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "bus_tree.cuh"

namespace b200conv {

constexpr int kTcRows = 128;    // MMA M: rows of a slab = samples per tap column (K depth of one row-block)
constexpr int kTcMaxCols = 128; // MMA N <= 128: output-slab columns per group (TMEM accumulator columns); one group = one work item
constexpr int kTcKSteps = 16;   // 128 taps / 8 (K of one kind::tf32 instruction)
constexpr int kTcPlanes = 32;   // 128 taps / 4 (16-byte K chunks)
constexpr int kTcThreads = 160; // 4 epilogue/band warps + 1 load/MMA warp
constexpr int kTcMaxA = 8;      // block <= 1024
constexpr int kTcTraceSlots = 16;  // diagnostics: stamps per CTA

struct TcGeometry {
    int A;      // row blocks per buffer = B / 128
    int C;      // tap columns = ceil(L / 128)
    int NE;     // slab columns that receive a contribution = C + A - 1; columns < A are the buffer's own samples (FP32 FMA)
    int N;      // MMA N: columns per group = min(128, roundup(C - 1, 16))
    int NGRP;   // column groups of N over the tensor-core columns A .. NE-1
    int R;      // image rows per group = N + A - 1
    int tmem_cols;  // TMEM allocation: power of two >= 3 N (two accumulators + the staged ring values)
    int capP;   // pending-output ring capacity in floats: roundup(128 * NE, B)
    size_t image_floats;  // floats of one (track, group, part) image = 32 * R * 4
    size_t smem_bytes;
};
TcGeometry tc_geometry(int B, int L);

struct TcParams {
    const float* d_in;   // [T][B]
    float* xprev;        // [2][T][128] the 128 samples before the current buffer, ping-pong: read [xpar], write [xpar ^ 1]
    const float* bimg;   // [T][NGRP][2][32][R][4] tap images (hi part, lo part), zero padded
    const float* hhead;  // [T][B] the first B taps as they are (zero padded): the buffer's own samples, FP32 FMA
    float* pend;         // [T][capP] pending-output ring
    float* out;          // [T][out_stride] or column tile of [B_full][Tg]
    // B is the block of THIS launch.  A caller's buffer longer than 1024 samples is streamed through the kernel in
    // sub-blocks (the engine state is a stream state): the launch then works on samples n_off .. n_off + B of rows
    // that are in_stride / out_stride / bus.B floats long, and on the bus chunks chunk0 .. chunk0 + nchunk.
    int in_stride, out_stride, n_off, chunk0, nchunk;
    int T, B, A, C, NE, N, NGRP, R, capP;
    uint32_t tmem_cols;
    int ppos;            // ring index of output sample 0 of the current buffer
    int xpar;            // which half of xprev holds the previous buffer's tail
    int commit;
    int sample_major, Tg, toff;
    int debug;           // B200CONV_TC_DEBUG (measurement only): 1 skip the MMAs, 2 skip the pending-ring traffic, 4 skip the image load,
                         // 32 skip staging the ring values in TMEM (wrong results, timing only)
    unsigned long long* trace;  // diagnostics (B200CONV_TC_TRACE=1): [grid][kTcTraceSlots] %globaltimer stamps, else null
    BusTreeParams bus;   // bus.mix == null: no bus
    BusSlice slice;      // slice.target != 0: the column-slice bus (every track on its own co-resident CTA), else the tree
};

cudaError_t launch_tc_toeplitz(const TcParams& p, int grid, cudaStream_t st);

// Host: the hi/lo TF32 images of one track's taps for every group, in the layout above.
void tc_build_images(const float* h, int L, const TcGeometry& g, float* dst);

}  // namespace b200conv
