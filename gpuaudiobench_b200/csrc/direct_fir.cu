// direct_fir.cu — direct-form per-track FIR for sm_100a.
//
// Replaces Conv1DTextureMemoryImplKernel (reference cuda/bench_conv1d.cu:7-27: one thread per
// track, B*L serial taps, texture fetch + stride-B global load per tap) with a design bounded by
// the FP32 FMA pipe:
//
//   y_t[n0+i] = sum_j h_t[j] * x_t[n0+i-j]          ("Direct engine")
//
// Work is cut into 16-sample output blocks (index a) and 16-tap blocks (index c).  For one (a, c)
// pair the 16x16 Toeplitz product needs exactly two 16-sample input blocks, D_{a-c-1} | D_{a-c},
// so a lane that walks c upward re-uses one of them: per 256 FMAs it loads 16 taps + 16 samples
// (8 LDS.128).  Lanes of a warp own different output blocks (taps are a shared-memory
// broadcast), warps own different tap ranges, and a CTA owns (track, 512-output tile, tap split).
// Taps and history are kept PRE-SWIZZLED in HBM (common.cuh: swz_chunk) so that one elected
// producer lane stages them with 1-D TMA bulk copies (cp.async.bulk -> SASS UBLKCP) into a
// multi-stage mbarrier ring while 8 consumer warps stay on the FMA pipe; the 64 B lane stride of
// the block reads is bank-conflict free because of that swizzle.
//
// Kernels: ring_append_kernel (new block -> history ring), fir_direct_kernel<A> (the hot kernel),
// fir_finish_kernel (fixed-order sum of the tap-split partials + output layout).
#include "direct_fir.cuh"

#include "common.cuh"

namespace b200conv {

// ---------------------------------------------------------------------------------------------
// History ring append: ring[t][swz(pos + i)] = in[t][i].  One float4 per thread.
// ---------------------------------------------------------------------------------------------
__global__ void ring_append_kernel(const float4* __restrict__ in, float4* __restrict__ ring, int T, int B4,
                                   int cap4, int pos4) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= T * B4) return;
    int t = idx / B4;
    int f = idx - t * B4;
    ring[static_cast<size_t>(t) * cap4 + swz_chunk(static_cast<uint32_t>(pos4 + f))] = in[idx];
}

// ---------------------------------------------------------------------------------------------
// The FIR kernel.
// ---------------------------------------------------------------------------------------------
// Load one 16-float block (index blk) of a swizzled tile into registers: four conflict-free
// LDS.128.  Physical chunk of logical chunk 4*blk+i is 8*(blk>>1) + (hi | (i ^ m)).
__device__ __forceinline__ void load_block(float (&v)[16], const float* tile, int blk) {
    const uint32_t m = blk & 3;
    const uint32_t hi = ((blk ^ (blk >> 2)) & 1) << 2;
    const float4* row = reinterpret_cast<const float4*>(tile) + ((blk >> 1) << 3);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float4 q = row[hi | (i ^ m)];
        v[4 * i + 0] = q.x;
        v[4 * i + 1] = q.y;
        v[4 * i + 2] = q.z;
        v[4 * i + 3] = q.w;
    }
}

// acc[r] += sum_s hv[s] * w[16 + r - s],  w = [lo | hi]  (256 FFMA, all indices static).
__device__ __forceinline__ void toeplitz_tile(float (&acc)[16], const float (&hv)[16], const float (&lo)[16],
                                              const float (&hi)[16]) {
#pragma unroll
    for (int s = 0; s < 16; ++s) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int idx = 16 + r - s;
            const float xv = (idx >= 16) ? hi[idx - 16] : lo[idx];
            acc[r] = fmaf(hv[s], xv, acc[r]);
        }
    }
}

template <int A>
__global__ void __launch_bounds__(kFirThreads, 2) fir_direct_kernel(FirParams p) {
    constexpr int CL = 32 / A;  // tap groups per warp
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* empty_bar = full_bar + kFirMaxStages;
    float* stage_base = reinterpret_cast<float*>(smem_raw + 128);
    const int stage_floats = (p.JSb + p.xtile_blocks) * 16;
    float* red = stage_base + static_cast<size_t>(p.nbuf) * stage_floats;

    const int s = blockIdx.x;   // tap split
    const int ot = blockIdx.y;  // 16*A-output tile
    const int t = blockIdx.z;   // track
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int a0 = ot * A;                 // first output block of this tile
    const int cs0 = s * p.nst * p.JSb;     // first tap block of this split
    const int qbase = p.posb + p.capb + a0;  // unwrapped ring block index of output block a0

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.nbuf; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], kFirWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == kFirWarps) {
        // ===== producer: one lane drives the TMA engine =====
        if (lane == 0) {
            const float* hsrc = p.h + static_cast<size_t>(t) * p.Lc * 16;
            const float* rsrc = p.ring + static_cast<size_t>(t) * p.capb * 16;
            for (int k = 0; k < p.nst; ++k) {
                const int slot = k % p.nbuf;
                const int round = k / p.nbuf;
                if (round > 0) mbar_wait(&empty_bar[slot], (round - 1) & 1);
                const int c0 = cs0 + k * p.JSb;
                const int qs = (qbase - c0 - p.JSb) & ~7;        // tile start, 512 B aligned in the ring
                const int nblk = (qbase + A - 1 - c0) - qs + 1;  // <= xtile_blocks
                const int src_b = qs % p.capb;
                const int first = min(nblk, p.capb - src_b);
                float* hs = stage_base + static_cast<size_t>(slot) * stage_floats;
                float* xs = hs + p.JSb * 16;
                mbar_arrive_expect_tx(&full_bar[slot], static_cast<uint32_t>((p.JSb + nblk) * 64));
                bulk_g2s(hs, hsrc + static_cast<size_t>(c0) * 16, static_cast<uint32_t>(p.JSb * 64), &full_bar[slot]);
                bulk_g2s(xs, rsrc + static_cast<size_t>(src_b) * 16, static_cast<uint32_t>(first * 64), &full_bar[slot]);
                if (first < nblk)
                    bulk_g2s(xs + first * 16, rsrc, static_cast<uint32_t>((nblk - first) * 64), &full_bar[slot]);
            }
        }
        __syncwarp();
    } else {
        // ===== consumers: 8 warps on the FMA pipe =====
        const int a = lane & (A - 1);
        const int g = lane / A;
        float acc[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) acc[r] = 0.0f;

        for (int k = 0; k < p.nst; ++k) {
            const int slot = k % p.nbuf;
            const int round = k / p.nbuf;
            mbar_wait(&full_bar[slot], round & 1);
            const int c0 = cs0 + k * p.JSb;
            const int qs = (qbase - c0 - p.JSb) & ~7;
            const float* hs = stage_base + static_cast<size_t>(slot) * stage_floats;
            const float* xs = hs + p.JSb * 16;
            const int hb = (warp * CL + g) * p.SPS;     // lane's first tap block inside the stage
            const int sb = qbase + a - (c0 + hb) - qs;  // smem block index of D_{a-c} for c = c0+hb

            float P[16], Q[16], hv[16];
            load_block(Q, xs, sb);
            for (int q = 0; q < p.SPS; q += 2) {
                load_block(hv, hs, hb + q);
                load_block(P, xs, sb - q - 1);
                toeplitz_tile(acc, hv, P, Q);
                load_block(hv, hs, hb + q + 1);
                load_block(Q, xs, sb - q - 2);
                toeplitz_tile(acc, hv, Q, P);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[slot]);
        }

        // tap groups of one warp -> lanes 0..A-1
        if (CL > 1) {
#pragma unroll
            for (int off = A; off < 32; off <<= 1) {
#pragma unroll
                for (int r = 0; r < 16; ++r) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], off);
            }
        }
        if (g == 0) {
            float4* dst = reinterpret_cast<float4*>(red + (warp * A + a) * 16);
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[i] = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
        }
    }
    __syncthreads();

    // warps -> one partial per output, summed in warp order (deterministic)
    constexpr int OT = A * 16;
    float* dst = p.partial + (static_cast<size_t>(s) * p.T + t) * p.B + ot * OT;
    for (int o = threadIdx.x; o < OT; o += kFirThreads) {
        float v = 0.0f;
#pragma unroll
        for (int w = 0; w < kFirWarps; ++w) v += red[w * OT + o];
        dst[o] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// Finish: y = sum over tap splits (fixed order), written track-major [T][B] or as this engine's
// column tile of the sample-major [B][Tg] matrix (bench_conv1d_accel.cu:249 layout).
// ---------------------------------------------------------------------------------------------
__global__ void fir_finish_kernel(const float* __restrict__ partial, float* __restrict__ out, int S, int T, int B,
                                  int sample_major, int Tg, int toff) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = blockIdx.y;
    if (n >= B) return;
    float v = 0.0f;
    for (int s = 0; s < S; ++s) v += partial[(static_cast<size_t>(s) * T + t) * B + n];
    if (sample_major)
        out[static_cast<size_t>(n) * Tg + toff + t] = v;
    else
        out[static_cast<size_t>(t) * B + n] = v;
}

// ---------------------------------------------------------------------------------------------
// Host-side launchers
// ---------------------------------------------------------------------------------------------
cudaError_t launch_ring_append(const float* d_in, float* ring, int T, int B, int cap, int pos, cudaStream_t st) {
    const int total = T * (B / 4);
    ring_append_kernel<<<(total + 255) / 256, 256, 0, st>>>(reinterpret_cast<const float4*>(d_in),
                                                            reinterpret_cast<float4*>(ring), T, B / 4, cap / 4, pos / 4);
    return cudaGetLastError();
}

template <int A>
static cudaError_t launch_fir_t(const FirParams& p, dim3 grid, size_t smem, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(fir_direct_kernel<A>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(kFirMaxSmem));
        if (e != cudaSuccess) return e;
        configured = true;
    }
    fir_direct_kernel<A><<<grid, kFirThreads, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_fir(const FirParams& p, int A, int S, int ntiles, size_t smem, cudaStream_t st) {
    dim3 grid(S, ntiles, p.T);
    switch (A) {
        case 32: return launch_fir_t<32>(p, grid, smem, st);
        case 16: return launch_fir_t<16>(p, grid, smem, st);
        case 8: return launch_fir_t<8>(p, grid, smem, st);
        case 4: return launch_fir_t<4>(p, grid, smem, st);
        case 2: return launch_fir_t<2>(p, grid, smem, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_fir_finish(const float* partial, float* out, int S, int T, int B, int sample_major, int Tg,
                              int toff, cudaStream_t st) {
    dim3 grid((B + 127) / 128, T);
    fir_finish_kernel<<<grid, 128, 0, st>>>(partial, out, S, T, B, sample_major, Tg, toff);
    return cudaGetLastError();
}

}  // namespace b200conv
