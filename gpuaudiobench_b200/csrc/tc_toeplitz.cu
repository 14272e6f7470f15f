// tc_toeplitz.cu — direct-form FIR on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// The tensor-core variant of the direct engine named by the north star ("a tensor-core Toeplitz-tile x IR
// contraction variant"): the same sum as Conv1DTextureMemoryImplKernel (reference cuda/bench_conv1d.cu:7-27),
//     y_t[n] = sum_k h_t[k] x_t[n - k],
// written INPUT-side so that it becomes one dense GEMM per track and buffer with no padding waste
// (2*B*L flops) and a wide N:
//
//   taps in columns of 128:  k = 128 c + d          (c < C = ceil(L/128), d < 128)
//   rows in slabs  of 128:   rho = 128 a + r        (a < A = B/128,       r < 128)   rho = sample of this buffer
//   S[r][e] = sum_a sum_d x[128 a + r - d] * h[128 (e - a) + d]                       e = a + c  < NE = C + A - 1
//   and then simply          y[128 e + r] += S[r][e]   for EVERY e: the buffer's contribution to all outputs it
//                                                       reaches, this buffer's (e < A) and the L future ones.
//
// The future outputs live in a per-track pending-output ring (overlap-add in the time domain): a buffer adds
// S to it, emits its own B samples and hands the rest on.  S accumulates over (a, d) inside TMEM — 64 K-steps
// of one 128 x 80 x 8 kind::tf32 MMA each per 80-column group — so the skewed sum costs nothing outside the
// tensor core.
//
// Operands, both K-major, no swizzle ("interleaved" canonical layout: 8-row x 16-byte core matrices, 8-row
// groups SBO apart, the two 16-byte K chunks of an instruction LBO apart):
//   A (input, Hankel after reversing d):  A[r][d~] = x[128 a + r - 127 + d~].  With SBO = 128 B the rows are
//      16 B apart, and with LBO = 64 B a K chunk further is the same as 4 rows further — so ONE array of
//      "4-sample windows", band[g] = x[g-127 .. g-124], serves every (a, K-step) by moving the start address
//      (+2048 B per row block, +128 B per K-step).  No Toeplitz matrix is ever materialised: B+124 windows.
//   B (taps): image[plane S][row][4] = h[128 (row - (A-1) + e0) + 127 - (4 S + j)], rows 16 B apart
//      (SBO = 128 B), planes LBO = 16 R apart; the row block a reads it (A-1-a) rows down.  Built once per IR.
//
// fp32 accuracy from TF32 tensor cores: both operands are split x = hi + lo (hi = RN to TF32, lo = RN of the
// remainder) and three products are accumulated, lo*hi + hi*lo + hi*hi; the dropped lo*lo term and the two
// roundings of lo are 2^-22 relative.  Measured SNR against the fp32 oracle is stated in the parity tests.
#include "tc_toeplitz.cuh"

#include <algorithm>
#include <cstring>

#include "common.cuh"

namespace b200conv {

namespace {

constexpr int kTmemCols = 256;  // allocation: power of two >= 2 * kTcCols (two accumulators; two CTAs per SM use all 512 columns)
constexpr int kTmemAcc1 = 128;  // column of the second accumulator

__host__ __device__ constexpr int round_up(int v, int m) { return (v + m - 1) / m * m; }

// ---- tcgen05 / TMEM wrappers (PTX as CUTLASS's cute/arch/*_sm100*.hpp emits it) ----
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], one elected thread issues for the CTA
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all MMAs issued so far by this thread arrive on `bar` when they have completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, leading
// (K chunk) and stride (8-row group) byte offsets in 16-byte units, version 1 (Blackwell), layout type 0.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    return d;
}

// cute::UMMA::InstrDescriptor for kind::tf32: D fp32, A and B TF32, both K-major, M x N
__host__ __device__ constexpr uint32_t instr_desc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

__device__ __forceinline__ float tf32_rn(float v) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    return __uint_as_float(u);
}

// 32 lanes x 16 consecutive 32-bit columns of TMEM -> 16 registers per thread (lane = row)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct SmemMap {
    int xw_off, band_hi_off, band_lo_off, bimg_off, total;
};
__host__ __device__ inline SmemMap smem_map(int B, int R) {
    SmemMap m{};
    m.xw_off = 128;
    m.band_hi_off = m.xw_off + round_up((B + 128) * 4, 128);
    const int band_bytes = round_up((B + 127) * 16, 128);
    m.band_lo_off = m.band_hi_off + band_bytes;
    m.bimg_off = round_up(m.band_lo_off + band_bytes, 1024);
    m.total = m.bimg_off + 2 * kTcPlanes * R * 16;
    return m;
}

}  // namespace

__global__ void __launch_bounds__(kTcThreads, 2) tc_toeplitz_kernel(const __grid_constant__ TcParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bfull = reinterpret_cast<uint64_t*>(smem);       // tap images of a group have landed (TMA bytes)
    uint64_t* dfull = bfull + 1;                               // the group's MMAs have completed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 16);
    int* s_flag = reinterpret_cast<int*>(smem + 20);
    const SmemMap sm = smem_map(p.B, p.R);
    float* xw = reinterpret_cast<float*>(smem + sm.xw_off);    // xw[i] = x[i - 128]
    unsigned char* band_hi = smem + sm.band_hi_off;            // band[g] = x[g-127 .. g-124], g < B + 124
    unsigned char* band_lo = smem + sm.band_lo_off;
    unsigned char* bimg_s = smem + sm.bimg_off;                // [2 parts][32 planes][R rows][16 B]
    const uint32_t plane_bytes = static_cast<uint32_t>(p.R) * 16u;
    const uint32_t part_bytes = kTcPlanes * plane_bytes;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int B = p.B;

    if (tid == 0) {
        mbar_init(bfull, 1);
        mbar_init(dfull, 1);
        mbar_fence_init();
    }
    if (warp == 4) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    uint32_t bphase = 0, dphase = 0;
    constexpr uint32_t idesc = instr_desc_tf32(kTcRows, kTcCols);

    auto load_images = [&](int t, int grp) {  // one elected thread: 8 bulk copies of 8 planes each
        const unsigned char* src = reinterpret_cast<const unsigned char*>(p.bimg) +
                                   (static_cast<size_t>(t) * p.NGRP + grp) * 2 * part_bytes;
        if (p.debug & 4) {
            mbar_arrive(bfull);
            return;
        }
        mbar_arrive_expect_tx(bfull, 2 * part_bytes);
        const uint32_t piece = 8 * plane_bytes;
        for (int i = 0; i < 8; ++i) bulk_g2s(bimg_s + i * piece, src + static_cast<size_t>(i) * piece, piece, bfull);
    };

    // Work items are (track, column group): the kTcCols-column groups of one track are independent GEMMs over the
    // same band, so they go to different CTAs (two CTAs fit an SM: one's MMAs overlap the other's prologue /
    // epilogue; at C2 that is 256 items for 148 SMs instead of 128 tracks).  Each item rebuilds the small band.
    const int n_items = p.T * p.NGRP;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int t = item / p.NGRP, grp = item - t * p.NGRP;
        if (warp == 4) {
            if (lane == 0) load_images(t, grp);  // the image buffer is free: the previous item's MMAs completed
        } else {
            // ---- band of 4-sample windows of [previous 128 | this buffer], split hi + lo ----
            const float4* xin4 = reinterpret_cast<const float4*>(p.d_in + static_cast<size_t>(t) * B);
            const float4* xp4 = reinterpret_cast<const float4*>(p.xprev + (static_cast<size_t>(p.xpar) * p.T + t) * 128);
            float4* xw4 = reinterpret_cast<float4*>(xw);
            for (int i = tid; i < 32; i += 128) xw4[i] = xp4[i];
            for (int i = tid; i < B / 4; i += 128) xw4[32 + i] = xin4[i];
            named_bar_sync(1, 128);
            for (int g = tid; g < B + 124; g += 128) {  // the last window ends at x[B-1]
                float v[4], hi[4], lo[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[j] = xw[g + 1 + j];
                    hi[j] = tf32_rn(v[j]);
                    lo[j] = tf32_rn(v[j] - hi[j]);
                }
                *reinterpret_cast<float4*>(band_hi + g * 16) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4*>(band_lo + g * 16) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
            if (p.commit && grp == 0) {
                // the next buffer's "previous 128" goes to the OTHER half of the ping-pong: the CTAs that work on
                // this track's other column groups may still be reading the current one
                float4* xpw = reinterpret_cast<float4*>(p.xprev + (static_cast<size_t>(p.xpar ^ 1) * p.T + t) * 128);
                for (int i = tid; i < 32; i += 128) xpw[i] = xw4[B / 4 + i];
            }
            fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
        }
        __syncthreads();

        float* pring = p.pend + static_cast<size_t>(t) * p.capP;
        const int row = warp * 32 + lane;  // r (epilogue warps)
        {
            if (warp == 4) {
                if (lane == 0) {
                    mbar_wait(bfull, bphase);
                    tc_fence_after();
                    const uint32_t a_hi0 = smem_u32(band_hi), a_lo0 = smem_u32(band_lo);
                    const uint32_t b_hi0 = smem_u32(bimg_s), b_lo0 = b_hi0 + part_bytes;
                    // Two accumulators, even / odd K-steps, added in the epilogue in fp32 RN: the tensor core adds each
                    // MMA into TMEM with truncation, an error that grows with the number of accumulation steps
                    // (measured 102 dB at B = 1024 with one accumulator: 384 steps) — two chains of half the length
                    for (int a = 0; a < ((p.debug & 1) ? 0 : p.A); ++a) {
                        const uint32_t boff = 16u * static_cast<uint32_t>(p.A - 1 - a);
#pragma unroll 4
                        for (int q = 0; q < kTcKSteps; ++q) {
                            const uint32_t aoff = 2048u * a + 128u * q;
                            const uint64_t da_hi = smem_desc(a_hi0 + aoff, 64, 128);
                            const uint64_t da_lo = smem_desc(a_lo0 + aoff, 64, 128);
                            const uint64_t db_hi = smem_desc(b_hi0 + 2u * q * plane_bytes + boff, plane_bytes, 128);
                            const uint64_t db_lo = smem_desc(b_lo0 + 2u * q * plane_bytes + boff, plane_bytes, 128);
                            const uint32_t d = tmem + ((q & 1) ? kTmemAcc1 : 0);
                            mma_tf32(d, da_lo, db_hi, idesc, (a > 0 || q > 1) ? 1u : 0u);  // small terms first
                            mma_tf32(d, da_hi, db_lo, idesc, 1u);
                            mma_tf32(d, da_hi, db_hi, idesc, 1u);
                        }
                    }
                    mma_commit(dfull);
                }
                __syncwarp();
            } else {
                // ---- epilogue: S (TMEM) + pending ring -> this buffer's samples and the new pending ring.
                // Kept a compact LOOP over 16-column batches: the first version unrolled all 144 columns with
                // their ring arithmetic into 14 K instructions per warp and ncu showed 58 % of the stall samples
                // as "no instruction" (instruction-cache misses; cold they come from DRAM): 82 us for 6 us of MMA.
                const int e0 = grp * kTcCols;
                const int capP = p.capP, NE = p.NE, nA = p.A, ppos = p.ppos;
                const bool commit = p.commit != 0, ring_io = !(p.debug & 2);
                float* prow = pring + row;
                auto ring_index = [&](int e) {
                    int idx = ppos + 128 * e;
                    return idx >= capP ? idx - capP : idx;
                };
                // the ring lines this warp will read: into L2 while the MMAs run (first touch after a flush is DRAM)
                for (int k = lane; k < kTcCols; k += 32)
                    if (e0 + k < NE && ring_io)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(pring + ring_index(e0 + k) + warp * 32));
                float pn[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) pn[j] = (e0 + j < NE && ring_io) ? __ldcg(prow + ring_index(e0 + j)) : 0.0f;
                mbar_wait(dfull, dphase);
                tc_fence_after();
                const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
#pragma unroll 1
                for (int cb = 0; cb < kTcCols / 16; ++cb) {
                    float pv[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) pv[j] = pn[j];
                    const int eb = e0 + cb * 16;
                    if (cb + 1 < kTcCols / 16) {  // next batch's ring values in flight while this one is finished
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            pn[j] = (eb + 16 + j < NE && ring_io) ? __ldcg(prow + ring_index(eb + 16 + j)) : 0.0f;
                    }
                    uint32_t r[16], r1[16];
                    tmem_ld16(taddr + cb * 16, r);
                    tmem_ld16(taddr + kTmemAcc1 + cb * 16, r1);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int e = eb + j;
                        if (e < NE) {
                            const float v = (__uint_as_float(r[j]) + __uint_as_float(r1[j])) + pv[j];
                            if (e < nA) {  // this buffer's own samples: n = 128 e + row < B
                                const int n = 128 * e + row;
                                if (p.sample_major)
                                    p.out[static_cast<size_t>(n) * p.Tg + p.toff + t] = v;
                                else
                                    p.out[static_cast<size_t>(t) * B + n] = v;
                                if (p.bus.mix) p.bus.ybus[static_cast<size_t>(t) * B + n] = v;
                                if (commit) prow[ring_index(e)] = 0.0f;  // becomes the farthest future slot
                            } else if (commit && ring_io) {
                                prow[ring_index(e)] = v;
                            }
                        }
                    }
                }
                tc_fence_before();
            }
            bphase ^= 1u;
            dphase ^= 1u;
            __syncthreads();  // TMEM drained, band / x window / images free before the next item overwrites them
        }
        if (warp < 4 && p.bus.mix && grp == 0) {  // group 0 carries this buffer's own samples
            for (int chunk = 0; chunk < p.bus.NC; ++chunk) bus_tree_arrive<2>(p.bus, t, chunk, tid, 128, 1, s_flag);
        }
    }
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem, kTmemCols);
}

TcGeometry tc_geometry(int B, int L) {
    TcGeometry g{};
    g.A = B / kTcRows;
    g.C = (L + kTcRows - 1) / kTcRows;
    g.NE = g.C + g.A - 1;
    g.NGRP = (g.NE + kTcCols - 1) / kTcCols;
    g.R = kTcCols + g.A - 1;
    g.capP = round_up(128 * g.NE, B);
    g.image_floats = static_cast<size_t>(kTcPlanes) * g.R * 4;
    g.smem_bytes = static_cast<size_t>(smem_map(B, g.R).total);
    return g;
}

static float host_tf32_rn(float v) {  // cvt.rna.tf32.f32: round to nearest, ties away from zero, 10 mantissa bits
    uint32_t u;
    std::memcpy(&u, &v, 4);
    u += 0x1000u;
    u &= 0xFFFFE000u;
    float r;
    std::memcpy(&r, &u, 4);
    return r;
}

void tc_build_images(const float* h, int L, const TcGeometry& g, float* dst) {
    // dst [NGRP][2][32][R][4]:  image[S][row][j] = h[128 (e0 + row - (A-1)) + 127 - (4 S + j)], zero outside [0, L)
    for (int grp = 0; grp < g.NGRP; ++grp) {
        float* hi = dst + (static_cast<size_t>(grp) * 2) * g.image_floats;
        float* lo = hi + g.image_floats;
        for (int S = 0; S < kTcPlanes; ++S)
            for (int row = 0; row < g.R; ++row)
                for (int j = 0; j < 4; ++j) {
                    const long long c = static_cast<long long>(grp) * kTcCols + row - (g.A - 1);
                    const long long k = 128 * c + 127 - (4 * S + j);
                    float v = 0.0f;
                    if (c >= 0 && c < g.C && k >= 0 && k < L) v = h[k];
                    const float vh = host_tf32_rn(v);
                    const size_t o = (static_cast<size_t>(S) * g.R + row) * 4 + j;
                    hi[o] = vh;
                    lo[o] = host_tf32_rn(v - vh);
                }
    }
}

cudaError_t launch_tc_toeplitz(const TcParams& p, int grid, cudaStream_t st) {
    const size_t smem = static_cast<size_t>(smem_map(p.B, p.R).total);
    cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(tc_toeplitz_kernel), smem);
    if (e != cudaSuccess) return e;
    tc_toeplitz_kernel<<<grid, kTcThreads, smem, st>>>(p);
    return cudaGetLastError();
}

}  // namespace b200conv
