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
// S to it, emits its own B samples and hands the rest on.  S accumulates over (a, d) inside TMEM — 16 A K-steps
// of three 128 x N x 8 kind::tf32 MMAs per N-column group (N <= 128) — so the skewed sum costs nothing outside
// the tensor core.
//
// The columns are split by WHEN they are needed.  The buffer's own samples are the columns e < A: they involve
// the first B taps only (1.6 % of the work at C2), and everything after the kernel — the stereo bus, its
// multi-GPU exchange, the caller — waits for them.  They are computed FIRST, in FP32 FMA by the four band /
// epilogue warps (not beside the MMAs: the tensor core's operand reads saturate shared memory), written out, and
// handed to the in-kernel bus (bus_tree.cuh), so that its chain of L2 round trips and the NVLink exchange run
// UNDER the MMAs instead of after them.  The tensor core does the columns A <= e < NE, which only feed the pending
// ring (C - 1 columns: 127 = one group of 128 at L = 16384).  One (group, track) item per CTA, one CTA per SM.
//
// Operands, both K-major, no swizzle ("interleaved" canonical layout: 8-row x 16-byte core matrices, 8-row
// groups SBO apart, the two 16-byte K chunks of an instruction LBO apart):
//   A (input, Hankel after reversing d):  A[r][d~] = x[128 a + r - 127 + d~].  With SBO = 128 B the rows are
//      16 B apart, and with LBO = 64 B a K chunk further is the same as 4 rows further — so ONE array of
//      "4-sample windows", band[g] = x[g-127 .. g-124], serves every (a, K-step) by moving the start address
//      (+2048 B per row block, +128 B per K-step).  No Toeplitz matrix is ever materialised: B+124 windows.
//   B (taps): image[plane S][row][4] = h[128 (row + 1 + N grp) + 127 - (4 S + j)] (row 0 of group 0 is tap
//      column 1: column e = A of row block a = A-1), rows 16 B apart (SBO = 128 B), planes LBO = 16 R apart;
//      the row block a reads it (A-1-a) rows down.  Built once per IR.
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

// 32 lanes x 32 consecutive 32-bit columns of TMEM -> 32 registers per thread (lane = row); no wait inside
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
          "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
          "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// An empty volatile asm that "rewrites" the registers of a tcgen05.ld: placed after tmem_wait_ld() it keeps the
// compiler from hoisting the arithmetic on them above the wait (volatile asms keep their order).
__device__ __forceinline__ void tmem_ld_landed(uint32_t (&r)[32]) {
#pragma unroll
    for (int h = 0; h < 32; h += 16)
        asm volatile("" : "+r"(r[h]), "+r"(r[h + 1]), "+r"(r[h + 2]), "+r"(r[h + 3]), "+r"(r[h + 4]), "+r"(r[h + 5]), "+r"(r[h + 6]),
                          "+r"(r[h + 7]), "+r"(r[h + 8]), "+r"(r[h + 9]), "+r"(r[h + 10]), "+r"(r[h + 11]), "+r"(r[h + 12]),
                          "+r"(r[h + 13]), "+r"(r[h + 14]), "+r"(r[h + 15])::"memory");
}

// diagnostics (B200CONV_TC_TRACE=1): %globaltimer stamps of the first item of every CTA, [grid][kTcTraceSlots]
__device__ __forceinline__ void tc_stamp(const TcParams& p, int slot) {
    if (p.trace) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        p.trace[static_cast<size_t>(blockIdx.x) * kTcTraceSlots + slot] = now;
    }
}

// 16 registers per thread -> 32 lanes x 16 consecutive 32-bit columns of TMEM
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

struct SmemMap {
    int xw_off, hs_off, red_off, part_off, band_hi_off, band_lo_off, bimg_off, total;
};
__host__ __device__ inline SmemMap smem_map(int B, int R) {
    SmemMap m{};
    m.xw_off = 128;
    m.hs_off = m.xw_off + round_up((B + 128) * 4, 128);
    m.red_off = m.hs_off + B * 4;               // [4 warps][A][32 lanes] float4 = 16 B bytes
    m.part_off = m.red_off + B * 16;            // column-slice bus: [128 threads][2] float4 + [128] segment sums
    m.band_hi_off = m.part_off + 4096 + 512;
    const int band_bytes = round_up((B + 127) * 16, 128);
    m.band_lo_off = m.band_hi_off + band_bytes;
    m.bimg_off = round_up(m.band_lo_off + band_bytes, 1024);
    m.total = m.bimg_off + 2 * kTcPlanes * R * 16;
    return m;
}

__device__ __forceinline__ void fma16(const float4& h, const float4& xa, const float4& xb, float4& acc) {
    // taps k .. k+3 (h) against the window xa | xb = x[m-4 .. m+3], outputs n0 .. n0+3 with m = n0 - k
    acc.x = fmaf(h.x, xb.x, acc.x); acc.y = fmaf(h.x, xb.y, acc.y); acc.z = fmaf(h.x, xb.z, acc.z); acc.w = fmaf(h.x, xb.w, acc.w);
    acc.x = fmaf(h.y, xa.w, acc.x); acc.y = fmaf(h.y, xb.x, acc.y); acc.z = fmaf(h.y, xb.y, acc.z); acc.w = fmaf(h.y, xb.z, acc.w);
    acc.x = fmaf(h.z, xa.z, acc.x); acc.y = fmaf(h.z, xa.w, acc.y); acc.z = fmaf(h.z, xb.x, acc.z); acc.w = fmaf(h.z, xb.y, acc.w);
    acc.x = fmaf(h.w, xa.y, acc.x); acc.y = fmaf(h.w, xa.z, acc.y); acc.z = fmaf(h.w, xa.w, acc.z); acc.w = fmaf(h.w, xb.x, acc.w);
}

}  // namespace

// kA: upper bound of the row blocks per buffer (the own-sample loop is unrolled over them; smaller = less code)
template <int kA>
__global__ void __launch_bounds__(kTcThreads, 1) tc_toeplitz_kernel(const __grid_constant__ TcParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bfull = reinterpret_cast<uint64_t*>(smem);       // tap images of a group have landed (TMA bytes)
    uint64_t* dfull = bfull + 1;                               // the group's MMAs have completed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 16);
    int* s_flag = reinterpret_cast<int*>(smem + 20);
    const SmemMap sm = smem_map(p.B, p.R);
    float* xw = reinterpret_cast<float*>(smem + sm.xw_off);    // xw[i] = x[i - 128]
    float4* hs4 = reinterpret_cast<float4*>(smem + sm.hs_off); // the first B taps (own samples)
    float4* red4 = reinterpret_cast<float4*>(smem + sm.red_off);
    float* bus_part = reinterpret_cast<float*>(smem + sm.part_off);
    unsigned char* band_hi = smem + sm.band_hi_off;            // band[g] = x[g-127 .. g-124], g < B + 124
    unsigned char* band_lo = smem + sm.band_lo_off;
    unsigned char* bimg_s = smem + sm.bimg_off;                // [2 parts][32 planes][R rows][16 B]
    const uint32_t plane_bytes = static_cast<uint32_t>(p.R) * 16u;
    const uint32_t part_bytes = kTcPlanes * plane_bytes;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int B = p.B, N = p.N;
    const uint32_t tmem_cols = p.tmem_cols;

    // Warp 4 sets up the barriers and the TMEM allocation while warps 0-3 already fetch the first item's input: no
    // CTA-wide barrier here, the first item's band barrier publishes both.
    if (warp == 4) {
        if (lane == 0) {
            mbar_init(bfull, 1);
            mbar_init(dfull, 1);
            mbar_fence_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, tmem_cols);
        tc_fence_before();
    }
    uint32_t tmem = 0;
    uint32_t bphase = 0, dphase = 0;
    const uint32_t idesc = instr_desc_tf32(kTcRows, N);

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

    // Work items are (column group, track): the N-column groups of one track are independent GEMMs over the same
    // band, so they go to different CTAs; each item rebuilds the small band.  Group 0 of every track comes first
    // in the item order: it also carries the buffer's own samples and the bus.
    const int n_items = p.T * p.NGRP;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int grp = item / p.T, t = item - grp * p.T;
        float* pring = p.pend + static_cast<size_t>(t) * p.capP;
        const int capP = p.capP, ppos = p.ppos;
        const bool commit = p.commit != 0, ring_io = !(p.debug & 2);
        const bool own = (grp == 0);
        uint32_t owed = 0;  // bus chunks whose multi-GPU sum this CTA still has to collect
        auto ring_index = [&](int e) {
            int idx = ppos + 128 * e;
            return idx >= capP ? idx - capP : idx;
        };
        const bool stamp = p.trace && item == static_cast<int>(blockIdx.x) && lane == 0;
        if (stamp && warp == 0) tc_stamp(p, 0);
        if (warp == 4) {
            if (lane == 0) load_images(t, grp);  // the image buffer is free: the previous item's MMAs completed
        } else {
            // ---- band of 4-sample windows of [previous 128 | this buffer], split hi + lo ----
            const float4* xin4 = reinterpret_cast<const float4*>(p.d_in + static_cast<size_t>(t) * p.in_stride + p.n_off);
            const float4* xp4 = reinterpret_cast<const float4*>(p.xprev + (static_cast<size_t>(p.xpar) * p.T + t) * 128);
            float4* xw4 = reinterpret_cast<float4*>(xw);
            for (int i = tid; i < 32; i += 128) xw4[i] = xp4[i];
            for (int i = tid; i < B / 4; i += 128) xw4[32 + i] = xin4[i];
            if (own) {
                const float4* hh4 = reinterpret_cast<const float4*>(p.hhead + static_cast<size_t>(t) * B);
                for (int i = tid; i < B / 4; i += 128) hs4[i] = __ldg(hh4 + i);
                // the ring lines of the own samples: towards L2 now (cold they come from DRAM)
                if (ring_io)
                    for (int i = tid; i < B / 32; i += 128)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(pring + ring_index(i >> 2) + (i & 3) * 32));
            }
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
            if (commit && own) {
                // the next buffer's "previous 128" goes to the OTHER half of the ping-pong: the CTAs that work on
                // this track's other column groups may still be reading the current one
                float4* xpw = reinterpret_cast<float4*>(p.xprev + (static_cast<size_t>(p.xpar ^ 1) * p.T + t) * 128);
                for (int i = tid; i < 32; i += 128) xpw[i] = xw4[B / 4 + i];
            }
            fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
        }
        __syncthreads();
        tc_fence_after();
        tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
        if (stamp && warp == 0) tc_stamp(p, 1);

        if (warp == 4) {
            // The own samples go first: the tensor core reads 8 KB of shared memory per MMA and starves every other
            // shared-memory client while it runs (measured: the FP32 loop below took 8 us under the MMAs, 1 us alone)
            if (own) named_bar_sync(2, kTcThreads);
            if (lane == 0) {
                mbar_wait(bfull, bphase);
                tc_fence_after();
                if (stamp) tc_stamp(p, 6);
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
                        const uint32_t d = tmem + ((q & 1) ? static_cast<uint32_t>(N) : 0u);
                        mma_tf32(d, da_lo, db_hi, idesc, (a > 0 || q > 1) ? 1u : 0u);  // small terms first
                        mma_tf32(d, da_hi, db_lo, idesc, 1u);
                        mma_tf32(d, da_hi, db_hi, idesc, 1u);
                    }
                }
                mma_commit(dfull);
                if (stamp) tc_stamp(p, 7);
            }
            __syncwarp();
        } else {
            const int e0 = p.A + grp * N;
            const int NE = p.NE;
            if (ring_io)  // the ring lines this warp will read: into L2 now (first touch after a flush is DRAM)
                for (int k = lane; k < N; k += 32)
                    if (e0 + k < NE) asm volatile("prefetch.global.L2 [%0];" ::"l"(pring + ring_index(e0 + k) + warp * 32));
            if (own) {
                // ---- this buffer's own samples (columns e < A) in FP32 FMA ----
                //   y[n] = ring[n] + sum_{k < 128 (e + 1)} h[k] x[n - k],   e = n / 128
                // (the taps beyond reach back past the previous 128 samples: earlier buffers put them in the ring).
                // Lane l owns the 4 outputs n0 = 128 e + 4 l of EVERY row block e; warp w owns taps 32 w .. 32 w + 31
                // of every tap column c (equal work for the four warps); per 16 FMAs one 16-byte window load, the
                // taps are a broadcast.  The four partial sums meet in shared memory, added in warp order.
                float4 acc[kA];
#pragma unroll
                for (int e = 0; e < kA; ++e) acc[e] = make_float4(0.f, 0.f, 0.f, 0.f);
                float4 rv[(kA + 3) / 4];  // the ring values of the row blocks this warp finalises (e = warp, warp + 4): in flight now
#pragma unroll
                for (int k = 0; k < (kA + 3) / 4; ++k)
                    rv[k] = (ring_io && warp + 4 * k < p.A) ? __ldcg(reinterpret_cast<const float4*>(pring + ring_index(warp + 4 * k) + 4 * lane))
                                                            : make_float4(0.f, 0.f, 0.f, 0.f);
                const float4* xw4 = reinterpret_cast<const float4*>(xw);
#pragma unroll 1
                for (int c = 0; c < p.A; ++c) {
                    float4 h[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) h[i] = hs4[32 * c + 8 * warp + i];
#pragma unroll
                    for (int e = 0; e < kA; ++e) {
                        if (e >= c && e < p.A) {
                            // tap quad kq = 32 c + 8 w + i against output quad qd = 32 e + l: window index 32 + qd - kq
                            const float4* base = xw4 + 32 + 32 * (e - c) + lane - 8 * warp;
                            float4 xb = base[0];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float4 xa = base[-(i + 1)];
                                fma16(h[i], xa, xb, acc[e]);
                                xb = xa;
                            }
                        }
                    }
                }
#pragma unroll
                for (int e = 0; e < kA; ++e)
                    if (e < p.A) red4[(warp * p.A + e) * 32 + lane] = acc[e];
                named_bar_sync(1, 128);
#pragma unroll
                for (int k = 0; k < (kA + 3) / 4; ++k) {
                    const int e = warp + 4 * k;
                    if (e >= p.A) break;
                    const int n0 = 128 * e + 4 * lane;
                    float4 v = red4[e * 32 + lane];
#pragma unroll
                    for (int w2 = 1; w2 < 4; ++w2) {
                        const float4 r = red4[(w2 * p.A + e) * 32 + lane];
                        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
                    }
                    float* rslot = pring + ring_index(e) + 4 * lane;
                    v.x += rv[k].x; v.y += rv[k].y; v.z += rv[k].z; v.w += rv[k].w;
                    if (p.sample_major) {
                        float* o = p.out + static_cast<size_t>(p.n_off + n0) * p.Tg + p.toff + t;
                        o[0] = v.x;
                        o[p.Tg] = v.y;
                        o[2 * static_cast<size_t>(p.Tg)] = v.z;
                        o[3 * static_cast<size_t>(p.Tg)] = v.w;
                    } else {
                        *reinterpret_cast<float4*>(p.out + static_cast<size_t>(t) * p.out_stride + p.n_off + n0) = v;
                    }
                    if (p.bus.mix) *reinterpret_cast<float4*>(p.bus.ybus + static_cast<size_t>(t) * p.bus.B + p.n_off + n0) = v;
                    if (commit && ring_io) *reinterpret_cast<float4*>(rslot) = make_float4(0.f, 0.f, 0.f, 0.f);  // becomes the farthest future slot
                }
                named_bar_arrive(2, kTcThreads);  // shared memory is the tensor core's from here
                if (stamp && warp == 0) tc_stamp(p, 2);
                // the bus: last-arriver tree over the tracks, and on a multi-GPU job its exchange — all of it under the MMAs
                // (a multi-GPU job: the CTA that completes the local bus pushes it to the peers here and comes back for
                // their values after its epilogue — by then they have arrived, nobody waits on NVLink)
                if (p.bus.mix) {
                    if (p.slice.target) {
                        if (bus_slice_reduce(p.bus, p.slice, t, p.n_off, B, tid, 1, bus_part, s_flag)) owed = 1u;
                    } else {
                        for (int chunk = 0; chunk < p.nchunk; ++chunk)
                            if (bus_tree_arrive<2, true>(p.bus, t, p.chunk0 + chunk, tid, 128, 1, s_flag)) owed |= 1u << chunk;
                    }
                }
                if (stamp && warp == 0) tc_stamp(p, 3);
            }
            // ---- epilogue: S (TMEM) + pending ring -> the new pending ring, columns e = A + N grp + j.
            // While the MMAs run, the ring values of the group's columns are staged in TMEM next to the accumulators
            // (columns 2N .. 3N; a warp owns its 32 lanes, so no other warp is involved): what follows the MMAs is
            // three TMEM loads, two adds and a store per value, in a compact loop.  (The first version unrolled every
            // column with its ring arithmetic into 14 K instructions per warp and ncu showed 58 % of the stall samples
            // as "no instruction" — the L1.5 instruction cache is 32 KB, beyond it code streams from L2 and, after a
            // flush, from DRAM: 82 us for 6 us of MMA.  Keep this kernel small.)
            const int row = warp * 32 + lane;  // r
            // the group's ring slots are consecutive and wrap at most once: column j is at pa + 128 j before the wrap
            // (j < nwrap) and at pb + 128 j = pa - capP + 128 j after it.  Both cases are a PREDICATED access with an
            // immediate offset: one warp per scheduler has nothing to hide a branch behind (the version with a
            // pointer select and a branch per column spent 1 us per 32 columns on an empty loop body).
            const int idx0 = ring_index(e0);
            float* pa = pring + idx0 + row;
            float* pb = pa - capP;
            const int nwrap = (capP - idx0) >> 7;
            const int ncol = min(N, NE - e0);
            const int n1 = min(nwrap, ncol);
            const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
            if (ring_io && !(p.debug & 32)) {
#pragma unroll 1
                for (int cb = 0; cb < N; cb += 32) {
                    uint32_t pn[32];
                    const float* pc = pa + 128 * cb;
                    const float* pd = pb + 128 * cb;
                    const int l1 = n1 - cb, l2 = nwrap - cb, l3 = ncol - cb;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float a = 0.0f;
                        if (j < l1) a = __ldcg(pc + 128 * j);
                        if (j >= l2 && j < l3) a = __ldcg(pd + 128 * j);
                        pn[j] = __float_as_uint(a);
                    }
                    tmem_st16(taddr + 2 * N + cb, reinterpret_cast<const uint32_t(&)[16]>(pn[0]));
                    if (cb + 16 < N) tmem_st16(taddr + 2 * N + cb + 16, reinterpret_cast<const uint32_t(&)[16]>(pn[16]));
                }
                tmem_wait_st();
            }
            if (stamp && warp == 0) tc_stamp(p, 8);
            mbar_wait(dfull, dphase);
            tc_fence_after();
            if (stamp && warp == 0) tc_stamp(p, 4);
            if (commit && ring_io) {
#pragma unroll 1
                for (int cb = 0; cb < N; cb += 32) {  // (N is a multiple of 16: the last batch may reach 16 columns past
                    uint32_t r[32], r1[32], rr[32];    //  the group — inside the allocation, never stored)
                    tmem_ld32(taddr + cb, r);  // three loads in flight, one wait
                    tmem_ld32(taddr + N + cb, r1);
                    tmem_ld32(taddr + 2 * N + cb, rr);
                    tmem_wait_ld();
                    tmem_ld_landed(r);
                    tmem_ld_landed(r1);
                    tmem_ld_landed(rr);
                    float* pc = pa + 128 * cb;
                    float* pd = pb + 128 * cb;
                    const int l1 = n1 - cb, l2 = nwrap - cb, l3 = ncol - cb;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float v = (__uint_as_float(r[j]) + __uint_as_float(r1[j])) + __uint_as_float(rr[j]);
                        if (j < l1) pc[128 * j] = v;
                        if (j >= l2 && j < l3) pd[128 * j] = v;
                    }
                    if (stamp && warp == 0 && (cb == 0 || cb == 64)) tc_stamp(p, cb == 0 ? 9 : 10);
                }
            }
            tc_fence_before();
            if (stamp && warp == 0) tc_stamp(p, 5);
            if (owed && p.slice.target) {
                bus_slice_finish(p.bus, p.slice, t, p.n_off, B, tid);
            } else {
                for (int chunk = 0; owed >> chunk; ++chunk)
                    if ((owed >> chunk) & 1u) bus_tree_finish<2>(p.bus, p.chunk0 + chunk, tid, 128);
            }
        }
        bphase ^= 1u;
        dphase ^= 1u;
        __syncthreads();  // TMEM drained, band / x window / images free before the next item overwrites them
    }
    __syncthreads();
    if (warp == 4) tmem_dealloc(*reinterpret_cast<volatile uint32_t*>(tmem_slot), tmem_cols);
}

TcGeometry tc_geometry(int B, int L) {
    TcGeometry g{};
    g.A = B / kTcRows;
    g.C = (L + kTcRows - 1) / kTcRows;
    g.NE = g.C + g.A - 1;
    // the tensor core does columns A .. NE-1 (C - 1 of them) in groups of N <= 128 (a multiple of 16): one shared-memory
    // read of the band serves N columns, and at N = 64 the MMAs were bound by shared-memory reads, not by the tensor pipe
    g.N = std::min(kTcMaxCols, std::max(16, round_up(g.C - 1, 16)));
    g.NGRP = std::max(1, (g.C - 1 + g.N - 1) / g.N);
    g.R = g.N + g.A - 1;
    g.tmem_cols = 32;  // two accumulators and the staged ring values: 3 N columns, a power of two
    while (g.tmem_cols < 3 * g.N) g.tmem_cols *= 2;
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
    // dst [NGRP][2][32][R][4]:  image[S][row][j] = h[128 c + 127 - (4 S + j)], tap column c = N grp + row + 1 (row block
    // a reads from row A-1-a: column e = A + N grp + j of row block a is tap column e - a), zero outside [0, L)
    for (int grp = 0; grp < g.NGRP; ++grp) {
        float* hi = dst + (static_cast<size_t>(grp) * 2) * g.image_floats;
        float* lo = hi + g.image_floats;
        for (int S = 0; S < kTcPlanes; ++S)
            for (int row = 0; row < g.R; ++row)
                for (int j = 0; j < 4; ++j) {
                    const long long c = static_cast<long long>(grp) * g.N + row + 1;
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
    auto launch = [&](auto kernel) -> cudaError_t {
        cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(kernel), smem);
        if (e != cudaSuccess) return e;
        kernel<<<grid, kTcThreads, smem, st>>>(p);
        return cudaGetLastError();
    };
    if (p.A <= 2) return launch(tc_toeplitz_kernel<2>);
    if (p.A <= 4) return launch(tc_toeplitz_kernel<4>);
    return launch(tc_toeplitz_kernel<kTcMaxA>);
}

}  // namespace b200conv
