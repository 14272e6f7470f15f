// bus_tree.cuh — the stereo mix bus as an EPILOGUE of the last compute kernel, and its multi-GPU sum.
//
// mix[c][n] = sum over tracks of gains[t][c] * y_t[n] is a reduction ACROSS tracks, while the
// convolution kernels finish tracks one by one in whatever order their CTAs retire.  Round 1 paid a
// second launch for it (fir_finish_mix_kernel / mix_cluster_kernel, 5-9 us) and a third for the
// multi-GPU all-reduce (bus_allreduce_kernel, 12-15 us of launch + flag latency).  Here both ride on
// "last arriver" tickets inside the convolution kernel itself:
//
//   level 0  the CTA that completes track t's columns [chunk*CH, +CH) has just written them to
//            ybus[t][..] (device scratch, track-major) and calls bus_tree_arrive();
//   level 1  tracks are grouped G1 at a time; the LAST arriver of a (group, chunk) sums the group's
//            rows in track order -> gpart[group][c][n];
//   level 2  the LAST group of a chunk sums the group partials in group order -> the local bus chunk;
//   level 3  (world > 1) that CTA pushes the chunk into its slot of EVERY peer's symmetric buffer over
//            NVLink (plain P2P stores), fences, raises its flag on every peer, acquire-polls the
//            world flags of its own buffer and adds the world slots in rank order.
//
// Every sum runs in a fixed order whoever executes it, so the bus is deterministic run to run and
// bit-identical on all ranks; no float atomics.  Counters re-arm themselves (the last arriver
// stores 0), so a launch needs no memset.  Per-track work is O(B) reads of L2-resident rows; the
// critical path after the last track is two ticket round trips (+ one NVLink flag round trip).
//
// Symmetric buffer layout per rank (b200conv_bus_buffer_bytes): float data[2][world][n] with n = 2*B,
// then uint32 flags[2][world][kBusMaxChunks]; slot = epoch & 1.  Two slots suffice: a rank cannot
// finish epoch e+1 before every peer has signalled e+1, which a peer only does after it has consumed
// epoch e.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace b200conv {

constexpr int kBusMaxWorld = 16;
constexpr int kBusMaxChunks = 16;            // B <= 8192 in chunks of >= 512 columns (or one chunk of B < 512)
constexpr unsigned kBusSpinLimit = 1u << 22;  // bounded wait for a peer (system-scope acquire polls: seconds)

struct BusTreeParams {
    const float* gains;  // [T][2]
    float* ybus;         // [T][B] rows the bus is summed from (track-major, device memory)
    float* gpart;        // [NG][2][B]
    unsigned* gcount;    // [NG][NC] arrival tickets, zero between launches
    unsigned* ccount;    // [NC]
    float* mix;          // [2][B] destination (device, or pinned host); null: no bus wanted, tree disabled
    int T, B;
    int G1, NG;          // tracks per group, groups
    int CH, NC;          // columns per chunk, chunks
    // multi-GPU exchange (world == 1: none)
    float* peers[kBusMaxWorld];
    int rank, world;
    uint32_t epoch;
    uint32_t* err;       // set to 1 if a peer did not signal within the spin bound
};

#ifdef __CUDACC__
__device__ __forceinline__ void bus_st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t bus_ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void bus_bar(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// All-reduce of the local bus chunk held in registers by threads tid < ... (columns col = tid + i*nthr):
// called by every thread of the arriving group.  `l`/`r` callbacks are avoided: the local chunk is first
// written to this rank's own slot like everybody else's, then summed in rank order.
__device__ __forceinline__ void bus_exchange_chunk(const BusTreeParams& bt, int chunk, int tid, int nthr, uint32_t bar_id) {
    const int n = 2 * bt.B;
    const int slot = bt.epoch & 1u;
    const size_t data_floats = static_cast<size_t>(2) * bt.world * n;
    const int c0 = chunk * bt.CH;
    // push: my chunk sits in my own slot already (written by the caller); copy it to every peer
    const float* mine = bt.peers[bt.rank] + (static_cast<size_t>(slot) * bt.world + bt.rank) * n;
    for (int col = tid; col < bt.CH; col += nthr) {
        const float l = mine[c0 + col], r = mine[bt.B + c0 + col];
        for (int p = 0; p < bt.world; ++p) {
            if (p == bt.rank) continue;
            float* dst = bt.peers[p] + (static_cast<size_t>(slot) * bt.world + bt.rank) * n;
            dst[c0 + col] = l;
            dst[bt.B + c0 + col] = r;
        }
    }
    __threadfence_system();
    bus_bar(bar_id, nthr);
    if (tid < bt.world) {
        uint32_t* peer_flags = reinterpret_cast<uint32_t*>(bt.peers[tid] + data_floats);
        bus_st_release_sys(peer_flags + (slot * bt.world + bt.rank) * kBusMaxChunks + chunk, bt.epoch);
        const uint32_t* my_flags = reinterpret_cast<const uint32_t*>(bt.peers[bt.rank] + data_floats);
        unsigned spins = 0;
        while (bus_ld_acquire_sys(my_flags + (slot * bt.world + tid) * kBusMaxChunks + chunk) != bt.epoch) {
            if (++spins > kBusSpinLimit) {
                *reinterpret_cast<volatile uint32_t*>(bt.err) = 1u;  // mapped host memory: a plain store
                break;
            }
        }
    }
    bus_bar(bar_id, nthr);
    const float* base = bt.peers[bt.rank] + static_cast<size_t>(slot) * bt.world * n;
    for (int col = tid; col < bt.CH; col += nthr) {
        float l = 0.0f, r = 0.0f;
        for (int q = 0; q < bt.world; ++q) {
            l += __ldcg(base + static_cast<size_t>(q) * n + c0 + col);
            r += __ldcg(base + static_cast<size_t>(q) * n + bt.B + c0 + col);
        }
        bt.mix[c0 + col] = l;
        bt.mix[bt.B + c0 + col] = r;
    }
}

// Called by all `nthr` threads of the arriving group (tid = 0 .. nthr-1; they synchronise on hardware
// barrier `bar_id`) after they have written ybus[t][chunk*CH .. +CH).  `flag` is one int of shared memory.
__device__ __forceinline__ void bus_tree_arrive(const BusTreeParams& bt, int t, int chunk, int tid, int nthr,
                                                uint32_t bar_id, int* flag) {
    const int g = t / bt.G1;
    const int gsize = min(bt.G1, bt.T - g * bt.G1);
    const int c0 = chunk * bt.CH;
    // ---- level 1: last arriver of (group, chunk) ----
    __threadfence();
    bus_bar(bar_id, nthr);
    if (tid == 0) {
        unsigned* cnt = bt.gcount + g * bt.NC + chunk;
        const unsigned ticket = (gsize > 1) ? atomicAdd(cnt, 1u) : 0u;
        const int last = (ticket == static_cast<unsigned>(gsize) - 1u);
        if (last && gsize > 1) *cnt = 0;  // re-armed for the next launch
        *flag = last;
    }
    bus_bar(bar_id, nthr);
    if (!*flag) return;
    __threadfence();
    const bool single_group = (bt.NG == 1);
    float* my_slot = nullptr;
    if (bt.world > 1)
        my_slot = bt.peers[bt.rank] + (static_cast<size_t>(bt.epoch & 1u) * bt.world + bt.rank) * (2 * bt.B);
    for (int col = tid; col < bt.CH; col += nthr) {
        float l = 0.0f, r = 0.0f;
        const float* row = bt.ybus + static_cast<size_t>(g) * bt.G1 * bt.B + c0 + col;
        const float2* gn = reinterpret_cast<const float2*>(bt.gains) + g * bt.G1;
        int tt = 0;
        for (; tt + 4 <= gsize; tt += 4) {  // loads of four rows in flight, adds in track order
            const float v0 = __ldcg(row + static_cast<size_t>(tt) * bt.B);
            const float v1 = __ldcg(row + static_cast<size_t>(tt + 1) * bt.B);
            const float v2 = __ldcg(row + static_cast<size_t>(tt + 2) * bt.B);
            const float v3 = __ldcg(row + static_cast<size_t>(tt + 3) * bt.B);
            const float2 g0 = gn[tt], g1 = gn[tt + 1], g2 = gn[tt + 2], g3 = gn[tt + 3];
            l = fmaf(g0.x, v0, l); r = fmaf(g0.y, v0, r);
            l = fmaf(g1.x, v1, l); r = fmaf(g1.y, v1, r);
            l = fmaf(g2.x, v2, l); r = fmaf(g2.y, v2, r);
            l = fmaf(g3.x, v3, l); r = fmaf(g3.y, v3, r);
        }
        for (; tt < gsize; ++tt) {
            const float v = __ldcg(row + static_cast<size_t>(tt) * bt.B);
            const float2 gg = gn[tt];
            l = fmaf(gg.x, v, l);
            r = fmaf(gg.y, v, r);
        }
        if (single_group) {
            if (bt.world == 1) {
                bt.mix[c0 + col] = l;
                bt.mix[bt.B + c0 + col] = r;
            } else {
                my_slot[c0 + col] = l;
                my_slot[bt.B + c0 + col] = r;
            }
        } else {
            bt.gpart[(static_cast<size_t>(g) * 2) * bt.B + c0 + col] = l;
            bt.gpart[(static_cast<size_t>(g) * 2 + 1) * bt.B + c0 + col] = r;
        }
    }
    if (!single_group) {
        // ---- level 2: last group of the chunk ----
        __threadfence();
        bus_bar(bar_id, nthr);
        if (tid == 0) {
            const unsigned ticket = atomicAdd(bt.ccount + chunk, 1u);
            const int last = (ticket == static_cast<unsigned>(bt.NG) - 1u);
            if (last) bt.ccount[chunk] = 0;
            *flag = last;
        }
        bus_bar(bar_id, nthr);
        if (!*flag) return;
        __threadfence();
        for (int col = tid; col < bt.CH; col += nthr) {
            float l = 0.0f, r = 0.0f;
            for (int gg = 0; gg < bt.NG; ++gg) {
                l += __ldcg(bt.gpart + (static_cast<size_t>(gg) * 2) * bt.B + c0 + col);
                r += __ldcg(bt.gpart + (static_cast<size_t>(gg) * 2 + 1) * bt.B + c0 + col);
            }
            if (bt.world == 1) {
                bt.mix[c0 + col] = l;
                bt.mix[bt.B + c0 + col] = r;
            } else {
                my_slot[c0 + col] = l;
                my_slot[bt.B + c0 + col] = r;
            }
        }
    }
    if (bt.world > 1) {
        // ---- level 3: NVLink exchange of this chunk ----
        __threadfence();
        bus_bar(bar_id, nthr);
        bus_exchange_chunk(bt, chunk, tid, nthr, bar_id);
    }
}
#endif  // __CUDACC__

}  // namespace b200conv
