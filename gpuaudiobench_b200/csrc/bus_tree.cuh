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
//   level 2  the LAST group of a chunk sums the group partials in group order -> the local bus chunk
//            (an ordered running sum over groups was tried for UPOLS and lost: the CTAs of a wave retire
//            together, so the chain serialised ~14 groups at the end of C3: 171 -> 183 us);
//   level 3  (world > 1) that CTA pushes the chunk into its slot of EVERY peer's symmetric buffer over
//            NVLink and adds the world slots of its own buffer in rank order (bus_ll_* below).
//
// Every sum runs in a fixed order whoever executes it, so the bus is deterministic run to run and
// bit-identical on all ranks; no float atomics.  Counters re-arm themselves (the last arriver
// stores 0), so a launch needs no memset.  Per-track work is O(B) reads of L2-resident rows; the
// critical path after the last track is two ticket round trips (+ one NVLink flag round trip).
//
// Exchange protocol ("LL": data and flag travel in ONE 8-byte store, as in NCCL's low-latency protocol):
// the symmetric buffer of a rank is uint64 ll[2][world][n], n = 2*B, slot = epoch & 1; a value is written as
// (epoch << 32) | float bits with a single 8-byte store — 8-byte stores are single-copy atomic, so a reader
// that sees the epoch in the upper half has the value in the lower half.  No fence, no separate flag store,
// no barrier: every thread pushes its own values to all ranks and polls its own values from all ranks, and
// the cost is ONE NVLink one-way latency (the first version — P2P stores, __threadfence_system, flag store,
// acquire poll — paid ~5 us for two hops and a system fence).  Two slots suffice: a rank cannot push epoch e+2
// before it has read every peer's epoch e+1, which a peer only writes after it has consumed epoch e.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace b200conv {

constexpr int kBusMaxWorld = 16;
constexpr int kBusMaxChunks = 16;            // B <= 8192 in chunks of >= 512 columns (or one chunk of B < 512)
constexpr unsigned kBusSpinLimit = 1u << 22;  // bounded wait for a peer (volatile polls: seconds)

// The multi-GPU half of the bus: who the peers are.  world == 1: no exchange.
struct BusExchange {
    unsigned long long* peers[kBusMaxWorld];  // rank p's symmetric buffer uint64 [2][world][n], mapped on this device
    int rank, world;
    uint32_t epoch;  // >= 1, advances by one per exchanged block on every rank
    uint32_t debug;  // measurement only (B200CONV_BUS_DEBUG, profiles/experiments/n2_fixed_cost.py): 1 push to the own buffer
                     // only, 2 poll the own slot only, 4 collect right after the push instead of after the epilogue
    uint32_t* err;   // set to 1 if a peer's value did not arrive within the spin bound
    unsigned long long* trace;  // optional diagnostics (b200conv_bus_trace): [kBusTraceLen][2] = %globaltimer ns when this
                                // rank's bus was ready to push / when the summed bus was complete, indexed by epoch
};
constexpr int kBusTraceLen = 4096;

struct BusTreeParams {
    const float* gains;  // [T][2]
    float* ybus;         // [T][B] rows the bus is summed from (track-major, device memory)
    float4* gpart;       // [NG][B/2] group partials, one float4 {l0, r0, l1, r1} per column pair
    unsigned* gcount;    // [NG][NC] arrival tickets, zero between launches
    unsigned* ccount;    // [NC] group tickets
    float* mix;          // [2][B] destination (device, or pinned host); null: no bus wanted, tree disabled
    int T, B;
    int G1, NG;          // tracks per group, groups
    int CH, NC;          // columns per chunk, chunks
    BusExchange x;       // multi-GPU exchange (x.world == 1: none)
};

// The column-slice bus (bus_slice_* below): when every track of the job is finished by a DIFFERENT, co-resident CTA
// (tracks <= CTAs of one wave), the tracks meet at one counter and every CTA sums a slice of the bus columns over all
// tracks — one level of L2 round trips instead of the tree's two, and on a multi-GPU job every CTA pushes its own
// slice to the peers at once.
struct BusSlice {
    unsigned long long* arrive;  // device counter, never reset between launches: launch `seq` waits for T * seq
    unsigned long long target;   // 0: slice bus off (the ticket tree is used)
    int slice;                   // columns per CTA: a power of two >= 4, slice * T >= columns of the launch
};

#ifdef __CUDACC__
__device__ __forceinline__ void bus_bar(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// What the exchange costs (2 x B200 over NVLink, b200conv_bus_trace device timestamps, profiles/experiments/
// n2_exchange_*.jsonl): push -> summed bus 0.5 us when blocks follow each other back to back, but 9.5-12 us when
// bench.py's L2 flush (512 MB streamed through each GPU) sits between blocks — symmetric on both ranks while their
// ready times differ by < 2 us, i.e. one-way latency of the first remote store after the flush, not rank skew.
// None of these moved it: weak / relaxed.sys / volatile stores, a system fence after the push, prefetching the
// receive slot into L2, a throw-away remote load at kernel start from one thread (from every CTA it cost +5 us).
// value i (0 <= i < n) of this rank's bus -> slot [epoch & 1][rank][i] of EVERY rank's buffer (own included)
__device__ __forceinline__ void bus_ll_push(const BusExchange& x, int n, int i, float v) {
    const unsigned long long word = (static_cast<unsigned long long>(x.epoch) << 32) | __float_as_uint(v);
    const size_t off = (static_cast<size_t>(x.epoch & 1u) * x.world + x.rank) * n + i;
#pragma unroll 1
    for (int p = 0; p < x.world; ++p)
        if (!(x.debug & 1u) || p == x.rank)
            asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(x.peers[p] + off), "l"(word) : "memory");
}
// two adjacent values (i even) in ONE 16-byte store per rank.  Each 8-byte half still carries its own epoch, so a
// reader that finds one half early and the other late simply keeps polling.  (Device timestamps at N = 2 after an L2
// flush: 11 us from push to summed bus with four 8-byte store instructions per thread and peer, 5.3 us with these two
// — the remote stores of a thread to one peer complete one after the other, ~2.6 us each when cold, 0.25 us warm.
// Spreading the final over NG CTAs and the two planes over different threads did not shorten it further and cost
// +2 us on one GPU: profiles/experiments/n2_exchange_*.jsonl.)
__device__ __forceinline__ void bus_ll_push2(const BusExchange& x, int n, int i, float v0, float v1) {
    const unsigned long long w0 = (static_cast<unsigned long long>(x.epoch) << 32) | __float_as_uint(v0);
    const unsigned long long w1 = (static_cast<unsigned long long>(x.epoch) << 32) | __float_as_uint(v1);
    const size_t off = (static_cast<size_t>(x.epoch & 1u) * x.world + x.rank) * n + i;
#pragma unroll 1  // (kernel size matters more than this loop: the L1.5 instruction cache is 32 KB)
    for (int p = 0; p < x.world; ++p)
        asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(x.peers[p] + off), "l"(w0), "l"(w1) : "memory");
}

// Sums over ranks, in rank order, of NV values of this thread (indices idx[j], ignored where !act[j]): polls this
// rank's own buffer until every rank's word carries the epoch.  All NV x 4 words of a group of four ranks are
// loaded together and checked together — the wait is one round trip per group of ranks, not one per word (the
// first version polled word after word: 16 dependent L2 round trips, ~10 us, at the end of the C2 step).
template <int NV>
__device__ __forceinline__ void bus_ll_gather(const BusExchange& x, int n, const int (&idx)[NV], const bool (&act)[NV],
                                              float (&acc)[NV]) {
    const unsigned long long* mine = x.peers[x.rank] + static_cast<size_t>(x.epoch & 1u) * x.world * n;
#pragma unroll
    for (int j = 0; j < NV; ++j) acc[j] = 0.0f;
    for (int q0 = 0; q0 < x.world; q0 += 4) {
        unsigned long long w[4][NV];
        unsigned spins = 0;
        for (;;) {
            bool ok = true;
#pragma unroll
            for (int dq = 0; dq < 4; ++dq) {
                const unsigned long long* src = mine + static_cast<size_t>(min(q0 + dq, x.world - 1)) * n;
#pragma unroll
                for (int j = 0; j < NV; ++j)
                    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w[dq][j]) : "l"(src + (act[j] ? idx[j] : 0)) : "memory");
            }
#pragma unroll
            for (int dq = 0; dq < 4; ++dq)
#pragma unroll
                for (int j = 0; j < NV; ++j)
                    if (q0 + dq < x.world && act[j] && static_cast<uint32_t>(w[dq][j] >> 32) != x.epoch &&
                        (!(x.debug & 2u) || q0 + dq == x.rank))
                        ok = false;
            if (ok) break;
            if (++spins > kBusSpinLimit) {
                *reinterpret_cast<volatile uint32_t*>(x.err) = 1u;  // mapped host memory: a plain store
                break;
            }
        }
#pragma unroll
        for (int dq = 0; dq < 4; ++dq)
            if (q0 + dq < x.world) {
#pragma unroll
                for (int j = 0; j < NV; ++j) acc[j] += __uint_as_float(static_cast<uint32_t>(w[dq][j]));
            }
    }
}
// one value
__device__ __forceinline__ float bus_ll_sum(const BusExchange& x, int n, int i) {
    const int idx[1] = {i};
    const bool act[1] = {true};
    float acc[1];
    bus_ll_gather<1>(x, n, idx, act, acc);
    return acc[0];
}

// Every level is "data, barrier, ONE thread fences and takes the ticket, barrier" — the barrier orders the other
// threads' stores before thread 0's fence (cumulativity; the pattern of cooperative-groups grid sync), which
// keeps 255 redundant membars off the critical path.  Loads of a level are issued in batches (all in flight,
// then added in the fixed order): the tail after the last track is a chain of L2 round trips, not of bytes.

// The chunk's final local bus {l0, r0, l1, r1} of this thread's NP column pairs -> mix, or over NVLink first.
// Thread `tid` owns the pairs tid + q * nthr (q < NP) of the chunk; a pair is active when 2 * pair < CH.
// kPart: 0 = everything; 1 = the push only (multi-GPU; the caller finishes later with kPart = 2, when the peers'
// values have had time to arrive — v is not needed then); 2 = gather, sum and write.
template <int NP, int kPart = 0>
__device__ __forceinline__ void bus_finish_chunk(const BusTreeParams& bt, int chunk, const float4 (&v)[NP], int tid, int nthr) {
    const int c0 = chunk * bt.CH;
    const int n = 2 * bt.B;
    if (kPart != 2 && bt.x.world > 1 && bt.x.trace && tid == 0 && chunk == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        bt.x.trace[(bt.x.epoch % kBusTraceLen) * 2] = now;
    }
    if (kPart != 2 && bt.x.world > 1) {
#pragma unroll
        for (int q = 0; q < NP; ++q) {  // all pushes first, then the polls: the NVLink latencies overlap
            const int pair = tid + q * nthr;
            if (2 * pair < bt.CH) {
                const int col = c0 + 2 * pair;
                bus_ll_push2(bt.x, n, col, v[q].x, v[q].z);          // (col is even and n = 2B is even: 16-byte aligned)
                bus_ll_push2(bt.x, n, bt.B + col, v[q].y, v[q].w);
            }
        }
    }
    if (kPart == 1) return;
    float sum[4 * NP];
    if (bt.x.world > 1) {  // fixed rank order: every rank computes the bit-identical sum
        int idx[4 * NP];
        bool act[4 * NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const int pair = tid + q * nthr;
            const int col = c0 + 2 * pair;
            idx[4 * q] = col; idx[4 * q + 1] = col + 1; idx[4 * q + 2] = bt.B + col; idx[4 * q + 3] = bt.B + col + 1;
            act[4 * q] = act[4 * q + 1] = act[4 * q + 2] = act[4 * q + 3] = (2 * pair < bt.CH);
        }
        bus_ll_gather<4 * NP>(bt.x, n, idx, act, sum);
        if (bt.x.trace && tid == 0 && chunk == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            bt.x.trace[(bt.x.epoch % kBusTraceLen) * 2 + 1] = now;
        }
    } else {
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            sum[4 * q] = v[q].x; sum[4 * q + 1] = v[q].z; sum[4 * q + 2] = v[q].y; sum[4 * q + 3] = v[q].w;
        }
    }
#pragma unroll
    for (int q = 0; q < NP; ++q) {
        const int pair = tid + q * nthr;
        if (2 * pair < bt.CH) {
            const int col = c0 + 2 * pair;
            *reinterpret_cast<float2*>(bt.mix + col) = make_float2(sum[4 * q], sum[4 * q + 1]);
            *reinterpret_cast<float2*>(bt.mix + bt.B + col) = make_float2(sum[4 * q + 2], sum[4 * q + 3]);
        }
    }
}

// Called by all `nthr` threads of the arriving group (tid = 0 .. nthr-1; they synchronise on hardware
// barrier `bar_id`) after they have written ybus[t][chunk*CH .. +CH).  `flag` is one int of shared memory.
// Thread `tid` owns the column pairs tid + q * nthr, q < NP; NP * nthr >= CH / 2 is required.
// Returns true (to all threads alike) only with kDefer on a multi-GPU job, in the one CTA that completed the chunk's
// local bus: its values are pushed to the peers, and the caller owes a bus_tree_finish for the chunk later.
template <int NP = 1, bool kDefer = false>
__device__ __forceinline__ bool bus_tree_arrive(const BusTreeParams& bt, int t, int chunk, int tid, int nthr,
                                                uint32_t bar_id, int* flag) {
    const int g = t / bt.G1;
    const int gsize = min(bt.G1, bt.T - g * bt.G1);
    const int c0 = chunk * bt.CH;
    // ---- level 1: last arriver of (group, chunk) ----
    bus_bar(bar_id, nthr);
    if (tid == 0) {
        int last = 1;
        if (gsize > 1) {
            __threadfence();
            unsigned* cnt = bt.gcount + g * bt.NC + chunk;
            last = (atomicAdd(cnt, 1u) == static_cast<unsigned>(gsize) - 1u);
            if (last) {
                *cnt = 0;  // re-armed for the next launch
                __threadfence();
            }
        }
        *flag = last;
    }
    bus_bar(bar_id, nthr);
    if (!*flag) return false;
    float4 part[NP];  // {l0, r0, l1, r1} per pair
    const float2* gn = reinterpret_cast<const float2*>(bt.gains) + g * bt.G1;
#pragma unroll
    for (int q = 0; q < NP; ++q) part[q] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    for (int t0 = 0; t0 < gsize; t0 += 16) {  // 16 rows (x NP pairs) in flight, then the adds in track order
        float2 v[NP][16];
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const int pair = tid + q * nthr;
            const float* row = bt.ybus + static_cast<size_t>(g) * bt.G1 * bt.B + c0 + 2 * pair;
#pragma unroll
            for (int j = 0; j < 16; ++j)
                v[q][j] = (t0 + j < gsize && 2 * pair < bt.CH)
                              ? __ldcg(reinterpret_cast<const float2*>(row + static_cast<size_t>(t0 + j) * bt.B))
                              : make_float2(0.0f, 0.0f);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (t0 + j < gsize) {
                const float2 gg = gn[t0 + j];
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    part[q].x = fmaf(gg.x, v[q][j].x, part[q].x);
                    part[q].y = fmaf(gg.y, v[q][j].x, part[q].y);
                    part[q].z = fmaf(gg.x, v[q][j].y, part[q].z);
                    part[q].w = fmaf(gg.y, v[q][j].y, part[q].w);
                }
            }
        }
    }
    if (bt.NG > 1) {
        const size_t hp = static_cast<size_t>(bt.B) >> 1;  // column pairs per bus row
        // ---- level 2: the last group of the chunk adds the NG partials in group order ----
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const int pair = tid + q * nthr;
            if (2 * pair < bt.CH) bt.gpart[static_cast<size_t>(g) * hp + (c0 >> 1) + pair] = part[q];
        }
        bus_bar(bar_id, nthr);
        if (tid == 0) {
            __threadfence();
            const int last = (atomicAdd(bt.ccount + chunk, 1u) == static_cast<unsigned>(bt.NG) - 1u);
            if (last) {
                bt.ccount[chunk] = 0;
                __threadfence();
            }
            *flag = last;
        }
        bus_bar(bar_id, nthr);
        if (!*flag) return false;
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const int pair = tid + q * nthr;
            part[q] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (2 * pair < bt.CH) {
                const float4* gp = bt.gpart + (c0 >> 1) + pair;
                for (int g0 = 0; g0 < bt.NG; g0 += 8) {
                    float4 v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        v[j] = (g0 + j < bt.NG) ? __ldcg(gp + static_cast<size_t>(g0 + j) * hp) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (g0 + j < bt.NG) {
                            part[q].x += v[j].x; part[q].y += v[j].y; part[q].z += v[j].z; part[q].w += v[j].w;
                        }
                    }
                }
            }
        }
    }
    if (kDefer && bt.x.world > 1) {
        bus_finish_chunk<NP, 1>(bt, chunk, part, tid, nthr);
        return true;
    }
    bus_finish_chunk<NP>(bt, chunk, part, tid, nthr);
    return false;
}
// the deferred half of bus_tree_arrive<NP, true>: poll the peers' values (long arrived by now), sum in rank order, write
template <int NP = 1>
__device__ __forceinline__ void bus_tree_finish(const BusTreeParams& bt, int chunk, int tid, int nthr) {
    float4 none[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) none[q] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    bus_finish_chunk<NP, 2>(bt, chunk, none, tid, nthr);
}

// ---- column-slice bus -------------------------------------------------------------------------------------------
// Called by the `nthr` = 128 threads that have just written ybus[t][n_off .. n_off + Bs) (hardware barrier `bar_id`;
// `part` is 4.5 KB of shared memory, `flag` one int).  CTA `t` owns the columns n_off + t * slice .. + slice.  Fixed
// summation order (row lanes in track order, then lane partials in lane order): the same on every rank.
// Returns true when the slice's multi-GPU values were pushed and bus_slice_finish is owed.
__device__ __forceinline__ bool bus_slice_reduce(const BusTreeParams& bt, const BusSlice& sl, int t, int n_off, int Bs,
                                                 int tid, uint32_t bar_id, float* part, int* flag) {
    constexpr int nthr = 128;
    bus_bar(bar_id, nthr);  // the row is written by all threads
    const int c_lo = t * sl.slice;
    const bool has_slice = c_lo < Bs;
    if (tid == 0) {
        __threadfence();
        atomicAdd(sl.arrive, 1ULL);
        if (has_slice) {
            unsigned spins = 0;
            unsigned long long seen;
            do {
                asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(seen) : "l"(sl.arrive) : "memory");
                if (++spins > kBusSpinLimit) {
                    *reinterpret_cast<volatile uint32_t*>(bt.x.err) = 1u;  // a track never arrived: report, do not hang
                    break;
                }
            } while (seen < sl.target);
            __threadfence();
        }
    }
    if (!has_slice) return false;
    bus_bar(bar_id, nthr);  // every track's row is in L2
    const int Q = sl.slice >> 2, RL = nthr / Q;  // column quads of the slice x row lanes
    const int rl = tid / Q, q = tid - rl * Q;
    const int col = n_off + c_lo + 4 * q;        // column in the caller's block
    const bool live = c_lo + 4 * q < Bs;
    float4 L = make_float4(0.f, 0.f, 0.f, 0.f), R = L;
    if (live) {
        const float* src = bt.ybus + col;
        const float2* gn = reinterpret_cast<const float2*>(bt.gains);
        for (int t0 = rl; t0 < bt.T; t0 += 4 * RL) {  // four rows in flight, added in track order
            float4 y[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                y[j] = (t0 + j * RL < bt.T) ? __ldcg(reinterpret_cast<const float4*>(src + static_cast<size_t>(t0 + j * RL) * bt.B))
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (t0 + j * RL < bt.T) {
                    const float2 g = gn[t0 + j * RL];
                    L.x = fmaf(g.x, y[j].x, L.x); L.y = fmaf(g.x, y[j].y, L.y); L.z = fmaf(g.x, y[j].z, L.z); L.w = fmaf(g.x, y[j].w, L.w);
                    R.x = fmaf(g.y, y[j].x, R.x); R.y = fmaf(g.y, y[j].y, R.y); R.z = fmaf(g.y, y[j].z, R.z); R.w = fmaf(g.y, y[j].w, R.w);
                }
            }
        }
    }
    float4* part4 = reinterpret_cast<float4*>(part);  // [RL][Q][2] float4
    part4[(rl * Q + q) * 2] = L;
    part4[(rl * Q + q) * 2 + 1] = R;
    bus_bar(bar_id, nthr);
    const int n = 2 * bt.B;
    const bool multi = bt.x.world > 1;
    if (multi && bt.x.trace && tid == 0 && t == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        bt.x.trace[(bt.x.epoch % kBusTraceLen) * 2] = now;
    }
    // The RL lane partials of each of the slice's NO = 8 Q bus values, in lane order — in two steps when there are more
    // threads than values (C2: 8 values, 128 lanes: 16 segments of 8 lanes, then the 16 segment sums), because a single
    // thread adding 128 shared-memory values one after the other is 2 us on the way to the bus.
    const int NO = 8 * Q;
    const int nseg = NO < nthr ? nthr / NO : 1, per = RL / nseg;
    float* part2 = part + 8 * nthr;  // [nseg][NO]
    auto emit = [&](int o, float sum) {
        const int fq = o >> 3, k = o & 7;
        if (c_lo + 4 * fq >= Bs) return;
        const int i = (k < 4 ? 0 : bt.B) + n_off + c_lo + 4 * fq + (k & 3);
        if (multi)
            bus_ll_push(bt.x, n, i, sum);
        else
            bt.mix[i] = sum;
    };
    if (nseg > 1) {
        const int o = tid % NO, sg = tid / NO;
        float sum = 0.0f;
        for (int r = sg * per; r < (sg + 1) * per; ++r) sum += part[(r * Q + (o >> 3)) * 8 + (o & 7)];
        part2[sg * NO + o] = sum;
        bus_bar(bar_id, nthr);
        if (tid < NO) {
            float tot = 0.0f;
            for (int g2 = 0; g2 < nseg; ++g2) tot += part2[g2 * NO + tid];
            emit(tid, tot);
        }
    } else {
        for (int o = tid; o < NO; o += nthr) {
            float sum = 0.0f;
            for (int r = 0; r < RL; ++r) sum += part[(r * Q + (o >> 3)) * 8 + (o & 7)];
            emit(o, sum);
        }
    }
    (void)flag;
    if (multi && (bt.x.debug & 4u)) {  // measurement only: collect right away instead of after the epilogue
        for (int f = tid; f < 8 * Q; f += nthr) {
            const int fq = f >> 3, k = f & 7;
            if (c_lo + 4 * fq >= Bs) continue;
            const int i = (k < 4 ? 0 : bt.B) + n_off + c_lo + 4 * fq + (k & 3);
            bt.mix[i] = bus_ll_sum(bt.x, n, i);
        }
        return false;
    }
    return multi;
}
// the owed half on a multi-GPU job: the peers' values of this CTA's slice (long arrived), summed in rank order
__device__ __forceinline__ void bus_slice_finish(const BusTreeParams& bt, const BusSlice& sl, int t, int n_off, int Bs, int tid) {
    const int c_lo = t * sl.slice, Q = sl.slice >> 2, n = 2 * bt.B;
    for (int f = tid; f < 8 * Q; f += 128) {
        const int fq = f >> 3, k = f & 7;
        if (c_lo + 4 * fq >= Bs) continue;
        const int i = (k < 4 ? 0 : bt.B) + n_off + c_lo + 4 * fq + (k & 3);
        bt.mix[i] = bus_ll_sum(bt.x, n, i);
    }
    if (bt.x.trace && tid == 0 && t == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        bt.x.trace[(bt.x.epoch % kBusTraceLen) * 2 + 1] = now;
    }
}
#endif  // __CUDACC__

}  // namespace b200conv
