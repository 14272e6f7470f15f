// common.cuh — shared device helpers for the sm_100a convolution kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <mutex>
#include <utility>

namespace b200conv {

constexpr int kWarp = 32;

// ---------------------------------------------------------------------------------------------
// Host: opt a kernel into more than 48 KB of dynamic shared memory.  The attribute belongs to the
// (function, DEVICE) pair, and one process may drive several devices from concurrent threads
// (group.cu), so the "already done" cache is keyed by both and guarded by a mutex — a process-wide
// flag would leave every device but the first without the opt-in.
// ---------------------------------------------------------------------------------------------
inline cudaError_t ensure_dyn_smem(const void* fn, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(mu);
    size_t& have = done[{fn, dev}];
    if (have >= bytes) return cudaSuccess;
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
    if (e == cudaSuccess) have = bytes;
    return e;
}

// Make `device` current for the lifetime of the guard and restore the caller's device afterwards:
// every ABI entry point that takes an engine runs on the engine's device without side effects on
// the calling thread's current device.
struct DeviceGuard {
    int prev = -1;
    cudaError_t status = cudaSuccess;
    explicit DeviceGuard(int device) {
        status = cudaGetDevice(&prev);
        if (status == cudaSuccess && prev != device) status = cudaSetDevice(device);
        else if (status == cudaSuccess) prev = -1;  // nothing to restore
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// ---------------------------------------------------------------------------------------------
// 16-byte-chunk swizzle used by the direct FIR for BOTH the tap table and the input-history ring.
// A "block" is 16 floats (64 B = 4 chunks); a lane of the FIR kernel reads whole blocks with four
// LDS.128, and neighbouring lanes read neighbouring blocks (64 B apart), which on a plain layout
// is a 4-way bank conflict.  XOR-ing the chunk's position inside its 128 B line with the block
// index (mod 8) makes any 8 consecutive blocks hit 8 distinct 16 B bank groups.  The permutation
// stays inside one 128 B line, so the data can be kept pre-swizzled in HBM and moved by plain
// 1-D bulk copies (TMA) with no per-element address work.
//   f : logical chunk index (float4 index);   returns the physical chunk index.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t swz_chunk(uint32_t f) { return f ^ ((f >> 2) & 7u); }

// Physical float index of logical float index n.
__host__ __device__ __forceinline__ uint32_t swz_float(uint32_t n) { return (swz_chunk(n >> 2) << 2) | (n & 3u); }

// ---------------------------------------------------------------------------------------------
// mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) wrappers.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// Make barrier initialisation visible to the async proxy before the first bulk copy.
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

// try_wait with a suspend-time hint: the warp sleeps in hardware (up to ~hint ns) instead of
// burning issue slots in a poll loop (ncu on the first FIR kernel showed 8 % of all issued
// instructions were SYNCS/BRA/YIELD from un-hinted polling).
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// Named barrier among a subset of the CTA's warps (id 1..15; 0 is __syncthreads).
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// the non-blocking half: these threads signal, the bar.sync side waits for `nthreads` arrivals in total
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// global -> shared bulk copy; bytes % 16 == 0, both addresses 16 B aligned; completion is
// signalled on `bar` as `bytes` transaction bytes.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// Programmatic dependent launch (PDL): the primary kernel allows its dependent to be scheduled
// early; the dependent blocks here until the primary grid has completed and flushed.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait_primary() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// 128-bit streaming loads/stores that do not allocate in L1 (data touched once per launch).
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}

}  // namespace b200conv
