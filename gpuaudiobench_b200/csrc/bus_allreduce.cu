// bus_allreduce.cu — one-shot all-reduce of the stereo mix bus over NVLink peer memory.
//
// The only collective of the path is the sum of the per-GPU bus float[2][B] (4 KiB at B = 512;
// SURVEY.md §8e).  At that size an NCCL all-reduce is pure latency (measured here: ~15 us at 2
// GPUs, ~40 us at 8 with ms-scale p99 spikes, against a 60-190 us convolution step), so the engine
// ships its own kernel over a symmetric (peer-mapped) buffer:
//   1. push : every rank stores its partial into slot [epoch&1][rank] of EVERY peer's buffer
//             (plain P2P stores through NVLink / NVSwitch),
//   2. signal: __threadfence_system, then a release store of `epoch` into the peer's flag
//             [epoch&1][rank],
//   3. wait : acquire-poll the own flags until all `world` of them carry `epoch` (bounded spin),
//   4. sum  : add the `world` slots in rank order -> every rank gets the bit-identical bus.
// Two slots (epoch parity) are enough: a rank cannot finish epoch e+1 before every peer has
// signalled e+1, which a peer only does after its epoch-e kernel has completed.
// `out` may alias `local` (element i is read in step 1 and written in step 4 by the same thread).
// Buffer layout per rank (floats unless noted): data[2][world][n] | flags uint32 [2][world][kBusMaxChunks]
// (+pad) — shared with the in-kernel exchange of bus_tree.cuh, which signals per column chunk; this
// stand-alone kernel (used by the paths whose last kernel cannot carry the bus tree, and by callers
// that reduce a bus of their own) uses chunk 0's flag.
#include "../../include/b200conv.h"

#include <cuda_runtime.h>
#include <stdint.h>

#include "bus_tree.cuh"

namespace {

using b200conv::kBusMaxChunks;
constexpr int kMaxWorld = b200conv::kBusMaxWorld;
constexpr unsigned kSpinLimit = b200conv::kBusSpinLimit;  // ~seconds; then give up instead of hanging the GPU

struct PeerTable {
    float* buf[kMaxWorld];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(1024) bus_allreduce_kernel(PeerTable peers, const float* local,
                                                             float* out, int rank, int world, int n,
                                                             uint32_t epoch, uint32_t* __restrict__ error_flag) {
    const int slot = epoch & 1u;
    const size_t data_floats = static_cast<size_t>(2) * world * n;
    asm volatile("griddepcontrol.wait;" ::: "memory");  // PDL: `local` is written by the kernel launched before us
    // 1. push my partial to every rank (including myself)
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = local[i];
        for (int p = 0; p < world; ++p) peers.buf[p][(static_cast<size_t>(slot) * world + rank) * n + i] = v;
    }
    __threadfence_system();
    __syncthreads();
    // 2. signal, 3. wait
    if (threadIdx.x < world) {
        uint32_t* peer_flags = reinterpret_cast<uint32_t*>(peers.buf[threadIdx.x] + data_floats);
        st_release_sys(peer_flags + (slot * world + rank) * kBusMaxChunks, epoch);
        const uint32_t* my_flags = reinterpret_cast<const uint32_t*>(peers.buf[rank] + data_floats);
        unsigned spins = 0;
        while (ld_acquire_sys(my_flags + (slot * world + threadIdx.x) * kBusMaxChunks) != epoch) {
            if (++spins > kSpinLimit) {
                *reinterpret_cast<volatile uint32_t*>(error_flag) = 1u;  // may be mapped host memory: a plain store
                break;
            }
        }
    }
    __syncthreads();
    // 4. fixed-order sum of the slots that landed in my buffer
    const float* mine = peers.buf[rank] + static_cast<size_t>(slot) * world * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float acc = 0.0f;
        for (int q = 0; q < world; ++q) acc += __ldcg(mine + static_cast<size_t>(q) * n + i);
        out[i] = acc;
    }
}

}  // namespace

extern "C" size_t b200conv_bus_buffer_bytes(int world, int n) {
    const size_t data = static_cast<size_t>(2) * world * n * sizeof(float);
    const size_t flags = static_cast<size_t>(2) * world * kBusMaxChunks * sizeof(uint32_t);
    return (data + flags + 255) / 256 * 256;
}

extern "C" int b200conv_bus_allreduce(const float* d_local, float* d_out, const uint64_t* peer_buffers, int rank,
                                      int world, int n, uint32_t epoch, uint32_t* d_error_flag, void* stream) {
    if (!d_local || !d_out || !peer_buffers || !d_error_flag || world < 1 || world > kMaxWorld || rank < 0 ||
        rank >= world || n < 1 || epoch == 0)
        return B200CONV_ERR_INVALID;
    PeerTable t{};
    for (int p = 0; p < world; ++p) t.buf[p] = reinterpret_cast<float*>(peer_buffers[p]);
    const int threads = n >= 1024 ? 1024 : ((n + 31) / 32 * 32 < 32 ? 32 : (n + 31) / 32 * 32);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(threads);
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // hide our launch under the primary's tail
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const cudaError_t err = cudaLaunchKernelEx(&cfg, bus_allreduce_kernel, t, d_local, d_out, rank, world, n, epoch, d_error_flag);
    return err == cudaSuccess ? B200CONV_OK : B200CONV_ERR_CUDA;
}
