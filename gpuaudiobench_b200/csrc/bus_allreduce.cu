// bus_allreduce.cu — stand-alone one-shot all-reduce of a stereo mix bus over NVLink peer memory.
//
// The only collective of the path is the sum of the per-GPU bus float[2][B] (4 KiB at B = 512;
// SURVEY.md §8e).  At that size an NCCL all-reduce is pure latency (measured: ~15 us at 2 GPUs, ~40 us
// at 8 with ms-scale p99 spikes, against a 40-190 us convolution step), so the engines do it themselves,
// INSIDE their last kernel (bus_tree.cuh: the FIR / tensor-core kernels' bus tree, the UPOLS bus kernel).
// This file is the same exchange as a kernel of its own, for callers that reduce a bus they own:
// every thread pushes its values into its slot of every rank's symmetric buffer as 8-byte
// (epoch, value) words and polls its own values from all ranks — one NVLink one-way latency, no fence,
// no flag, fixed rank order -> the bit-identical sum on every rank.  `out` may alias `local`.
// Buffer layout per rank: uint64 ll[2][world][n] (b200conv_bus_buffer_bytes), zero-initialised once.
#include "../../include/b200conv.h"

#include <cuda_runtime.h>
#include <stdint.h>

#include "bus_tree.cuh"

namespace {

using b200conv::BusExchange;

__global__ void __launch_bounds__(1024) bus_allreduce_kernel(BusExchange x, const float* local, float* out, int n) {
    asm volatile("griddepcontrol.wait;" ::: "memory");  // PDL: `local` is written by the kernel launched before us
    for (int i = threadIdx.x; i < n; i += blockDim.x) b200conv::bus_ll_push(x, n, i, local[i]);
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = b200conv::bus_ll_sum(x, n, i);
}

}  // namespace

extern "C" size_t b200conv_bus_buffer_bytes(int world, int n) {
    const size_t bytes = static_cast<size_t>(2) * world * n * sizeof(unsigned long long);
    return (bytes + 255) / 256 * 256;
}

extern "C" int b200conv_bus_allreduce(const float* d_local, float* d_out, const uint64_t* peer_buffers, int rank,
                                      int world, int n, uint32_t epoch, uint32_t* d_error_flag, void* stream) {
    if (!d_local || !d_out || !peer_buffers || !d_error_flag || world < 1 || world > b200conv::kBusMaxWorld || rank < 0 ||
        rank >= world || n < 1 || epoch == 0)
        return B200CONV_ERR_INVALID;
    BusExchange x{};
    for (int p = 0; p < world; ++p) x.peers[p] = reinterpret_cast<unsigned long long*>(peer_buffers[p]);
    x.rank = rank;
    x.world = world;
    x.epoch = epoch;
    x.err = d_error_flag;
    x.trace = nullptr;
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
    const cudaError_t err = cudaLaunchKernelEx(&cfg, bus_allreduce_kernel, x, d_local, d_out, n);
    return err == cudaSuccess ? B200CONV_OK : B200CONV_ERR_CUDA;
}
