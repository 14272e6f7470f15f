// group.cu — multi-GPU convolution in ONE process: the b200conv_group_* entry points.
//
// Tracks are independent, so a group is simply one engine per GPU owning a contiguous track range
// (IRs, bus gains and sample-major columns use the global track index — SURVEY.md §8e), one
// persistent host thread per device that submits that device's work (so 8 GPUs are fed in
// parallel, not 8 x launch latency in series), and ONE collective per block: the stereo bus
// all-reduce, which every member engine performs INSIDE its last convolution kernel (bus_tree.cuh)
// over peer-mapped buffers (cudaDeviceEnablePeerAccess + b200conv_attach_bus; every device stores
// into every other device's slot over NVLink).  A member's block is exactly the single-GPU host call,
// b200conv_process_host, on the member's slice of the caller's buffers: pinned buffers are read and
// written in place over PCIe by the kernels, pageable ones take the staged copies.
// The one-process-per-GPU variant of the same thing is bench.py + gpuaudiobench_b200/distributed.py.
#include "../../include/b200conv.h"

#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_group_error;

int gfail(int code, const std::string& msg) {
    g_group_error = msg;
    return code;
}

struct Job {
    const float* h_in = nullptr;
    float* h_out = nullptr;
    float* h_mix = nullptr;
    uint32_t flags = 0;
};

struct Member {
    int device = 0;
    int t0 = 0, t1 = 0;  // global track range
    b200conv_engine* engine = nullptr;
    float* bus_buf = nullptr;       // symmetric slot buffer of this device
    float* h_mix = nullptr;         // pinned [2][B]: every member takes part in the exchange and needs a place for
                                    // the (identical) result; member 0 writes the caller's buffer instead
    int rc = 0;
    std::string err;
    std::thread worker;
};

}  // namespace

struct b200conv_group {
    b200conv_config cfg{};  // tracks = total tracks
    int n = 0;
    std::vector<Member> members;
    std::vector<uint64_t> peer_ptrs;
    // worker coordination
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    uint64_t generation = 0;
    int pending = 0;
    bool stop = false;
    Job job;
};

namespace {

int member_submit(b200conv_group* g, Member& m, const Job& job) {
    const int B = static_cast<int>(g->cfg.block);
    const bool sample_major = g->cfg.out_layout == B200CONV_OUT_SAMPLE_MAJOR;
    if (cudaSetDevice(m.device) != cudaSuccess) return B200CONV_ERR_CUDA;
    float* out = nullptr;
    if (job.h_out)  // sample-major: the engine writes its column tile of the full [B][Tg] matrix
        out = sample_major ? job.h_out : job.h_out + static_cast<size_t>(m.t0) * B;
    float* mix = nullptr;
    if (job.h_mix) mix = (m.t0 == 0) ? job.h_mix : m.h_mix;  // every rank holds the identical bus; member 0 returns it
    return b200conv_process_host(m.engine, job.h_in + static_cast<size_t>(m.t0) * B, out, mix, job.flags);
}

void worker_loop(b200conv_group* g, int idx) {
    uint64_t seen = 0;
    for (;;) {
        Job job;
        {
            std::unique_lock<std::mutex> lk(g->mu);
            g->cv_go.wait(lk, [&] { return g->stop || g->generation != seen; });
            if (g->stop) return;
            seen = g->generation;
            job = g->job;
        }
        Member& m = g->members[idx];
        m.rc = member_submit(g, m, job);
        if (m.rc) m.err = b200conv_last_error();
        {
            std::lock_guard<std::mutex> lk(g->mu);
            if (--g->pending == 0) g->cv_done.notify_all();
        }
    }
}

// The calling thread's current device is left as it was found (the engine calls guard themselves; the few
// runtime calls made here directly are bracketed by this).
struct CallerDevice {
    int dev = -1;
    CallerDevice() { if (cudaGetDevice(&dev) != cudaSuccess) dev = -1; }
    ~CallerDevice() { if (dev >= 0) cudaSetDevice(dev); }
};

template <typename F> int for_each_member(b200conv_group* g, F&& fn) {
    CallerDevice keep;
    for (Member& m : g->members) {
        if (cudaSetDevice(m.device) != cudaSuccess) return gfail(B200CONV_ERR_CUDA, "cudaSetDevice failed");
        const int rc = fn(m);
        if (rc) return gfail(rc, b200conv_last_error());
    }
    return B200CONV_OK;
}

}  // namespace

extern "C" {

const char* b200conv_group_last_error(void) { return g_group_error.c_str(); }

int b200conv_group_create(const b200conv_config* cfg, int n_gpus, b200conv_group** out) {
    if (!cfg || !out || n_gpus < 1) return gfail(B200CONV_ERR_INVALID, "b200conv_group_create: bad argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < n_gpus)
        return gfail(B200CONV_ERR_NO_DEVICE, "b200conv_group_create: " + std::to_string(n_gpus) + " GPUs requested, " +
                                                 std::to_string(ndev) + " visible (no CPU fallback)");
    if (cfg->tracks < static_cast<uint32_t>(n_gpus)) return gfail(B200CONV_ERR_INVALID, "b200conv_group_create: fewer tracks than GPUs");
    CallerDevice keep;
    auto* g = new b200conv_group();
    g->cfg = *cfg;
    g->n = n_gpus;
    g->members.resize(n_gpus);
    const int Tg = static_cast<int>(cfg->tracks), B = static_cast<int>(cfg->block);
    auto bail = [&](int code, const std::string& msg) {
        const std::string keep = msg;
        b200conv_group_destroy(g);
        return gfail(code, keep);
    };
    for (int i = 0; i < n_gpus; ++i) {
        Member& m = g->members[i];
        m.device = i;
        m.t0 = static_cast<int>(static_cast<long long>(Tg) * i / n_gpus);
        m.t1 = static_cast<int>(static_cast<long long>(Tg) * (i + 1) / n_gpus);
        b200conv_config c = *cfg;
        c.device = i;
        c.tracks = static_cast<uint32_t>(m.t1 - m.t0);
        c.track_offset = static_cast<uint32_t>(m.t0);
        c.total_tracks = static_cast<uint32_t>(Tg);
        if (int rc = b200conv_create(&c, &m.engine)) return bail(rc, b200conv_last_error());
        cudaSetDevice(i);
        const size_t bus_bytes = b200conv_bus_buffer_bytes(n_gpus, 2 * B);
        if (cudaMalloc(&m.bus_buf, bus_bytes) != cudaSuccess ||
            cudaMallocHost(&m.h_mix, static_cast<size_t>(2) * B * sizeof(float)) != cudaSuccess)
            return bail(B200CONV_ERR_CUDA, "b200conv_group_create: allocation failed");
        cudaMemset(m.bus_buf, 0, bus_bytes);
        cudaDeviceSynchronize();
        g->peer_ptrs.push_back(reinterpret_cast<uint64_t>(m.bus_buf));
    }
    for (int i = 0; i < n_gpus; ++i) {  // every device may store into every other device's bus buffer
        cudaSetDevice(i);
        for (int j = 0; j < n_gpus; ++j) {
            if (i == j) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, i, j);
            if (!can) return bail(B200CONV_ERR_NO_DEVICE, "b200conv_group_create: no peer access between GPU " + std::to_string(i) +
                                                              " and " + std::to_string(j));
            const cudaError_t e = cudaDeviceEnablePeerAccess(j, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return bail(B200CONV_ERR_CUDA, "cudaDeviceEnablePeerAccess failed");
            cudaGetLastError();
        }
    }
    if (n_gpus > 1)
        for (int i = 0; i < n_gpus; ++i)
            if (int rc = b200conv_attach_bus(g->members[i].engine, g->peer_ptrs.data(), i, n_gpus)) return bail(rc, b200conv_last_error());
    for (int i = 0; i < n_gpus; ++i) g->members[i].worker = std::thread(worker_loop, g, i);
    *out = g;
    return B200CONV_OK;
}

void b200conv_group_destroy(b200conv_group* g) {
    if (!g) return;
    CallerDevice keep;
    {
        std::lock_guard<std::mutex> lk(g->mu);
        g->stop = true;
    }
    g->cv_go.notify_all();
    for (Member& m : g->members) {
        if (m.worker.joinable()) m.worker.join();
        cudaSetDevice(m.device);
        cudaDeviceSynchronize();
        b200conv_destroy(m.engine);
        cudaFree(m.bus_buf);
        if (m.h_mix) cudaFreeHost(m.h_mix);
    }
    delete g;
}

int b200conv_group_size(const b200conv_group* g) { return g ? g->n : 0; }

int b200conv_group_load_ir(b200conv_group* g, const float* host_ir) {
    if (!g || !host_ir) return gfail(B200CONV_ERR_INVALID, "b200conv_group_load_ir: null argument");
    const size_t L = g->cfg.ir_len;
    return for_each_member(g, [&](Member& m) { return b200conv_load_ir(m.engine, host_ir + static_cast<size_t>(m.t0) * L); });
}

int b200conv_group_prime_history(b200conv_group* g, const float* host_hist) {
    if (!g) return gfail(B200CONV_ERR_INVALID, "b200conv_group_prime_history: null group");
    const size_t H = g->cfg.ir_len - 1;
    return for_each_member(g, [&](Member& m) {
        return b200conv_prime_history(m.engine, host_hist ? host_hist + static_cast<size_t>(m.t0) * H : nullptr);
    });
}

int b200conv_group_reset(b200conv_group* g) {
    if (!g) return gfail(B200CONV_ERR_INVALID, "b200conv_group_reset: null group");
    return for_each_member(g, [&](Member& m) { return b200conv_reset(m.engine); });
}

// channel strip on every member: per-track arrays are sliced at the member's first track
int b200conv_group_set_strip(b200conv_group* g, const b200conv_strip* strip) {
    if (!g) return gfail(B200CONV_ERR_INVALID, "b200conv_group_set_strip: null group");
    return for_each_member(g, [&](Member& m) {
        if (!strip) return b200conv_set_strip(m.engine, nullptr);
        b200conv_strip part = *strip;
        if (part.gains) part.gains += m.t0;
        if (part.biquad && !(part.ops & B200CONV_STRIP_SHARED_COEFFS)) part.biquad += static_cast<size_t>(5) * m.t0;
        return b200conv_set_strip(m.engine, &part);
    });
}

int b200conv_group_strip_state(b200conv_group* g, float* host_state, int set) {
    if (!g || !host_state) return gfail(B200CONV_ERR_INVALID, "b200conv_group_strip_state: null argument");
    return for_each_member(g, [&](Member& m) { return b200conv_strip_state(m.engine, host_state + static_cast<size_t>(2) * m.t0, set); });
}

int b200conv_group_strip_stats(b200conv_group* g, float* host_stats) {
    if (!g || !host_stats) return gfail(B200CONV_ERR_INVALID, "b200conv_group_strip_stats: null argument");
    return for_each_member(g, [&](Member& m) { return b200conv_strip_stats(m.engine, host_stats + static_cast<size_t>(2) * m.t0); });
}

int b200conv_group_process_host(b200conv_group* g, const float* h_in, float* h_out, float* h_mix, uint32_t flags) {
    if (!g || !h_in) return gfail(B200CONV_ERR_INVALID, "b200conv_group_process_host: null argument");
    {
        std::lock_guard<std::mutex> lk(g->mu);
        g->job = Job{h_in, h_out, h_mix, flags};
        g->pending = g->n;
        g->generation += 1;
    }
    g->cv_go.notify_all();
    {
        std::unique_lock<std::mutex> lk(g->mu);
        g->cv_done.wait(lk, [&] { return g->pending == 0; });
    }
    for (Member& m : g->members)
        if (m.rc) return gfail(m.rc, "GPU " + std::to_string(m.device) + ": " + m.err);
    return B200CONV_OK;
}

}  // extern "C"
