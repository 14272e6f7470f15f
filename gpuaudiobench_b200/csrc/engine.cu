// engine.cu — libb200conv.so: the C ABI of include/b200conv.h over the sm_100a kernels.
//
// Engine-owned device state:
//   direct : taps  h[T][Lc*16]   (zero padded, chunk-swizzled)       — replaces the reference's
//            ring  x[T][cap]     (input history, chunk-swizzled)        cudaArray/texture, bench_conv1d.cu:123-157
//            part  [S][T][B]     (tap-split partial sums)
//   UPOLS  : H[T][P][B] float2   (partition spectra, packed bins, /N) — replaces d_ir_fft,
//            X[T][P][B] float2   (frequency-domain delay line ring)     bench_conv1d_accel.cu:175-228
//            prev[T][B], Ypart[S][T][B] float2, twiddle tables
#include "../../include/b200conv.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "bus_tree.cuh"
#include "common.cuh"
#include "direct_fir.cuh"
#include "strip.cuh"
#include "tc_toeplitz.cuh"
#include "upols.cuh"

using namespace b200conv;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

#define CU_TRY(call)                                                                              \
    do {                                                                                          \
        cudaError_t _e = (call);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            return fail(B200CONV_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));  \
    } while (0)

// Run the rest of the entry point on the engine's device and give the caller's device back on return
// (b200conv.h: "every call that takes an engine runs on cfg.device and leaves the calling thread's
// current device as it found it").
#define ENGINE_DEVICE(dev)                                                                              \
    DeviceGuard _device_guard(dev);                                                                     \
    if (_device_guard.status != cudaSuccess)                                                            \
        return fail(B200CONV_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(_device_guard.status))

bool is_pow2(uint32_t v) { return v && !(v & (v - 1)); }
int ilog2(uint32_t v) {
    int l = 0;
    while ((1u << l) < v) ++l;
    return l;
}

int env_int(const char* name, int dflt) {
    const char* s = std::getenv(name);
    return (s && *s) ? std::atoi(s) : dflt;
}

struct DirectState {
    int A = 0, CL = 0, SPS = 0, JSb = 0, NS = 0, G = 0, MS = 1, Lc = 0, cap = 0, nbuf = 0, xtile_blocks = 0, ntiles = 0;
    size_t smem = 0;
    float* h = nullptr;
    float* ring = nullptr;
    float* partial = nullptr;
    int pos = 0;
};

struct TcState {
    TcGeometry g{};
    float* bimg = nullptr;   // [T][NGRP][2][32][R][4]
    float* hhead = nullptr;  // [T][Bs] first Bs taps, fp32
    unsigned long long* trace = nullptr;  // B200CONV_TC_TRACE=1: [grid][kTcTraceSlots] phase stamps of the last launch
    float* pend = nullptr;   // [T][capP]
    float* xprev = nullptr;  // [2][T][128] ping-pong
    int ppos = 0, xpar = 0;
    int grid = 0;
    int Bs = 0, nsub = 1;    // block of one launch; caller's block = nsub * Bs (buffers > 1024 samples stream through in sub-blocks)
    int nch = 1;             // bus chunks per launch (<= 512 columns each)
    float* shadow = nullptr; // [T][capP] + [T][128]: state saved around a PEEK of more than one sub-block
    unsigned long long* arrive = nullptr;  // column-slice bus: arrival counter, launch `seq` waits for T * seq
    unsigned long long seq = 0;
    int slice = 0;           // columns per CTA of the slice bus; 0: the ticket tree (more tracks than CTAs of one wave)
};

struct UpolsState {
    int P = 0, M = 0, logM = 0, S = 1;
    float2* H = nullptr;
    float2* X = nullptr;
    float2* Ypart = nullptr;
    float* prev = nullptr;        // [2][T][B] previous buffer, ping-pong (fused kernel: read [par], write [par ^ 1])
    int par = 0;
    unsigned* counters = nullptr;
    unsigned* xready = nullptr;   // [T] fused kernel with bin tiles: "X_m of launch `seq` is in the ring"
    unsigned seq = 0;
    bool fused = false;
};

}  // namespace

struct b200conv_engine {
    b200conv_config cfg{};
    uint32_t impl = 0;  // the engine actually built: cfg.algo, or DIRECT_TC when the planner dispatches ALGO_DIRECT to it
    int T = 0, B = 0, L = 0, Tg = 0, toff = 0;
    int sm_count = 0;
    bool ir_loaded = false;
    uint64_t blocks = 0, launches = 0, device_bytes = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t last_stream = nullptr;  // stream of the latest process(): state (ring, delay line) is ordered on it
    bool has_last_stream = false;
    float* d_in_stage = nullptr;   // staging slot 0 (slot 1 below: b200conv_submit double-buffers)
    float* d_out_stage = nullptr;
    float* d_mix_stage = nullptr;
    float* d_in_stage1 = nullptr;
    float* d_out_stage1 = nullptr;
    float* d_mix_stage1 = nullptr;
    cudaStream_t copy_stream = nullptr;              // device -> host copies of block m run here, under block m+1's kernels
    cudaEvent_t ev_done[2] = {nullptr, nullptr};     // kernels of the slot's block have finished
    cudaEvent_t ev_out[2] = {nullptr, nullptr};      // the slot's results are on the host
    uint64_t submitted = 0, completed = 0;           // tickets handed out / waited for
    bool slot_has_bus[2] = {false, false};
    float* d_gains = nullptr;
    // stereo-bus tree (bus_tree.cuh): scratch rows, group partials, tickets
    float* d_ybus = nullptr;      // [T][B]
    float4* d_gpart = nullptr;    // [NG][B/2] {l0, r0, l1, r1}
    unsigned* d_gcount = nullptr; // [NG][NC]
    unsigned* d_ccount = nullptr; // [NC]
    int bus_G1 = 1, bus_NG = 1, bus_CH = 0, bus_NC = 1;
    // multi-GPU bus group (b200conv_attach_bus): world == 1 means a stand-alone engine
    int bus_world = 1, bus_rank = 0;
    uint64_t bus_peers[kBusMaxWorld] = {};
    uint32_t bus_epoch = 0;
    unsigned long long* d_bus_trace = nullptr;  // B200CONV_BUS_TRACE=1: per-epoch (ready, done) %globaltimer stamps
    uint32_t* d_bus_err = nullptr;  // pinned, mapped host word (UVA): kernels store 1 on a spin timeout, the host
                                    // reads it after its stream synchronise without a copy
    DirectState dir;
    UpolsState up;
    TcState tc;
    bool bus_in_kernel_single = false;  // B200CONV_BUS_TREE=1: run the in-kernel bus tree on a stand-alone UPOLS engine too
    // channel strip (b200conv_set_strip): device copies of the per-track parameters
    uint32_t strip_ops = 0;
    float strip_gain = 1.0f;
    float* d_strip_gains = nullptr;  // [T] or null
    float* d_strip_coef = nullptr;   // [T][5] or [5]
    bool strip_shared_coef = false;
    bool strip_use_gains = false;
    float* d_strip_state = nullptr;  // [T][2]
    float* d_strip_stats = nullptr;  // [T][2]
    bool profiling = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    float stage_ms[4] = {0, 0, 0, 0};
    uint32_t stage_calls = 0;
    std::vector<void*> allocs;
};

namespace {

template <typename Tp>
int dev_alloc(b200conv_engine* e, Tp** out, size_t count, bool zero = true) {
    void* p = nullptr;
    size_t bytes = count * sizeof(Tp);
    cudaError_t err = cudaMalloc(&p, bytes);
    if (err != cudaSuccess)
        return fail(B200CONV_ERR_CUDA, "cudaMalloc(" + std::to_string(bytes) + " bytes): " + cudaGetErrorString(err));
    if (zero) {
        err = cudaMemset(p, 0, bytes);
        if (err != cudaSuccess) return fail(B200CONV_ERR_CUDA, std::string("cudaMemset: ") + cudaGetErrorString(err));
    }
    e->allocs.push_back(p);
    e->device_bytes += bytes;
    *out = static_cast<Tp*>(p);
    return B200CONV_OK;
}

// ALGO_DIRECT is a request for the direct-form sum; the planner picks the kernel.  The tensor-core variant
// (tc_toeplitz.cu) wins once the job amortises its fixed cost (launch, prologue, FP32 own samples, TMEM epilogue):
// measured at C2 (128 x 512 x 16384, 1.07e9 MAC) 24.6 us against 54.3 us for the FFMA kernel; below ~2.5e8 MAC
// per buffer the FFMA kernel's shorter fixed path is faster.  Blocks longer than 1024 samples stream through the
// kernel in sub-blocks (B = 4096: 125 us against 345 us).  cfg.flags & B200CONV_FLAG_FFMA_ONLY or
// B200CONV_DIRECT_TC=0 keep the FFMA kernel (bit-exact impulse behaviour, 126-131 dB instead of ~110 dB).
uint32_t resolve_impl(const b200conv_config& cfg) {
    if (cfg.algo != B200CONV_ALGO_DIRECT) return cfg.algo;
    if ((cfg.flags & B200CONV_FLAG_FFMA_ONLY) || env_int("B200CONV_DIRECT_TC", 1) == 0) return B200CONV_ALGO_DIRECT;
    // (B = 128, one row block per item: 22.5 us against 25.7 us for the FFMA kernel at 128 tracks x 16384 taps)
    const bool shape_ok = cfg.block % kTcRows == 0 && cfg.block >= kTcRows &&
                          (cfg.block <= kTcRows * kTcMaxA || cfg.block % 512 == 0) && cfg.block <= 512 * kBusMaxChunks;
    const double macs = static_cast<double>(cfg.tracks) * cfg.block * cfg.ir_len;
    return (shape_ok && macs >= 2.5e8) ? B200CONV_ALGO_DIRECT_TC : B200CONV_ALGO_DIRECT;
}

int plan_direct(b200conv_engine* e) {
    DirectState& d = e->dir;
    const int B = e->B, L = e->L;
    if (B >= 512) {
        if (B % 512) return fail(B200CONV_ERR_INVALID, "direct engine: block must be 32,64,128,256 or a multiple of 512");
        d.A = 32;
        d.ntiles = B / 512;
    } else {
        if (!(B == 32 || B == 64 || B == 128 || B == 256))
            return fail(B200CONV_ERR_INVALID, "direct engine: block must be 32,64,128,256 or a multiple of 512");
        d.A = B / 16;
        d.ntiles = 1;
    }
    d.CL = 32 / d.A;
    // steps per lane per stage: even; for A = 4 / 2 the tap groups of a quarter-warp must land on
    // distinct 16 B bank groups, which needs SPS = 4 (mod 8) / 2 (mod 4) (see common.cuh swizzle)
    d.SPS = (d.A >= 16) ? 8 : (d.A == 2 ? 2 : 4);
    int sps = env_int("B200CONV_DIRECT_SPS", 0);
    if (sps > 0 && sps % 2 == 0 && d.A >= 8) d.SPS = sps;
    d.JSb = kFirWarps * d.CL * d.SPS;
    const int Lc0 = (L + 15) / 16;
    d.NS = (Lc0 + d.JSb - 1) / d.JSb;
    d.Lc = d.NS * d.JSb;
    const long long units = static_cast<long long>(e->T) * d.ntiles * d.NS;
    if (units > 0x7fffffffLL) return fail(B200CONV_ERR_INVALID, "direct engine: job too large");
    int per_sm = env_int("B200CONV_DIRECT_CTAS_PER_SM", kFirCtasPerSm);
    per_sm = std::max(1, std::min(per_sm, kFirCtasPerSm));
    d.G = static_cast<int>(std::min<long long>(units, static_cast<long long>(per_sm) * e->sm_count));
    d.MS = fir_max_segments(e->T * d.ntiles, d.NS, d.G);
    d.xtile_blocks = (d.A + d.JSb + 8 + 7) & ~7;
    const size_t stage_bytes = static_cast<size_t>(d.JSb + d.xtile_blocks) * 64;
    const size_t red_bytes = static_cast<size_t>(kFirWarps) * d.A * 16 * sizeof(float);
    const int per_cta = static_cast<int>((units + d.G - 1) / d.G);
    const int depth = std::max(1, std::min(env_int("B200CONV_DIRECT_NBUF", 4), kFirMaxStages));
    d.nbuf = std::min({per_cta, depth, kFirMaxStages});
    while (d.nbuf > 1 && 128 + d.nbuf * stage_bytes + red_bytes > kFirMaxSmem) --d.nbuf;
    d.smem = 128 + d.nbuf * stage_bytes + red_bytes;
    if (d.smem > kFirMaxSmem) return fail(B200CONV_ERR_INVALID, "direct engine: stage does not fit shared memory");
    const long long need = static_cast<long long>(d.Lc) * 16 + B + 128;
    const long long unit = std::lcm<long long>(B, 128);
    d.cap = static_cast<int>((need + unit - 1) / unit * unit);
    return B200CONV_OK;
}

int plan_upols(b200conv_engine* e) {
    UpolsState& u = e->up;
    const int B = e->B;
    if (!is_pow2(B) || B < 16 || B > 8192)
        return fail(B200CONV_ERR_INVALID, "UPOLS engine: block must be a power of two in [16, 8192]");
    u.M = B;
    u.logM = ilog2(B);
    u.P = (e->L + B - 1) / B;
    const int KT = std::max(1, (B / 2) / 256);
    u.fused = (u.M <= kFusedMaxM) && env_int("B200CONV_UPOLS_FUSED", 1) != 0;
    int S = env_int("B200CONV_UPOLS_SPLIT", 0);
    int min_parts = 1;
    if (S <= 0) {
        const long long base = static_cast<long long>(e->T) * KT;
        if (u.fused) {
            // The fused kernel runs 4 CTAs per SM (upols.cu) and is fastest when the grid needs about two
            // waves: every CTA starts and ends with a transform phase that moves no HBM bytes, and in a
            // single resident wave those phases line up across the whole GPU.  Measured on the C4 shard
            // (512 tracks): S = 1 128.3 us, S = 2 124.0 us, S = 3 122.5 us; C3 (1024 tracks): S = 1 160.9 us,
            // S = 2 161.7 us.  A split costs a partial-spectrum round trip, so keep >= 16 partitions each.
            S = static_cast<int>((6LL * e->sm_count + base - 1) / base);
            min_parts = 16;
        } else {
            // three-kernel path (M > 512): split only when the tracks alone cannot put ~2 CTAs on every SM
            S = static_cast<int>((2LL * e->sm_count + base - 1) / base);
        }
    }
    u.S = std::max(1, std::min({S, std::max(1, u.P / min_parts), 32}));
    return B200CONV_OK;
}

// The block of one tensor-core launch: the caller's block up to 1024 samples, else the largest of 1024 / 512 that
// divides it — the engine's state is a stream state, so a longer buffer is the same as several shorter ones.
int tc_sub_block(int B) {
    if (B <= kTcRows * kTcMaxA) return B;
    if (B % 1024 == 0) return 1024;
    if (B % 512 == 0) return 512;
    return 0;
}

int plan_tc(b200conv_engine* e) {
    const int B = e->B;
    TcState& c = e->tc;
    c.Bs = (B % kTcRows == 0 && B >= kTcRows) ? tc_sub_block(B) : 0;
    if (!c.Bs || B > 512 * kBusMaxChunks)
        return fail(B200CONV_ERR_INVALID, "tensor-core direct engine: block must be a multiple of 128 up to 1024, or a multiple of 512 up to 8192");
    c.nsub = B / c.Bs;
    c.nch = (c.Bs + 511) / 512;
    if (c.Bs % c.nch || (c.Bs / c.nch) % 2)
        return fail(B200CONV_ERR_INVALID, "tensor-core direct engine: block does not split into even bus chunks");
    c.g = tc_geometry(c.Bs, e->L);
    if (c.g.smem_bytes > 227 * 1024) return fail(B200CONV_ERR_INVALID, "tensor-core direct engine: tile does not fit shared memory");
    c.grid = std::max(1, std::min(e->T * c.g.NGRP, e->sm_count));  // persistent over (column group, track) items, one CTA per SM
    // The column-slice bus needs every track on its own CTA of the first wave (items are ordered group 0 first, so
    // CTA t < T owns track t): all of them are resident at once and can wait for each other.
    c.slice = 0;
    if (e->T <= c.grid && env_int("B200CONV_BUS_SLICE", 1)) {
        int sl = 4;
        while (sl * e->T < c.Bs) sl *= 2;
        if (sl <= 512) c.slice = sl;
    }
    return B200CONV_OK;
}

int set_default_gains(b200conv_engine* e) {
    std::vector<float> g(static_cast<size_t>(e->T) * 2);
    const double half_pi = 1.5707963267948966192313216916398;
    const double scale = 1.0 / std::sqrt(static_cast<double>(e->Tg));
    for (int t = 0; t < e->T; ++t) {
        double theta = (e->toff + t + 0.5) / e->Tg * half_pi;
        g[2 * t] = static_cast<float>(std::cos(theta) * scale);
        g[2 * t + 1] = static_cast<float>(std::sin(theta) * scale);
    }
    CU_TRY(cudaMemcpy(e->d_gains, g.data(), g.size() * sizeof(float), cudaMemcpyHostToDevice));
    CU_TRY(cudaDeviceSynchronize());
    return B200CONV_OK;
}

int reset_state(b200conv_engine* e) {
    // engine streams are non-blocking: fence explicitly around the (legacy-stream) memsets
    CU_TRY(cudaDeviceSynchronize());
    if (e->impl == B200CONV_ALGO_DIRECT) {
        CU_TRY(cudaMemset(e->dir.ring, 0, static_cast<size_t>(e->T) * e->dir.cap * sizeof(float)));
        e->dir.pos = 0;
    } else if (e->impl == B200CONV_ALGO_DIRECT_TC) {
        CU_TRY(cudaMemset(e->tc.pend, 0, static_cast<size_t>(e->T) * e->tc.g.capP * sizeof(float)));
        CU_TRY(cudaMemset(e->tc.xprev, 0, static_cast<size_t>(2) * e->T * 128 * sizeof(float)));
        e->tc.ppos = 0;
        e->tc.xpar = 0;
        CU_TRY(cudaMemset(e->tc.arrive, 0, sizeof(unsigned long long)));
        e->tc.seq = 0;
    } else {
        CU_TRY(cudaMemset(e->up.X, 0, static_cast<size_t>(e->T) * e->up.P * e->up.M * sizeof(float2)));
        CU_TRY(cudaMemset(e->up.prev, 0, static_cast<size_t>(2) * e->T * e->B * sizeof(float)));
        e->up.par = 0;
    }
    if (e->d_strip_state) CU_TRY(cudaMemset(e->d_strip_state, 0, static_cast<size_t>(2) * e->T * sizeof(float)));
    CU_TRY(cudaDeviceSynchronize());
    e->blocks = 0;
    return B200CONV_OK;
}

// convolution output -> strip, in place, on the launch stream (PDL: waits for the producing kernel)
StripParams strip_params(b200conv_engine* e, float* d_out, bool commit);

int run_strip(b200conv_engine* e, float* d_out, bool commit, cudaStream_t st) {
    const StripParams sp = strip_params(e, d_out, commit);
    CU_TRY(launch_strip(sp, st));
    e->launches += 1;
    return B200CONV_OK;
}

StripParams strip_params(b200conv_engine* e, float* d_out, bool commit) {
    StripParams sp{};
    sp.in = d_out;
    sp.out = d_out;
    sp.T = e->T;
    sp.B = e->B;
    sp.sample_major = (e->cfg.out_layout == B200CONV_OUT_SAMPLE_MAJOR);
    sp.ld = e->Tg;
    sp.col0 = e->toff;
    sp.ops = e->strip_ops;
    sp.gain = e->strip_gain;
    sp.gains = e->strip_use_gains ? e->d_strip_gains : nullptr;
    sp.coef = e->d_strip_coef;
    sp.shared_coef = e->strip_shared_coef ? 1 : 0;
    sp.state = e->d_strip_state;
    sp.stats = e->d_strip_stats;
    sp.peek = commit ? 0 : 1;
    return sp;
}

// Who the peers are for this block's bus exchange (world == 1: none).  The epoch was advanced by the caller.
BusExchange bus_exchange(const b200conv_engine* e) {
    BusExchange x{};
    x.rank = e->bus_rank;
    x.world = e->bus_world;
    for (int i = 0; i < e->bus_world; ++i) x.peers[i] = reinterpret_cast<unsigned long long*>(e->bus_peers[i]);
    x.epoch = e->bus_epoch;
    x.err = e->d_bus_err;
    x.trace = e->d_bus_trace;
    x.debug = static_cast<uint32_t>(env_int("B200CONV_BUS_DEBUG", 0));
    return x;
}

// The bus tree of this block (bus_tree.cuh).  d_mix == null disables it.  On a multi-GPU job the epoch was
// advanced by the caller (once per block with a bus, in lockstep on every rank).
BusTreeParams bus_params(const b200conv_engine* e, float* d_mix) {
    BusTreeParams b{};
    b.mix = d_mix;
    if (!d_mix) return b;
    b.gains = e->d_gains;
    b.ybus = e->d_ybus;
    b.gpart = e->d_gpart;
    b.gcount = e->d_gcount;
    b.ccount = e->d_ccount;
    b.T = e->T;
    b.B = e->B;
    b.G1 = e->bus_G1;
    b.NG = e->bus_NG;
    b.CH = e->bus_CH;
    b.NC = e->bus_NC;
    b.x = bus_exchange(e);
    return b;
}

struct StageTimer {
    b200conv_engine* e;
    cudaStream_t st;
    int idx = 0;
    void mark() {
        if (e->profiling && idx < 4) cudaEventRecord(e->ev[idx++], st);
    }
};

int finish_profile(b200conv_engine* e, int marks) {
    if (!e->profiling) return B200CONV_OK;
    CU_TRY(cudaEventSynchronize(e->ev[marks - 1]));
    for (int i = 0; i + 1 < marks; ++i) {
        float ms = 0.0f;
        CU_TRY(cudaEventElapsedTime(&ms, e->ev[i], e->ev[i + 1]));
        e->stage_ms[i] += ms;
    }
    e->stage_calls += 1;
    return B200CONV_OK;
}

}  // namespace

extern "C" {

uint32_t b200conv_abi_version(void) { return B200CONV_ABI_VERSION; }

const char* b200conv_last_error(void) { return g_last_error.c_str(); }

int b200conv_plan(const b200conv_config* cfg, int sm_count, int32_t plan[16]) {
    if (!cfg || !plan || sm_count <= 0) return fail(B200CONV_ERR_INVALID, "b200conv_plan: bad argument");
    if (cfg->tracks == 0 || cfg->block == 0 || cfg->ir_len == 0)
        return fail(B200CONV_ERR_INVALID, "b200conv_plan: tracks, block and ir_len must be > 0");
    b200conv_engine tmp;
    tmp.cfg = *cfg;
    tmp.T = static_cast<int>(cfg->tracks);
    tmp.B = static_cast<int>(cfg->block);
    tmp.L = static_cast<int>(cfg->ir_len);
    tmp.sm_count = sm_count;
    std::fill(plan, plan + 16, 0);
    const uint32_t impl = resolve_impl(*cfg);
    if (impl == B200CONV_ALGO_DIRECT) {
        int rc = plan_direct(&tmp);
        if (rc) return rc;
        const DirectState& d = tmp.dir;
        const int32_t v[13] = {d.A, d.CL, d.SPS, d.JSb, d.NS, d.G, d.Lc, d.cap, d.nbuf, d.xtile_blocks, d.ntiles,
                               static_cast<int32_t>(d.smem), d.MS};
        std::copy(v, v + 13, plan);
    } else if (impl == B200CONV_ALGO_UPOLS) {
        int rc = plan_upols(&tmp);
        if (rc) return rc;
        plan[0] = tmp.up.P;
        plan[1] = tmp.up.M;
        plan[2] = tmp.up.logM;
        plan[3] = tmp.up.S;
    } else if (impl == B200CONV_ALGO_DIRECT_TC) {
        int rc = plan_tc(&tmp);
        if (rc) return rc;
        const TcGeometry& g = tmp.tc.g;
        const int32_t v[9] = {g.A, g.C, g.NE, g.NGRP, g.R, g.capP, static_cast<int32_t>(g.smem_bytes), tmp.tc.grid, g.N};
        std::copy(v, v + 9, plan);
    } else {
        return fail(B200CONV_ERR_INVALID, "b200conv_plan: unknown algo");
    }
    plan[15] = static_cast<int32_t>(impl);  // which engine the planner chose (b200conv_algo value)
    return B200CONV_OK;
}

int b200conv_create(const b200conv_config* cfg, b200conv_engine** out) {
    if (!cfg || !out) return fail(B200CONV_ERR_INVALID, "b200conv_create: null argument");
    *out = nullptr;
    if (cfg->abi_version != B200CONV_ABI_VERSION)
        return fail(B200CONV_ERR_ABI, "b200conv_create: abi_version mismatch");
    if (cfg->tracks == 0 || cfg->block == 0 || cfg->ir_len == 0)
        return fail(B200CONV_ERR_INVALID, "b200conv_create: tracks, block and ir_len must be > 0");
    if (cfg->algo > B200CONV_ALGO_DIRECT_TC || cfg->out_layout > B200CONV_OUT_SAMPLE_MAJOR)
        return fail(B200CONV_ERR_INVALID, "b200conv_create: unknown algo or layout");
    const uint32_t Tg = cfg->total_tracks ? cfg->total_tracks : cfg->tracks;
    if (cfg->track_offset + cfg->tracks > Tg)
        return fail(B200CONV_ERR_INVALID, "b200conv_create: track_offset + tracks exceeds total_tracks");

    int ndev = 0;
    cudaError_t err = cudaGetDeviceCount(&ndev);
    if (err != cudaSuccess || ndev == 0)
        return fail(B200CONV_ERR_NO_DEVICE, std::string("no CUDA device: ") +
                                                (err != cudaSuccess ? cudaGetErrorString(err) : "device count is 0") +
                                                " (this engine has no CPU fallback)");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(B200CONV_ERR_INVALID, "b200conv_create: bad device ordinal");
    cudaDeviceProp prop{};
    CU_TRY(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10)
        return fail(B200CONV_ERR_NO_DEVICE, std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                                                std::to_string(prop.minor) + "; kernels are built for sm_100a only");
    ENGINE_DEVICE(cfg->device);

    auto* e = new b200conv_engine();
    e->cfg = *cfg;
    e->T = static_cast<int>(cfg->tracks);
    e->B = static_cast<int>(cfg->block);
    e->L = static_cast<int>(cfg->ir_len);
    e->Tg = static_cast<int>(Tg);
    e->toff = static_cast<int>(cfg->track_offset);
    e->sm_count = prop.multiProcessorCount;

    const uint32_t impl = resolve_impl(*cfg);
    e->impl = impl;
    int rc = (impl == B200CONV_ALGO_DIRECT) ? plan_direct(e) : (impl == B200CONV_ALGO_UPOLS ? plan_upols(e) : plan_tc(e));
    e->bus_in_kernel_single = env_int("B200CONV_BUS_TREE", 0) != 0;
    auto bail = [&](int code) {
        b200conv_destroy(e);
        return code;
    };
    if (rc) return bail(rc);

    const size_t tb = static_cast<size_t>(e->T) * e->B;
    const size_t out_elems = (cfg->out_layout == B200CONV_OUT_SAMPLE_MAJOR) ? static_cast<size_t>(e->B) * e->Tg : tb;
    if ((rc = dev_alloc(e, &e->d_in_stage, tb))) return bail(rc);
    if ((rc = dev_alloc(e, &e->d_out_stage, out_elems))) return bail(rc);
    if ((rc = dev_alloc(e, &e->d_mix_stage, static_cast<size_t>(2) * e->B))) return bail(rc);
    if ((rc = dev_alloc(e, &e->d_gains, static_cast<size_t>(2) * e->T))) return bail(rc);
    if ((rc = set_default_gains(e))) return bail(rc);
    {   // bus tree geometry: groups of ~sqrt(T) tracks; column chunks = the kernels' output tiles
        int g1 = 1;
        while (g1 < 64 && g1 * g1 < e->T) g1 *= 2;
        e->bus_G1 = g1;
        e->bus_NG = (e->T + g1 - 1) / g1;
        if (impl == B200CONV_ALGO_DIRECT) {
            e->bus_CH = e->dir.A * 16;
            e->bus_NC = e->dir.ntiles;
        } else if (impl == B200CONV_ALGO_DIRECT_TC) {
            e->bus_CH = e->tc.Bs / e->tc.nch;  // <= 512: 128 epilogue threads own two column pairs each
            e->bus_NC = e->B / e->bus_CH;
        } else {
            e->bus_CH = e->B;
            e->bus_NC = 1;
        }
        if (e->bus_NC > kBusMaxChunks) return bail(fail(B200CONV_ERR_INVALID, "b200conv_create: block too large for the bus tree"));
        if ((rc = dev_alloc(e, &e->d_ybus, tb))) return bail(rc);
        if ((rc = dev_alloc(e, &e->d_gpart, static_cast<size_t>(e->bus_NG) * (e->B / 2)))) return bail(rc);
        if ((rc = dev_alloc(e, &e->d_gcount, static_cast<size_t>(e->bus_NG) * e->bus_NC))) return bail(rc);
        if ((rc = dev_alloc(e, &e->d_ccount, static_cast<size_t>(e->bus_NC)))) return bail(rc);
        err = cudaHostAlloc(reinterpret_cast<void**>(&e->d_bus_err), sizeof(uint32_t), cudaHostAllocMapped | cudaHostAllocPortable);
        if (err != cudaSuccess) return bail(fail(B200CONV_ERR_CUDA, std::string("cudaHostAlloc: ") + cudaGetErrorString(err)));
        *e->d_bus_err = 0;
        if (env_int("B200CONV_BUS_TRACE", 0))
            if ((rc = dev_alloc(e, &e->d_bus_trace, static_cast<size_t>(2) * kBusTraceLen))) return bail(rc);
    }

    if (impl == B200CONV_ALGO_DIRECT) {
        DirectState& d = e->dir;
        if ((rc = dev_alloc(e, &d.h, static_cast<size_t>(e->T) * d.Lc * 16))) return bail(rc);
        if ((rc = dev_alloc(e, &d.ring, static_cast<size_t>(e->T) * d.cap))) return bail(rc);
        if ((rc = dev_alloc(e, &d.partial, static_cast<size_t>(d.MS) * tb))) return bail(rc);
    } else if (impl == B200CONV_ALGO_DIRECT_TC) {
        TcState& c = e->tc;
        if ((rc = dev_alloc(e, &c.bimg, static_cast<size_t>(e->T) * c.g.NGRP * 2 * c.g.image_floats))) return bail(rc);
        if ((rc = dev_alloc(e, &c.hhead, tb))) return bail(rc);
        if ((rc = dev_alloc(e, &c.arrive, 1))) return bail(rc);
        if (env_int("B200CONV_TC_TRACE", 0))
            if ((rc = dev_alloc(e, &c.trace, static_cast<size_t>(c.grid) * kTcTraceSlots))) return bail(rc);
        if ((rc = dev_alloc(e, &c.pend, static_cast<size_t>(e->T) * c.g.capP))) return bail(rc);
        if (c.nsub > 1)
            if ((rc = dev_alloc(e, &c.shadow, static_cast<size_t>(e->T) * (c.g.capP + 128)))) return bail(rc);
        if ((rc = dev_alloc(e, &c.xprev, static_cast<size_t>(2) * e->T * 128))) return bail(rc);
    } else {
        UpolsState& u = e->up;
        const size_t spec = static_cast<size_t>(e->T) * u.P * u.M;
        if ((rc = dev_alloc(e, &u.H, spec))) return bail(rc);
        if ((rc = dev_alloc(e, &u.X, spec))) return bail(rc);
        if ((rc = dev_alloc(e, &u.Ypart, static_cast<size_t>(u.S) * e->T * u.M))) return bail(rc);
        if ((rc = dev_alloc(e, &u.prev, 2 * tb))) return bail(rc);
        if ((rc = dev_alloc(e, &u.counters, static_cast<size_t>(e->T)))) return bail(rc);
        if ((rc = dev_alloc(e, &u.xready, static_cast<size_t>(e->T)))) return bail(rc);
    }
    err = cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking);
    if (err != cudaSuccess) return bail(fail(B200CONV_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(err)));
    for (int i = 0; i < 2; ++i) {
        err = cudaEventCreateWithFlags(&e->ev_done[i], cudaEventDisableTiming);
        if (err == cudaSuccess) err = cudaEventCreateWithFlags(&e->ev_out[i], cudaEventDisableTiming);
        if (err != cudaSuccess) return bail(fail(B200CONV_ERR_CUDA, std::string("cudaEventCreate: ") + cudaGetErrorString(err)));
    }
    for (auto& ev : e->ev) {
        err = cudaEventCreate(&ev);
        if (err != cudaSuccess) return bail(fail(B200CONV_ERR_CUDA, std::string("cudaEventCreate: ") + cudaGetErrorString(err)));
    }
    *out = e;
    return B200CONV_OK;
}

void b200conv_destroy(b200conv_engine* e) {
    if (!e) return;
    DeviceGuard guard(e->cfg.device);
    cudaDeviceSynchronize();
    for (void* p : e->allocs) cudaFree(p);
    for (auto& ev : e->ev)
        if (ev) cudaEventDestroy(ev);
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    for (int i = 0; i < 2; ++i) {
        if (e->ev_done[i]) cudaEventDestroy(e->ev_done[i]);
        if (e->ev_out[i]) cudaEventDestroy(e->ev_out[i]);
    }
    if (e->d_bus_err) cudaFreeHost(e->d_bus_err);
    delete e;
}

int b200conv_load_ir(b200conv_engine* e, const float* host_ir) {
    if (!e || !host_ir) return fail(B200CONV_ERR_INVALID, "b200conv_load_ir: null argument");
    ENGINE_DEVICE(e->cfg.device);
    const int T = e->T, L = e->L, B = e->B;
    if (e->impl == B200CONV_ALGO_DIRECT) {
        DirectState& d = e->dir;
        const size_t row = static_cast<size_t>(d.Lc) * 16;
        const int chunk = static_cast<int>(std::max<size_t>(1, (64u << 20) / (row * sizeof(float))));
        std::vector<float> stage(static_cast<size_t>(std::min(chunk, T)) * row);
        for (int t0 = 0; t0 < T; t0 += chunk) {
            const int nt = std::min(chunk, T - t0);
            std::fill(stage.begin(), stage.begin() + static_cast<size_t>(nt) * row, 0.0f);
            for (int t = 0; t < nt; ++t) {
                const float* src = host_ir + static_cast<size_t>(t0 + t) * L;
                float* dst = stage.data() + static_cast<size_t>(t) * row;
                // swizzled only when lanes of a warp read different tap blocks (otherwise a broadcast)
                if (d.CL > 1) {
                    for (int j = 0; j < L; ++j) dst[swz_float(static_cast<uint32_t>(j))] = src[j];
                } else {
                    std::memcpy(dst, src, static_cast<size_t>(L) * sizeof(float));
                }
            }
            CU_TRY(cudaMemcpy(d.h + static_cast<size_t>(t0) * row, stage.data(), static_cast<size_t>(nt) * row * sizeof(float),
                              cudaMemcpyHostToDevice));
        }
    } else if (e->impl == B200CONV_ALGO_DIRECT_TC) {
        // hi / lo TF32 tap images in the tensor core's K-major core-matrix layout, one per (track, column group)
        TcState& c = e->tc;
        const size_t per_track = static_cast<size_t>(c.g.NGRP) * 2 * c.g.image_floats;
        const int chunk = static_cast<int>(std::max<size_t>(1, (64u << 20) / (per_track * sizeof(float))));
        std::vector<float> stage(static_cast<size_t>(std::min(chunk, T)) * per_track);
        for (int t0 = 0; t0 < T; t0 += chunk) {
            const int nt = std::min(chunk, T - t0);
            for (int t = 0; t < nt; ++t)
                tc_build_images(host_ir + static_cast<size_t>(t0 + t) * L, L, c.g, stage.data() + static_cast<size_t>(t) * per_track);
            CU_TRY(cudaMemcpy(c.bimg + static_cast<size_t>(t0) * per_track, stage.data(),
                              static_cast<size_t>(nt) * per_track * sizeof(float), cudaMemcpyHostToDevice));
        }
        std::vector<float> head(static_cast<size_t>(T) * c.Bs, 0.0f);  // the taps of a launch's own samples (k < Bs)
        for (int t = 0; t < T; ++t)
            std::memcpy(head.data() + static_cast<size_t>(t) * c.Bs, host_ir + static_cast<size_t>(t) * L,
                        static_cast<size_t>(std::min(c.Bs, L)) * sizeof(float));
        CU_TRY(cudaMemcpy(c.hhead, head.data(), head.size() * sizeof(float), cudaMemcpyHostToDevice));
    } else {
        UpolsState& u = e->up;
        const size_t row = static_cast<size_t>(u.P) * B;  // taps zero padded to P*B
        const int chunk = static_cast<int>(std::max<size_t>(1, (128u << 20) / (row * sizeof(float))));
        const int nmax = std::min(chunk, T);
        std::vector<float> stage(static_cast<size_t>(nmax) * row);
        float* d_pad = nullptr;
        CU_TRY(cudaMalloc(&d_pad, static_cast<size_t>(nmax) * row * sizeof(float)));
        int rc = B200CONV_OK;
        for (int t0 = 0; t0 < T && rc == B200CONV_OK; t0 += chunk) {
            const int nt = std::min(chunk, T - t0);
            for (int t = 0; t < nt; ++t) {
                float* dst = stage.data() + static_cast<size_t>(t) * row;
                std::memcpy(dst, host_ir + static_cast<size_t>(t0 + t) * L, static_cast<size_t>(L) * sizeof(float));
                std::fill(dst + L, dst + row, 0.0f);
            }
            cudaError_t err = cudaMemcpy(d_pad, stage.data(), static_cast<size_t>(nt) * row * sizeof(float), cudaMemcpyHostToDevice);
            if (err == cudaSuccess) {
                RfftParams p{};
                p.first = d_pad;
                p.first_stride = B;
                p.second = nullptr;
                p.second_stride = 0;
                p.out = u.H + static_cast<size_t>(t0) * u.P * u.M;
                p.out_stride = u.M;
                p.prev_out = nullptr;
                p.count = nt * u.P;
                p.M = u.M;
                p.logM = u.logM;
                p.scale = 1.0f / static_cast<float>(2 * B);
                err = launch_rfft_fwd(p, nullptr);
                e->launches += 1;
                if (err == cudaSuccess) err = cudaDeviceSynchronize();
            }
            if (err != cudaSuccess) rc = fail(B200CONV_ERR_CUDA, std::string("load_ir (UPOLS spectra): ") + cudaGetErrorString(err));
        }
        cudaFree(d_pad);
        if (rc) return rc;
    }
    e->ir_loaded = true;
    return reset_state(e);
}

int b200conv_reset(b200conv_engine* e) {
    if (!e) return fail(B200CONV_ERR_INVALID, "b200conv_reset: null engine");
    ENGINE_DEVICE(e->cfg.device);
    CU_TRY(cudaDeviceSynchronize());
    return reset_state(e);
}

int b200conv_prime_history(b200conv_engine* e, const float* host_hist) {
    if (!e) return fail(B200CONV_ERR_INVALID, "b200conv_prime_history: null engine");
    ENGINE_DEVICE(e->cfg.device);
    CU_TRY(cudaDeviceSynchronize());
    int rc = reset_state(e);
    if (rc || !host_hist || e->L < 2) return rc;
    const int T = e->T, L = e->L, B = e->B, H = L - 1;
    if (e->impl == B200CONV_ALGO_DIRECT_TC) {
        // the state of this engine is the pending OUTPUT of past input (overlap-add), so history is primed by
        // streaming it: ceil((L-1)/B) buffers, zero padded at the front, outputs discarded
        const int nb = (H + B - 1) / B;
        const size_t tb = static_cast<size_t>(T) * B;
        std::vector<float> blk(tb);
        for (int m = 0; m < nb; ++m) {
            for (int t = 0; t < T; ++t)
                for (int i = 0; i < B; ++i) {
                    const long long src = static_cast<long long>(m) * B + i - (static_cast<long long>(nb) * B - H);
                    blk[static_cast<size_t>(t) * B + i] = (src >= 0) ? host_hist[static_cast<size_t>(t) * H + src] : 0.0f;
                }
            // stream-ordered with the kernel that reads it: a plain cudaMemcpy from PAGEABLE memory may return before
            // its DMA has landed, and the engine's stream is non-blocking (does not wait for the legacy stream)
            CU_TRY(cudaMemcpyAsync(e->d_in_stage, blk.data(), tb * sizeof(float), cudaMemcpyHostToDevice, e->own_stream));
            rc = b200conv_process(e, e->d_in_stage, e->d_out_stage, nullptr, 0, e->own_stream);
            if (rc) return rc;
            CU_TRY(cudaStreamSynchronize(e->own_stream));
        }
        e->blocks = 0;
        return B200CONV_OK;
    }
    if (e->impl == B200CONV_ALGO_DIRECT) {
        // ring position pos = 0 is the next block; history occupies the last L-1 floats of the ring.
        DirectState& d = e->dir;
        const size_t row = d.cap;
        const int chunk = static_cast<int>(std::max<size_t>(1, (64u << 20) / (row * sizeof(float))));
        std::vector<float> stage(static_cast<size_t>(std::min(chunk, T)) * row);
        for (int t0 = 0; t0 < T; t0 += chunk) {
            const int nt = std::min(chunk, T - t0);
            std::fill(stage.begin(), stage.begin() + static_cast<size_t>(nt) * row, 0.0f);
            for (int t = 0; t < nt; ++t) {
                const float* src = host_hist + static_cast<size_t>(t0 + t) * H;
                float* dst = stage.data() + static_cast<size_t>(t) * row;
                for (int i = 0; i < H; ++i) dst[swz_float(static_cast<uint32_t>(d.cap - H + i))] = src[i];
            }
            CU_TRY(cudaMemcpy(d.ring + static_cast<size_t>(t0) * row, stage.data(), static_cast<size_t>(nt) * row * sizeof(float),
                              cudaMemcpyHostToDevice));
        }
    } else {
        // Push history blocks -(P-1) .. -1 through the forward transform into ring slots P-1 .. 1.
        UpolsState& u = e->up;
        const size_t row = static_cast<size_t>(u.P) * B;  // samples n = -P*B .. -1
        std::vector<float> hb(static_cast<size_t>(T) * row, 0.0f);
        for (int t = 0; t < T; ++t)
            std::memcpy(hb.data() + static_cast<size_t>(t) * row + (row - H), host_hist + static_cast<size_t>(t) * H,
                        static_cast<size_t>(H) * sizeof(float));
        float* d_hb = nullptr;
        CU_TRY(cudaMalloc(&d_hb, hb.size() * sizeof(float)));
        cudaError_t err = cudaMemcpy(d_hb, hb.data(), hb.size() * sizeof(float), cudaMemcpyHostToDevice);
        for (int j = -(u.P - 1); j <= -1 && err == cudaSuccess; ++j) {
            RfftParams p{};
            p.first = d_hb + static_cast<size_t>(u.P + j - 1) * B;
            p.first_stride = row;
            p.second = d_hb + static_cast<size_t>(u.P + j) * B;
            p.second_stride = row;
            p.out = u.X + static_cast<size_t>(-j) * u.M;
            p.out_stride = static_cast<size_t>(u.P) * u.M;
            p.prev_out = nullptr;
            p.count = T;
            p.M = u.M;
            p.logM = u.logM;
            p.scale = 1.0f;
            err = launch_rfft_fwd(p, nullptr);
            e->launches += 1;
        }
        if (err == cudaSuccess)
            err = cudaMemcpy2D(u.prev, static_cast<size_t>(B) * sizeof(float), d_hb + static_cast<size_t>(u.P - 1) * B,
                               row * sizeof(float), static_cast<size_t>(B) * sizeof(float), T, cudaMemcpyDeviceToDevice);
        if (err == cudaSuccess) err = cudaDeviceSynchronize();
        cudaFree(d_hb);
        if (err != cudaSuccess) return fail(B200CONV_ERR_CUDA, std::string("prime_history (UPOLS): ") + cudaGetErrorString(err));
    }
    CU_TRY(cudaDeviceSynchronize());  // pageable H2D copies may still be in flight when cudaMemcpy returns
    return B200CONV_OK;
}

int b200conv_set_mix_gains(b200conv_engine* e, const float* host_gains) {
    if (!e) return fail(B200CONV_ERR_INVALID, "b200conv_set_mix_gains: null engine");
    ENGINE_DEVICE(e->cfg.device);
    if (!host_gains) return set_default_gains(e);
    CU_TRY(cudaMemcpy(e->d_gains, host_gains, static_cast<size_t>(2) * e->T * sizeof(float), cudaMemcpyHostToDevice));
    CU_TRY(cudaDeviceSynchronize());  // pageable H2D: the DMA may still be in flight when cudaMemcpy returns
    return B200CONV_OK;
}

static int process_impl(b200conv_engine* e, const float* d_in, float* d_out, float* d_out2, float* d_mix, uint32_t flags,
                        void* stream);

int b200conv_process(b200conv_engine* e, const float* d_in, float* d_out, float* d_mix, uint32_t flags, void* stream) {
    return process_impl(e, d_in, d_out, nullptr, d_mix, flags, stream);
}

// d_out2: optional second copy of the output (fused UPOLS kernel only; null otherwise)
static int process_impl(b200conv_engine* e, const float* d_in, float* d_out, float* d_out2, float* d_mix, uint32_t flags,
                        void* stream) {
    if (!e || !d_in || !d_out) return fail(B200CONV_ERR_INVALID, "b200conv_process: null argument");
    if (!e->ir_loaded) return fail(B200CONV_ERR_STATE, "b200conv_process: call b200conv_load_ir first");
    ENGINE_DEVICE(e->cfg.device);  // launches, smem opt-ins and the null stream are those of the engine's device
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // the engine's state lives on whatever stream the previous block ran on: switching streams is
    // allowed, but the new stream must see the old one's work (cheap: only when the stream changes)
    if (e->has_last_stream && e->last_stream != st) CU_TRY(cudaStreamSynchronize(e->last_stream));
    e->last_stream = st;
    e->has_last_stream = true;
    const bool commit = !(flags & B200CONV_PEEK);
    const int sample_major = (e->cfg.out_layout == B200CONV_OUT_SAMPLE_MAJOR);
    StageTimer tm{e, st};
    int marks = 0;
    if (d_mix && e->bus_world > 1) e->bus_epoch += 1;  // one exchange per block with a bus, in lockstep on every rank

    if (e->impl == B200CONV_ALGO_DIRECT) {
        DirectState& d = e->dir;
        tm.mark();
        FirParams p{};
        p.h = d.h;
        p.ring = d.ring;
        p.d_in = d_in;
        p.partial = d.partial;
        p.T = e->T;
        p.B = e->B;
        p.capb = d.cap / 16;
        p.posb = d.pos / 16;
        p.Lc = d.Lc;
        p.JSb = d.JSb;
        p.NS = d.NS;
        p.SPS = d.SPS;
        p.nbuf = d.nbuf;
        p.xtile_blocks = d.xtile_blocks;
        p.ntiles = d.ntiles;
        p.U = e->T * d.ntiles * d.NS;
        p.G = d.G;
        CU_TRY(launch_fir(p, d.A, d.smem, st));
        tm.mark();
        // tail of the block (PDL-launched behind the FIR): rows -> output, ring append, bus (+ NVLink exchange)
        FinishParams f{};
        f.partial = d.partial;
        f.out = d_out;
        f.MS = d.MS;
        f.T = e->T;
        f.B = e->B;
        f.sample_major = sample_major;
        f.Tg = e->Tg;
        f.toff = e->toff;
        f.gains = e->d_gains;
        f.mix = e->strip_ops ? nullptr : d_mix;  // with a strip the bus is taken after it
        f.d_in = d_in;
        f.ring = commit ? d.ring : nullptr;
        f.cap = d.cap;
        f.pos = d.pos;
        f.x = bus_exchange(e);
        CU_TRY(launch_fir_finish_mix(f, st));
        e->launches += 2;
        if (e->strip_ops) {
            int rc = run_strip(e, d_out, commit, st);
            if (rc) return rc;
            if (d_mix) {
                CU_TRY(launch_mix_cluster(d_out, sample_major, e->Tg, e->toff, e->d_gains, d_mix, e->T, e->B, bus_exchange(e), st));
                e->launches += 1;
            }
        }
        tm.mark();
        marks = tm.idx;
        if (commit) d.pos = (d.pos + e->B) % d.cap;
    } else if (e->impl == B200CONV_ALGO_DIRECT_TC) {
        TcState& c = e->tc;
        tm.mark();
        TcParams p{};
        p.d_in = d_in;
        p.xprev = c.xprev;
        p.bimg = c.bimg;
        p.hhead = c.hhead;
        p.trace = c.trace;
        p.pend = c.pend;
        p.out = d_out;
        p.T = e->T;
        p.B = c.Bs;
        p.in_stride = e->B;
        p.out_stride = e->B;
        p.nchunk = c.nch;
        p.A = c.g.A;
        p.C = c.g.C;
        p.NE = c.g.NE;
        p.NGRP = c.g.NGRP;
        p.N = c.g.N;
        p.tmem_cols = static_cast<uint32_t>(c.g.tmem_cols);
        p.R = c.g.R;
        p.capP = c.g.capP;
        p.sample_major = sample_major;
        p.Tg = e->Tg;
        p.toff = e->toff;
        p.debug = env_int("B200CONV_TC_DEBUG", 0);  // measurement only (skip phases)
        p.bus = bus_params(e, e->strip_ops ? nullptr : d_mix);
        // A PEEK of several sub-blocks has to commit the first ones to compute the later ones: the stream state
        // (pending ring, previous-128 window) is saved before and put back afterwards.
        const bool peek_multi = !commit && c.nsub > 1;
        const size_t ring_floats = static_cast<size_t>(e->T) * c.g.capP;
        int ppos = c.ppos, xpar = c.xpar;
        if (peek_multi) {
            CU_TRY(cudaMemcpyAsync(c.shadow, c.pend, ring_floats * sizeof(float), cudaMemcpyDeviceToDevice, st));
            CU_TRY(cudaMemcpyAsync(c.shadow + ring_floats, c.xprev + static_cast<size_t>(xpar) * e->T * 128,
                                   static_cast<size_t>(e->T) * 128 * sizeof(float), cudaMemcpyDeviceToDevice, st));
        }
        for (int s = 0; s < c.nsub; ++s) {
            p.n_off = s * c.Bs;
            p.chunk0 = s * c.nch;
            p.ppos = ppos;
            p.xpar = xpar;
            p.commit = (commit || s + 1 < c.nsub) ? 1 : 0;
            p.slice = BusSlice{};
            if (p.bus.mix && c.slice) {
                p.slice.arrive = c.arrive;
                p.slice.target = static_cast<unsigned long long>(e->T) * ++c.seq;
                p.slice.slice = c.slice;
            }
            CU_TRY(launch_tc_toeplitz(p, c.grid, st));
            e->launches += 1;
            if (p.commit) {
                ppos = (ppos + c.Bs) % c.g.capP;
                xpar ^= 1;
            }
        }
        if (peek_multi) {
            CU_TRY(cudaMemcpyAsync(c.pend, c.shadow, ring_floats * sizeof(float), cudaMemcpyDeviceToDevice, st));
            CU_TRY(cudaMemcpyAsync(c.xprev + static_cast<size_t>(c.xpar) * e->T * 128, c.shadow + ring_floats,
                                   static_cast<size_t>(e->T) * 128 * sizeof(float), cudaMemcpyDeviceToDevice, st));
        } else if (commit) {
            c.ppos = ppos;
            c.xpar = xpar;
        }
        tm.mark();
        if (e->strip_ops) {
            int rc = run_strip(e, d_out, commit, st);
            if (rc) return rc;
            if (d_mix) {
                CU_TRY(launch_mix_cluster(d_out, sample_major, e->Tg, e->toff, e->d_gains, d_mix, e->T, e->B, bus_exchange(e), st));
                e->launches += 1;
            }
            tm.mark();
        }
        marks = tm.idx;
    } else {
        UpolsState& u = e->up;
        const int slot0 = static_cast<int>((u.P - (e->blocks % u.P)) % u.P);
        if (u.fused) {
            tm.mark();
            FusedParams fp{};
            fp.d_in = d_in;
            fp.prev = u.prev + static_cast<size_t>(u.par) * e->T * e->B;
            fp.prev_w = u.prev + static_cast<size_t>(u.par ^ 1) * e->T * e->B;
            fp.KT = std::max(1, u.M / 512);
            fp.xready = u.xready;
            if (++u.seq == 0) u.seq = 1;
            fp.seq = u.seq;
            fp.H = u.H;
            fp.X = u.X;
            fp.Ypart = u.Ypart;
            fp.counters = u.counters;
            fp.out = d_out;
            fp.out2 = d_out2;
            fp.T = e->T;
            fp.P = u.P;
            fp.M = u.M;
            fp.logM = u.logM;
            fp.S = u.S;
            fp.slot0 = slot0;
            fp.commit = commit ? 1 : 0;
            fp.sample_major = sample_major;
            fp.Tg = e->Tg;
            fp.toff = e->toff;
            if (e->strip_ops) fp.strip = strip_params(e, d_out, commit);  // strip runs inside the kernel's epilogue
            // The bus stays in the PDL-launched bus kernel (which also carries the NVLink exchange of a multi-GPU
            // job): inside this kernel it costs a ticket round trip per track that the streaming CTA cannot hide,
            // measured 128 -> 133 us on the C4 shard and 166 -> 171 us at C3.  B200CONV_BUS_TREE=1 selects the
            // in-kernel tree anyway (one launch per block; kept for measurement and tested).
            const bool tree = e->bus_in_kernel_single;
            fp.bus = bus_params(e, tree ? d_mix : nullptr);
            CU_TRY(launch_upols_fused(fp, st));
            if (commit) u.par ^= 1;
            e->launches += 1;
            tm.mark();
        } else {
            tm.mark();
            RfftParams f{};
            float* prev_cur = u.prev + static_cast<size_t>(u.par) * e->T * e->B;  // three-kernel path: in place, no flip
            f.first = prev_cur;
            f.first_stride = e->B;
            f.second = d_in;
            f.second_stride = e->B;
            f.out = u.X + static_cast<size_t>(slot0) * u.M;
            f.out_stride = static_cast<size_t>(u.P) * u.M;
            f.prev_out = commit ? prev_cur : nullptr;
            f.count = e->T;
            f.M = u.M;
            f.logM = u.logM;
            f.scale = 1.0f;
            CU_TRY(launch_rfft_fwd(f, st));
            tm.mark();
            MacParams m{};
            m.H = u.H;
            m.X = u.X;
            m.Ypart = u.Ypart;
            m.T = e->T;
            m.P = u.P;
            m.M = u.M;
            m.slot0 = slot0;
            m.S = u.S;
            CU_TRY(launch_fdl_mac(m, st));
            tm.mark();
            IrfftParams r{};
            r.Ypart = u.Ypart;
            r.S = u.S;
            r.out = d_out;
            r.T = e->T;
            r.M = u.M;
            r.logM = u.logM;
            r.sample_major = sample_major;
            r.Tg = e->Tg;
            r.toff = e->toff;
            CU_TRY(launch_irfft_ols(r, st));
            e->launches += 3;
        }
        if (e->strip_ops && !u.fused) {
            int rc = run_strip(e, d_out, commit, st);
            if (rc) return rc;
        }
        const bool tree_done = u.fused && e->bus_in_kernel_single;
        if (d_mix && !tree_done) {
            CU_TRY(launch_mix_cluster(d_out, sample_major, e->Tg, e->toff, e->d_gains, d_mix, e->T, e->B, bus_exchange(e), st));
            e->launches += 1;
        }
        if (!u.fused || (d_mix && !tree_done)) tm.mark();
        marks = tm.idx;
    }
    if (commit) e->blocks += 1;
    return finish_profile(e, marks);
}

namespace {
// Pinned (cudaMallocHost / cudaHostRegister) memory is mapped into the device address space under
// UVA: kernels can read and write it directly over PCIe, which removes the copy-engine hand-offs
// (H2D -> kernel -> D2H dependencies cost ~10 us each) from the host-buffer call.
bool is_pinned_host(const void* p) {
    if (!p) return false;
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost;
}
}  // namespace

namespace {
int wait_slot(b200conv_engine* e, int slot) {
    CU_TRY(cudaEventSynchronize(e->ev_out[slot]));
    if (e->bus_world > 1 && e->slot_has_bus[slot] && *static_cast<volatile uint32_t*>(e->d_bus_err)) {
        *e->d_bus_err = 0;  // did a peer engine miss the bus exchange of this block?
        return fail(B200CONV_ERR_CUDA, "bus all-reduce: a peer engine did not signal within the spin bound");
    }
    return B200CONV_OK;
}

// Enqueue one host-buffer block on staging slot `slot` and record ev_out[slot] when its results are on the host.
int submit_slot(b200conv_engine* e, const float* h_in, float* h_out, float* h_mix, uint32_t flags, int slot) {
    cudaStream_t st = e->own_stream;
    const size_t tb = static_cast<size_t>(e->T) * e->B;
    const size_t out_elems = (e->cfg.out_layout == B200CONV_OUT_SAMPLE_MAJOR) ? static_cast<size_t>(e->B) * e->Tg : tb;
    if (slot == 1 && !e->d_in_stage1) {  // second staging set: allocated on the first pipelined submit
        int rc = dev_alloc(e, &e->d_in_stage1, tb);
        if (!rc) rc = dev_alloc(e, &e->d_out_stage1, out_elems);
        if (!rc) rc = dev_alloc(e, &e->d_mix_stage1, static_cast<size_t>(2) * e->B);
        if (rc) return rc;
    }
    float* in_stage = slot ? e->d_in_stage1 : e->d_in_stage;
    float* out_stage = slot ? e->d_out_stage1 : e->d_out_stage;
    float* mix_stage = slot ? e->d_mix_stage1 : e->d_mix_stage;
    const int zc = env_int("B200CONV_ZEROCOPY", 3);  // bit 0: read the input in place; bit 1: write results in place
    const bool direct = (e->impl == B200CONV_ALGO_DIRECT || e->impl == B200CONV_ALGO_DIRECT_TC);
    // input: every engine reads d_in once or twice -> read it straight from pinned host memory
    const bool in_place = (zc & 1) && is_pinned_host(h_in);
    // results: the kernels that finish a track in their own epilogue (direct FIR, tensor-core FIR, fused UPOLS) can
    // post the output rows and the bus straight to pinned host memory — the bus tree sums from a device-side copy
    // of the rows, so nothing is read back over PCIe.  Not for: a strip kernel after the FIR (it works in place on
    // the output), the three-kernel UPOLS path, and sample-major UPOLS (a column tile written track by track is
    // scattered 4-byte PCIe writes: measured 2x slower than the staged copy).
    const bool fused_ok = !direct && e->up.fused && e->cfg.out_layout == B200CONV_OUT_TRACK_MAJOR;
    const bool tree = e->bus_in_kernel_single;  // UPOLS: the bus rides in the fused kernel (measurement option)
    const bool pinned_results = (zc & 2) && h_out && is_pinned_host(h_out) && (!h_mix || is_pinned_host(h_mix));
    const bool out_place = pinned_results && ((direct && !e->strip_ops) || (fused_ok && (tree || !h_mix)));
    // stand-alone fused UPOLS with a bus: the PDL-launched bus kernel reads the output back, so the fused kernel
    // keeps a device copy and posts a SECOND copy of each finished row straight to the pinned host buffer
    const bool dual = pinned_results && !out_place && fused_ok;
    e->slot_has_bus[slot] = (h_mix != nullptr);
    // the slot's staging buffers are free once the device -> host copies of the block that used them are done
    CU_TRY(cudaStreamWaitEvent(st, e->ev_out[slot], 0));
    const float* d_in = h_in;
    if (!in_place) {
        CU_TRY(cudaMemcpyAsync(in_stage, h_in, tb * sizeof(float), cudaMemcpyHostToDevice, st));
        d_in = in_stage;
    }
    if (out_place) {
        int rc = b200conv_process(e, d_in, h_out, h_mix, flags, st);
        if (rc) return rc;
        CU_TRY(cudaEventRecord(e->ev_out[slot], st));
        return B200CONV_OK;
    }
    const bool mix_place = (zc & 2) && h_mix && is_pinned_host(h_mix);  // 2*B floats: written in place whatever the output path
    int rc = process_impl(e, d_in, out_stage, dual ? h_out : nullptr, h_mix ? (mix_place ? h_mix : mix_stage) : nullptr, flags, st);
    if (rc) return rc;
    const bool need_copy = (h_out && !dual) || (h_mix && !mix_place);
    if (!need_copy) {
        CU_TRY(cudaEventRecord(e->ev_out[slot], st));
        return B200CONV_OK;
    }
    // device -> host on the copy stream: it overlaps the kernels of the NEXT submitted block
    CU_TRY(cudaEventRecord(e->ev_done[slot], st));
    cudaStream_t cs = e->copy_stream;
    CU_TRY(cudaStreamWaitEvent(cs, e->ev_done[slot], 0));
    if (h_out && !dual) {
        if (e->cfg.out_layout == B200CONV_OUT_SAMPLE_MAJOR && e->Tg != e->T) {
            CU_TRY(cudaMemcpy2DAsync(h_out + e->toff, static_cast<size_t>(e->Tg) * sizeof(float), out_stage + e->toff,
                                     static_cast<size_t>(e->Tg) * sizeof(float), static_cast<size_t>(e->T) * sizeof(float),
                                     e->B, cudaMemcpyDeviceToHost, cs));
        } else {
            CU_TRY(cudaMemcpyAsync(h_out, out_stage, tb * sizeof(float), cudaMemcpyDeviceToHost, cs));
        }
    }
    if (h_mix && !mix_place)
        CU_TRY(cudaMemcpyAsync(h_mix, mix_stage, static_cast<size_t>(2) * e->B * sizeof(float), cudaMemcpyDeviceToHost, cs));
    CU_TRY(cudaEventRecord(e->ev_out[slot], cs));
    return B200CONV_OK;
}
}  // namespace

int b200conv_submit(b200conv_engine* e, const float* h_in, float* h_out, float* h_mix, uint32_t flags, uint64_t* ticket) {
    if (!e || !h_in || !ticket) return fail(B200CONV_ERR_INVALID, "b200conv_submit: null argument");
    ENGINE_DEVICE(e->cfg.device);
    if (e->submitted - e->completed >= 2) {  // both staging slots in flight: the oldest block has to land first
        int rc = wait_slot(e, static_cast<int>(e->completed & 1));
        if (rc) return rc;
        e->completed += 1;
    }
    const int slot = static_cast<int>(e->submitted & 1);
    int rc = submit_slot(e, h_in, h_out, h_mix, flags, slot);
    if (rc) return rc;
    *ticket = e->submitted++;
    return B200CONV_OK;
}

int b200conv_wait(b200conv_engine* e, uint64_t ticket) {
    if (!e) return fail(B200CONV_ERR_INVALID, "b200conv_wait: null engine");
    if (ticket >= e->submitted) return fail(B200CONV_ERR_STATE, "b200conv_wait: no such ticket");
    ENGINE_DEVICE(e->cfg.device);
    while (e->completed <= ticket) {  // blocks complete in submission order
        int rc = wait_slot(e, static_cast<int>(e->completed & 1));
        e->completed += 1;
        if (rc) return rc;
    }
    return B200CONV_OK;
}

int b200conv_process_host(b200conv_engine* e, const float* h_in, float* h_out, float* h_mix, uint32_t flags) {
    if (!e || !h_in) return fail(B200CONV_ERR_INVALID, "b200conv_process_host: null argument");
    uint64_t ticket = 0;
    int rc = b200conv_submit(e, h_in, h_out, h_mix, flags, &ticket);
    if (rc) return rc;
    return b200conv_wait(e, ticket);
}

int b200conv_set_strip(b200conv_engine* e, const b200conv_strip* strip) {
    if (!e) return fail(B200CONV_ERR_INVALID, "b200conv_set_strip: null engine");
    ENGINE_DEVICE(e->cfg.device);
    CU_TRY(cudaDeviceSynchronize());
    if (!strip || !(strip->ops & (B200CONV_STRIP_STATS | B200CONV_STRIP_GAIN | B200CONV_STRIP_BIQUAD))) {
        e->strip_ops = 0;
        return B200CONV_OK;
    }
    if ((strip->ops & B200CONV_STRIP_BIQUAD) && !strip->biquad)
        return fail(B200CONV_ERR_INVALID, "b200conv_set_strip: BIQUAD needs coefficients");
    const size_t T = static_cast<size_t>(e->T);
    if (!e->d_strip_state || !e->d_strip_stats || !e->d_strip_gains || !e->d_strip_coef) {
        // all or nothing: a partial failure must not leave some buffers set and others null
        float *st = nullptr, *ss = nullptr, *sg = nullptr, *sc = nullptr;
        int rc = dev_alloc(e, &st, 2 * T);
        if (!rc) rc = dev_alloc(e, &ss, 2 * T);
        if (!rc) rc = dev_alloc(e, &sg, T);
        if (!rc) rc = dev_alloc(e, &sc, 5 * T);
        if (rc) return rc;  // whatever was allocated stays in e->allocs and is freed by destroy
        e->d_strip_state = st;
        e->d_strip_stats = ss;
        e->d_strip_gains = sg;
        e->d_strip_coef = sc;
    }
    CU_TRY(cudaMemset(e->d_strip_state, 0, 2 * T * sizeof(float)));
    CU_TRY(cudaMemset(e->d_strip_stats, 0, 2 * T * sizeof(float)));
    e->strip_gain = strip->gain;
    float* gains = nullptr;
    if ((strip->ops & B200CONV_STRIP_GAIN) && strip->gains) {
        CU_TRY(cudaMemcpy(e->d_strip_gains, strip->gains, T * sizeof(float), cudaMemcpyHostToDevice));
        gains = e->d_strip_gains;
    }
    e->strip_shared_coef = (strip->ops & B200CONV_STRIP_SHARED_COEFFS) != 0;
    if (strip->ops & B200CONV_STRIP_BIQUAD)
        CU_TRY(cudaMemcpy(e->d_strip_coef, strip->biquad, (e->strip_shared_coef ? 5 : 5 * T) * sizeof(float),
                          cudaMemcpyHostToDevice));
    e->strip_use_gains = gains != nullptr;
    e->strip_ops = strip->ops & (B200CONV_STRIP_STATS | B200CONV_STRIP_GAIN | B200CONV_STRIP_BIQUAD);
    CU_TRY(cudaDeviceSynchronize());  // pageable H2D copies above
    return B200CONV_OK;
}

int b200conv_strip_state(b200conv_engine* e, float* host_state, int set) {
    if (!e || !host_state) return fail(B200CONV_ERR_INVALID, "b200conv_strip_state: null argument");
    if (!e->d_strip_state) return fail(B200CONV_ERR_STATE, "b200conv_strip_state: no strip attached");
    ENGINE_DEVICE(e->cfg.device);
    CU_TRY(cudaDeviceSynchronize());
    const size_t bytes = static_cast<size_t>(2) * e->T * sizeof(float);
    if (set) {
        CU_TRY(cudaMemcpy(e->d_strip_state, host_state, bytes, cudaMemcpyHostToDevice));
        CU_TRY(cudaDeviceSynchronize());
    } else
        CU_TRY(cudaMemcpy(host_state, e->d_strip_state, bytes, cudaMemcpyDeviceToHost));
    return B200CONV_OK;
}

int b200conv_strip_stats(b200conv_engine* e, float* host_stats) {
    if (!e || !host_stats) return fail(B200CONV_ERR_INVALID, "b200conv_strip_stats: null argument");
    if (!e->d_strip_stats) return fail(B200CONV_ERR_STATE, "b200conv_strip_stats: no strip attached");
    ENGINE_DEVICE(e->cfg.device);
    CU_TRY(cudaDeviceSynchronize());
    CU_TRY(cudaMemcpy(host_stats, e->d_strip_stats, static_cast<size_t>(2) * e->T * sizeof(float), cudaMemcpyDeviceToHost));
    return B200CONV_OK;
}

int b200conv_strip_process(const float* d_in, float* d_out, uint32_t tracks, uint32_t block, uint32_t layout, uint32_t ld,
                           uint32_t col0, const b200conv_strip* strip, float* d_state, float* d_stats, uint32_t flags,
                           void* stream) {
    if (!d_in || !d_out || !strip || tracks == 0 || block == 0)
        return fail(B200CONV_ERR_INVALID, "b200conv_strip_process: null or empty argument");
    const uint32_t ops = strip->ops & (B200CONV_STRIP_STATS | B200CONV_STRIP_GAIN | B200CONV_STRIP_BIQUAD);
    if (!ops) return fail(B200CONV_ERR_INVALID, "b200conv_strip_process: no operation selected");
    if ((ops & B200CONV_STRIP_BIQUAD) && (!strip->biquad || !d_state))
        return fail(B200CONV_ERR_INVALID, "b200conv_strip_process: BIQUAD needs coefficients and a state buffer");
    if (layout == B200CONV_OUT_SAMPLE_MAJOR && static_cast<uint64_t>(col0) + tracks > ld)
        return fail(B200CONV_ERR_INVALID, "b200conv_strip_process: column tile exceeds the leading dimension");
    if (layout != B200CONV_OUT_SAMPLE_MAJOR && layout != B200CONV_OUT_TRACK_MAJOR)
        return fail(B200CONV_ERR_INVALID, "b200conv_strip_process: unknown layout");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return fail(B200CONV_ERR_NO_DEVICE, "b200conv_strip_process: no CUDA device");
    }
    StripParams sp{};
    sp.in = d_in;
    sp.out = d_out;
    sp.T = static_cast<int>(tracks);
    sp.B = static_cast<int>(block);
    sp.sample_major = (layout == B200CONV_OUT_SAMPLE_MAJOR);
    sp.ld = static_cast<int>(ld);
    sp.col0 = static_cast<int>(col0);
    sp.ops = ops;
    sp.gain = strip->gain;
    sp.gains = (ops & B200CONV_STRIP_GAIN) ? strip->gains : nullptr;
    sp.coef = strip->biquad;
    sp.shared_coef = (strip->ops & B200CONV_STRIP_SHARED_COEFFS) ? 1 : 0;
    sp.state = d_state;
    sp.stats = d_stats;
    sp.peek = (flags & B200CONV_PEEK) ? 1 : 0;
    CU_TRY(launch_strip(sp, static_cast<cudaStream_t>(stream)));
    return B200CONV_OK;
}

int b200conv_attach_bus(b200conv_engine* e, const uint64_t* peer_buffers, int rank, int world) {
    if (!e) return fail(B200CONV_ERR_INVALID, "b200conv_attach_bus: null engine");
    ENGINE_DEVICE(e->cfg.device);
    CU_TRY(cudaDeviceSynchronize());
    if (!peer_buffers || world <= 1) {  // detach: a stand-alone engine again
        e->bus_world = 1;
        e->bus_rank = 0;
        e->bus_epoch = 0;
        return B200CONV_OK;
    }
    if (world > kBusMaxWorld || rank < 0 || rank >= world)
        return fail(B200CONV_ERR_INVALID, "b200conv_attach_bus: bad rank / world (at most 16 engines)");
    for (int i = 0; i < world; ++i)
        if (!peer_buffers[i]) return fail(B200CONV_ERR_INVALID, "b200conv_attach_bus: null peer buffer");
    e->bus_world = world;
    e->bus_rank = rank;
    e->bus_epoch = 0;
    for (int i = 0; i < world; ++i) e->bus_peers[i] = peer_buffers[i];
    *e->d_bus_err = 0;
    return B200CONV_OK;
}

int b200conv_bus_trace(b200conv_engine* e, uint64_t* host_stamps, int count) {
    if (!e || !host_stamps || count < 1) return fail(B200CONV_ERR_INVALID, "b200conv_bus_trace: bad argument");
    if (!e->d_bus_trace) return fail(B200CONV_ERR_STATE, "b200conv_bus_trace: create the engine with B200CONV_BUS_TRACE=1 in the environment");
    ENGINE_DEVICE(e->cfg.device);
    CU_TRY(cudaDeviceSynchronize());
    count = std::min(count, kBusTraceLen);
    // rows = the last `count` epochs, oldest first
    std::vector<unsigned long long> all(static_cast<size_t>(2) * kBusTraceLen);
    CU_TRY(cudaMemcpy(all.data(), e->d_bus_trace, all.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    for (int i = 0; i < count; ++i) {
        const uint32_t ep = e->bus_epoch - static_cast<uint32_t>(count - 1 - i);
        host_stamps[2 * i] = all[(ep % kBusTraceLen) * 2];
        host_stamps[2 * i + 1] = all[(ep % kBusTraceLen) * 2 + 1];
    }
    return B200CONV_OK;
}

int b200conv_tc_trace(b200conv_engine* e, uint64_t* host_stamps, int max_ctas) {
    if (!e || !host_stamps || max_ctas < 1) return fail(B200CONV_ERR_INVALID, "b200conv_tc_trace: bad argument");
    if (e->impl != B200CONV_ALGO_DIRECT_TC || !e->tc.trace)
        return fail(B200CONV_ERR_STATE, "b200conv_tc_trace: tensor-core direct engine created with B200CONV_TC_TRACE=1 only");
    ENGINE_DEVICE(e->cfg.device);
    CU_TRY(cudaDeviceSynchronize());
    const int n = std::min(max_ctas, e->tc.grid);
    CU_TRY(cudaMemcpy(host_stamps, e->tc.trace, static_cast<size_t>(n) * kTcTraceSlots * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return n;
}

int b200conv_bus_status(b200conv_engine* e) {
    if (!e) return fail(B200CONV_ERR_INVALID, "b200conv_bus_status: null engine");
    ENGINE_DEVICE(e->cfg.device);
    if (e->has_last_stream) CU_TRY(cudaStreamSynchronize(e->last_stream));
    if (*static_cast<volatile uint32_t*>(e->d_bus_err)) {
        *e->d_bus_err = 0;  // reported once; later blocks start clean
        return fail(B200CONV_ERR_CUDA, "bus all-reduce: a peer engine did not signal within the spin bound "
                                       "(the bus of that block is incomplete)");
    }
    return B200CONV_OK;
}

int b200conv_query(b200conv_engine* e, b200conv_info* info) {
    if (!e || !info) return fail(B200CONV_ERR_INVALID, "b200conv_query: null argument");
    std::memset(info, 0, sizeof(*info));
    const uint64_t T = e->T, B = e->B, L = e->L;
    info->macs_per_block = T * B * L;
    info->device_bytes = e->device_bytes;
    info->blocks_processed = e->blocks;
    info->kernel_launches = e->launches;
    info->sm_count = e->sm_count;
    info->stage_count = 3;
    info->dominant_stage = 1;
    info->stage_calls = e->stage_calls;
    for (int i = 0; i < 4; ++i) info->stage_ms[i] = e->stage_ms[i];
    if (e->impl == B200CONV_ALGO_DIRECT_TC) {
        info->flops_per_block = 2 * T * B * L;  // algorithmic; the tensor cores issue 3x (TF32 split) on L padded to 128
        info->alg_bytes_per_block = 0;
        info->partitions = e->tc.g.NGRP;
        info->fft_size = 0;
        info->kernels_per_block = e->tc.nsub + (e->strip_ops ? 2 : 0);
        info->stage_count = e->strip_ops ? 2 : 1;
        info->dominant_stage = 0;
        std::snprintf(info->stage_name[0], 24, "tc_toeplitz");
        std::snprintf(info->stage_name[1], 24, "strip+mix");
    } else if (e->impl == B200CONV_ALGO_DIRECT) {
        info->flops_per_block = 2 * T * B * L;
        info->alg_bytes_per_block = 0;
        info->partitions = e->dir.MS;
        info->fft_size = 0;
        info->kernels_per_block = e->strip_ops ? 4 : 2;
        info->stage_count = 2;
        info->dominant_stage = 0;
        std::snprintf(info->stage_name[0], 24, "fir_direct");
        std::snprintf(info->stage_name[1], 24, e->strip_ops ? "finish+strip+mix" : "finish+mix+append");
    } else {
        const uint64_t P = e->up.P;
        info->flops_per_block = 8 * T * P * (B + 1);
        info->alg_bytes_per_block = T * 16 * P * (B + 1);
        info->partitions = e->up.P;
        info->fft_size = 2 * e->B;
        if (e->up.fused) {
            const bool tree = e->bus_in_kernel_single;
            info->kernels_per_block = tree ? 1 : 2;
            info->stage_count = tree ? 1 : 2;
            info->dominant_stage = 0;
            std::snprintf(info->stage_name[0], 24, "upols_fused");
            std::snprintf(info->stage_name[1], 24, "mix");
        } else {
            info->kernels_per_block = 3;
            std::snprintf(info->stage_name[0], 24, "rfft_fwd");
            std::snprintf(info->stage_name[1], 24, "fdl_mac");
            std::snprintf(info->stage_name[2], 24, "irfft_ols+mix");
        }
    }
    return B200CONV_OK;
}

int b200conv_rfft(const float* d_in, void* d_out, int count, int n, void* stream) {
    if (!d_in || !d_out || count < 1) return fail(B200CONV_ERR_INVALID, "b200conv_rfft: bad argument");
    if (!is_pow2(static_cast<uint32_t>(n)) || n < 32 || n > 8192)
        return fail(B200CONV_ERR_INVALID, "b200conv_rfft: n must be a power of two in [32, 8192]");
    RfftParams p{};
    const int M = n / 2;
    p.first = d_in;  // window = [first half | second half] of each row
    p.first_stride = n;
    p.second = d_in + M;
    p.second_stride = n;
    p.out = static_cast<float2*>(d_out);
    p.out_stride = M + 1;
    p.prev_out = nullptr;
    p.count = count;
    p.M = M;
    p.logM = ilog2(static_cast<uint32_t>(M));
    p.scale = 1.0f;
    p.unpacked = 1;
    CU_TRY(launch_rfft_fwd(p, static_cast<cudaStream_t>(stream)));
    return B200CONV_OK;
}

int b200conv_set_profiling(b200conv_engine* e, int on) {
    if (!e) return fail(B200CONV_ERR_INVALID, "b200conv_set_profiling: null engine");
    e->profiling = (on != 0);
    for (float& v : e->stage_ms) v = 0.0f;
    e->stage_calls = 0;
    return B200CONV_OK;
}

}  // extern "C"
