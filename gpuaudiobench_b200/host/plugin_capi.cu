// plugin_capi.cu — include/gpubench_plugin.h over the plugin classes (libgpubench_b200.so).
#include "gpubench_plugin.h"

#include <chrono>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "bench_conv1d.cuh"
#include "bench_conv1d_accel.cuh"
#include "bench_fft.cuh"
#include "bench_strip.cuh"
#include "conv_common.cuh"
#include "registry.cuh"

struct gpubench_plugin {
    std::unique_ptr<GPUABenchmark> bench;
    Conv1DBenchmark* direct = nullptr;
    Conv1DAccelBenchmark* accel = nullptr;
    FFTBenchmark* fft = nullptr;
    ChannelStripBenchmark* strip = nullptr;
};

namespace {
thread_local std::string g_err;

template <typename F> int guarded(F&& body) {
    try {
        body();
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}
}  // namespace

extern "C" {

const char* gpubench_last_error(void) { return g_err.c_str(); }

void gpubench_set_globals(int fs, int nruns, int stream_mode) {
    if (fs > 0) FS = fs;
    if (nruns > 0) NRUNS = nruns;
    STREAM_MODE = (stream_mode != 0);
}

void gpubench_set_ngpus(int n) { NGPUS = n > 0 ? n : 1; }

void gpubench_set_dawsim(int enable, int sleep_mode, double jitter_us) {
    DAWSIM = (enable != 0);
    DAWSIM_SLEEP = (sleep_mode != 0);
    DAWSIM_JITTER_US = jitter_us;
}

int gpubench_dawsim_probe(double period_s, int sleep_mode, double jitter_us, int n, double* wake_times_s) {
    BenchmarkUtils::DAWSimulator sim(period_s, sleep_mode ? BenchmarkUtils::DAWSimulator::Mode::SLEEP
                                                           : BenchmarkUtils::DAWSimulator::Mode::SPIN, jitter_us * 1e-6);
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < n; ++i) {
        sim.wait();
        wake_times_s[i] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    return 0;
}

gpubench_plugin* gpubench_create(const char* name, int ir_len, int buffer_size, int track_count) {
    if (!name) return nullptr;
    BUFSIZE = buffer_size;
    NTRACKS = track_count;
    IR_LEN = ir_len > 0 ? ir_len : 0;
    auto bench = createBenchmark(name);
    if (!bench) {
        g_err = std::string("Unknown benchmark: ") + name;
        return nullptr;
    }
    auto* p = new gpubench_plugin();
    p->direct = dynamic_cast<Conv1DBenchmark*>(bench.get());
    p->accel = dynamic_cast<Conv1DAccelBenchmark*>(bench.get());
    p->fft = dynamic_cast<FFTBenchmark*>(bench.get());
    p->strip = dynamic_cast<ChannelStripBenchmark*>(bench.get());
    p->bench = std::move(bench);
    return p;
}

void gpubench_destroy(gpubench_plugin* p) { delete p; }

int gpubench_setup(gpubench_plugin* p) {
    return guarded([&] { p->bench->setupBenchmark(); });
}

int gpubench_iterate(gpubench_plugin* p) {
    return guarded([&] { p->bench->performBenchmarkIteration(); });
}

int gpubench_run(gpubench_plugin* p, int iterations, int warmup, float* wall_ms, float* gpu_ms) {
    return guarded([&] {
        auto r = p->bench->runBenchmark(iterations, warmup);
        for (int i = 0; i < iterations; ++i) {
            if (wall_ms) wall_ms[i] = r.latencies[i];
            if (gpu_ms) gpu_ms[i] = i < static_cast<int>(r.gpu_latencies.size()) ? r.gpu_latencies[i] : 0.0f;
        }
    });
}

int gpubench_validate(gpubench_plugin* p, gpubench_validation* out, char* messages, size_t cap) {
    return guarded([&] {
        GPUABenchmark::ValidationData v;
        p->bench->validate(v);
        if (out) {
            out->status = static_cast<int>(v.status);
            out->max_error = v.max_error;
            out->mean_error = v.mean_error;
            if (p->fft) {
                out->snr_db = p->fft->lastSnrDb();
                out->max_abs_err = v.max_error;
                out->ref_peak = 0.0;
            } else {
                const float* ref = p->strip ? p->strip->cpuReference()
                                   : p->direct ? p->direct->cpuReference() : p->accel->cpuReference();
                const ConvCommon::Accuracy a = ConvCommon::measureAccuracy(p->bench->hostOutput(), ref, p->bench->getTotalElements());
                out->snr_db = a.snr_db;
                out->max_abs_err = a.max_abs_err;
                out->ref_peak = a.ref_peak;
            }
        }
        if (messages && cap) {
            std::string joined;
            for (const auto& m : v.messages) joined += m + "\n";
            std::strncpy(messages, joined.c_str(), cap - 1);
            messages[cap - 1] = '\0';
        }
    });
}

const float* gpubench_host_input(gpubench_plugin* p) { return p->bench->hostInput(); }
const float* gpubench_host_ir(gpubench_plugin* p) {
    return p->direct ? p->direct->hostIR() : p->accel ? p->accel->hostIR() : nullptr;
}
const float* gpubench_host_output(gpubench_plugin* p) { return p->bench->hostOutput(); }
const float* gpubench_cpu_reference(gpubench_plugin* p) {
    return p->strip ? p->strip->cpuReference() : p->direct ? p->direct->cpuReference() : p->accel ? p->accel->cpuReference() : nullptr;
}

const float* gpubench_strip_stats(gpubench_plugin* p, int cpu) {
    return !p->strip ? nullptr : cpu ? p->strip->cpuStats() : p->strip->hostStats();
}
const float* gpubench_strip_state(gpubench_plugin* p, int cpu) {
    return !p->strip ? nullptr : cpu ? p->strip->cpuState() : p->strip->hostState();
}
int gpubench_strip_coefficients(gpubench_plugin* p, float out5[5]) {
    if (!p->strip) return 1;
    const IIRCoefficients& c = p->strip->coefficients();
    out5[0] = c.b0; out5[1] = c.b1; out5[2] = c.b2; out5[3] = c.a1; out5[4] = c.a2;
    return 0;
}
int gpubench_strip_bit_exact(gpubench_plugin* p) { return p->strip && p->strip->lastRunBitExact() ? 1 : 0; }

const float* gpubench_fft_input(gpubench_plugin* p) { return p->fft ? p->fft->hostInputFFT() : nullptr; }
const float* gpubench_fft_output(gpubench_plugin* p) { return p->fft ? reinterpret_cast<const float*>(p->fft->hostOutputFFT()) : nullptr; }
const float* gpubench_fft_reference(gpubench_plugin* p) { return p->fft ? reinterpret_cast<const float*>(p->fft->cpuReferenceFFT()) : nullptr; }

int gpubench_json_results(const float* lat, size_t n, const char* name, int fs, int bufsize, int ntracks, char* out, size_t cap) {
    FS = fs;
    BUFSIZE = bufsize;
    NTRACKS = ntracks;
    const std::string js = generateJSONResults(std::vector<float>(lat, lat + n), name);
    if (js.size() + 1 > cap) return -1;
    std::memcpy(out, js.c_str(), js.size() + 1);
    return static_cast<int>(js.size());
}

int gpubench_statistics(const float* lat, size_t n, float out8[8]) {
    const BenchmarkUtils::Statistics s = BenchmarkUtils::calculateStatistics(std::vector<float>(lat, lat + n));
    const float v[8] = {s.mean, s.median, s.std_dev, s.min_val, s.max_val, s.p95, s.p99, static_cast<float>(s.count)};
    std::memcpy(out8, v, sizeof(v));
    return 0;
}

}  // extern "C"
