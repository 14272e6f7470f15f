#include "registry.cuh"

#include <cstdio>
#include <functional>

#include "bench_conv1d.cuh"
#include "bench_conv1d_accel.cuh"
#include "bench_fft.cuh"
#include "bench_strip.cuh"

namespace {
struct Entry {
    const char* name;
    std::function<std::unique_ptr<GPUABenchmark>()> make;
};

// Constructor defaults read BUFSIZE/NTRACKS at call time, after argument parsing — that is how the
// reference's CLI overrides reach a plugin (bench_conv1d.cuh:17); --irLen is this build's addition.
const std::vector<Entry>& registry() {
    static const std::vector<Entry> entries = {
        {"Conv1D", [] { return std::make_unique<Conv1DBenchmark>(IR_LEN > 0 ? IR_LEN : Conv1DBenchmark::DEFAULT_IR_LEN); }},
        {"Conv1D_accel",
         [] { return std::make_unique<Conv1DAccelBenchmark>(IR_LEN > 0 ? IR_LEN : Conv1DAccelBenchmark::DEFAULT_IR_LEN); }},
        {"FFT1D", [] { return std::make_unique<FFTBenchmark>(); }},  // first step beyond the conv path
        // the channel-strip ops next to the mix bus, registry names of cuda/main.cu:85-93
        {"gain", [] { return std::make_unique<GainBenchmark>(); }},
        {"GainStats", [] { return std::make_unique<GainStatsBenchmark>(); }},
        {"IIRFilter", [] { return std::make_unique<IIRBenchmark>(); }},
    };
    return entries;
}
}  // namespace

std::vector<std::string> listBenchmarks() {
    std::vector<std::string> names;
    for (const Entry& e : registry()) names.emplace_back(e.name);
    return names;
}

std::unique_ptr<GPUABenchmark> createBenchmark(const std::string& name) {
    for (const Entry& e : registry())
        if (name == e.name) return e.make();
    return nullptr;
}

int runSelectedBenchmark(std::unique_ptr<GPUABenchmark> benchmark, const std::string& benchmarkName) {
    if (!benchmark) {
        std::printf("Unknown benchmark: %s\n", benchmarkName.c_str());
        return -1;
    }
    int status = -1;
    try {
        std::printf("Setting up %s benchmark...\n", benchmarkName.c_str());
        benchmark->setupBenchmark();

        std::printf("Running %s benchmark (%d iterations with %d warmup)...\n", benchmarkName.c_str(), NRUNS, WARMUP_RUNS);
        auto result = benchmark->runBenchmark(NRUNS, WARMUP_RUNS);

        std::printf("Validating %s benchmark results...\n", benchmarkName.c_str());
        GPUABenchmark::ValidationData validation;
        benchmark->validate(validation);
        status = static_cast<int>(validation.status);
        if (validation.status != GPUABenchmark::ValidationStatus::SUCCESS)
            std::printf("Validation failed for %s:\n", benchmarkName.c_str());
        else
            std::printf("Validation passed for %s\n", benchmarkName.c_str());
        for (const auto& msg : validation.messages) std::printf("  %s\n", msg.c_str());

        if (JSON_OUTPUT) {
            writeJSONResults(result.latencies, benchmarkName, OUTPUT_FILE);
        } else {
            benchmark->printResults(result);
            benchmark->writeResults(result);
            if (!OUTPUT_FILE.empty()) writeCSVResults(result.latencies, benchmarkName, OUTPUT_FILE);
        }
        std::printf("%s benchmark completed successfully!\n", benchmarkName.c_str());
    } catch (const std::exception& e) {
        std::printf("Benchmark %s failed: %s\n", benchmarkName.c_str(), e.what());
    }
    return status;
}
