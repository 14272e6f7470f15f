// gpubench — command-line entry of the benchmark suite, convolution path only.
//
// Same flags, defaults, messages and exit codes as the reference's cuda/main.cu:236-328
// (--help --list --json --benchmark --fs --bufferSize --nTracks --nRuns --outputfile; unknown
// arguments warn; exit 1 only for a missing flag value, an unknown benchmark or no CUDA device;
// a failed validation is printed, not returned).  Added here: --irLen, --warmup, --mode, and the
// lower-case spellings the Metal port uses (--buffersize --ntracks --nruns, main.swift:48-163).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "globals.cuh"
#include "registry.cuh"

static void printBenchmarkList() {
    std::printf("Available benchmarks:\n");
    for (const std::string& name : listBenchmarks()) std::printf("%s\n", name.c_str());
}

static void printHelp() {
    std::printf("CUDA GPU Audio Benchmark Suite — B200 convolution engine\n");
    std::printf("=========================================================\n");
    std::printf("Real-time GPGPU audio processing benchmarks (per-track FIR convolution path)\n\n");
    std::printf("Usage: gpubench [options]\n\n");
    std::printf("Options:\n");
    std::printf("  --help              Print this help message\n");
    std::printf("  --list              List all available benchmarks\n");
    std::printf("  --benchmark [name]  Run specific benchmark (see list below)\n");
    std::printf("  --fs [rate]         Set sampling rate (default: 48000)\n");
    std::printf("  --bufferSize [size] Set buffer size (default: 512)\n");
    std::printf("  --nTracks [count]   Set number of tracks (default: 128)\n");
    std::printf("  --nRuns [count]     Set number of iterations (default: 100)\n");
    std::printf("  --outputfile [file] Save results to CSV file\n");
    std::printf("  --json              Output results in JSON format\n");
    std::printf("  --irLen [taps]      Impulse response length (default: 1024 Conv1D, 512 Conv1D_accel)\n");
    std::printf("  --warmup [count]    Warm-up iterations before timing (default: 3)\n");
    std::printf("  --nGpus [count]     Shard the tracks over this many GPUs of the box (default: 1)\n");
    std::printf("  --mode [stateless|stream]  stateless re-submits one buffer (reference behaviour, default);\n");
    std::printf("                      stream advances the convolution state every iteration\n");
    std::printf("  --dawsim            Pace iterations at the buffer period bufferSize/fs (DAW-style submission)\n");
    std::printf("  --dawsim-mode [spin|sleep]   how to wait for the next period (default: spin)\n");
    std::printf("  --dawsim-jitter-us [us]      uniform +/- jitter on every wake-up (default: 0)\n\n");
    std::printf("Digital Signal Processing:\n");
    std::printf("  Conv1D           - 1D convolution (direct form, FP32 FMA bound)\n");
    std::printf("  Conv1D_accel     - Accelerated 1D convolution (partitioned overlap-save FFT, HBM bound)\n");
    std::printf("  FFT1D            - 1D Fast Fourier Transform (1024-point R2C per track, shared-memory Stockham)\n\n");
    std::printf("Examples:\n");
    std::printf("  gpubench --benchmark Conv1D --nTracks 128 --irLen 16384\n");
    std::printf("  gpubench --benchmark Conv1D_accel --bufferSize 256 --nTracks 1024 --irLen 65536 --json\n\n");
}

namespace {
struct IntFlag {
    const char* name;
    const char* alias;
    int* target;
    const char* echo;  // message the reference prints after setting (or nullptr)
};
}  // namespace

int main(int argc, char** argv) {
    std::printf("GPGPU Audio Benchmark\n");
    std::string whichBenchmark = "Conv1D";  // the reference defaults to RndMemRead, which is out of scope here

    const IntFlag int_flags[] = {
        {"--fs", nullptr, &FS, nullptr},
        {"--bufferSize", "--buffersize", &BUFSIZE, "Buffer size set to: %d\n"},
        {"--nTracks", "--ntracks", &NTRACKS, "Number of tracks set to: %d\n"},
        {"--nRuns", "--nruns", &NRUNS, nullptr},
        {"--irLen", "--irlen", &IR_LEN, "IR length set to: %d\n"},
        {"--warmup", nullptr, &WARMUP_RUNS, nullptr},
        {"--nGpus", "--ngpus", &NGPUS, "Number of GPUs set to: %d\n"},
    };

    for (int i = 1; i < argc; ++i) {
        const char* arg = argv[i];
        const bool has_value = i + 1 < argc;
        auto need_value = [&](const char* flag) {
            if (!has_value) std::printf("Error: %s requires an argument\n", flag);
            return has_value;
        };
        if (!std::strcmp(arg, "--help")) {
            printHelp();
            return 0;
        }
        if (!std::strcmp(arg, "--list")) {
            printBenchmarkList();
            return 0;
        }
        if (!std::strcmp(arg, "--json")) {
            JSON_OUTPUT = true;
            continue;
        }
        if (!std::strcmp(arg, "--benchmark")) {
            if (!need_value("--benchmark")) return 1;
            whichBenchmark = argv[++i];
            continue;
        }
        if (!std::strcmp(arg, "--outputfile")) {
            if (!need_value("--outputfile")) return 1;
            OUTPUT_FILE = argv[++i];
            std::printf("Output file set to: %s\n", OUTPUT_FILE.c_str());
            continue;
        }
        if (!std::strcmp(arg, "--dawsim")) {
            DAWSIM = true;
            continue;
        }
        if (!std::strcmp(arg, "--dawsim-mode")) {
            if (!need_value("--dawsim-mode")) return 1;
            DAWSIM_SLEEP = !std::strcmp(argv[++i], "sleep");
            continue;
        }
        if (!std::strcmp(arg, "--dawsim-jitter-us")) {
            if (!need_value("--dawsim-jitter-us")) return 1;
            DAWSIM_JITTER_US = std::atof(argv[++i]);
            continue;
        }
        if (!std::strcmp(arg, "--mode")) {
            if (!need_value("--mode")) return 1;
            STREAM_MODE = !std::strcmp(argv[++i], "stream");
            continue;
        }
        bool matched = false;
        for (const IntFlag& f : int_flags) {
            if (!std::strcmp(arg, f.name) || (f.alias && !std::strcmp(arg, f.alias))) {
                if (!need_value(f.name)) return 1;
                *f.target = std::atoi(argv[++i]);
                if (f.echo) std::printf(f.echo, *f.target);
                matched = true;
                break;
            }
        }
        if (!matched) std::printf("Warning: Unparsed argument: %s\n", arg);
    }

    int deviceCount = 0;
    const cudaError_t err = cudaGetDeviceCount(&deviceCount);
    if (err != cudaSuccess) {
        std::printf("Failed to get CUDA device count: %s\n", cudaGetErrorString(err));
        return 1;
    }
    std::printf("Found %d CUDA device(s)\n", deviceCount);

    if (auto instance = createBenchmark(whichBenchmark)) {
        std::printf("Running %s benchmark...\n", whichBenchmark.c_str());
        runSelectedBenchmark(std::move(instance), whichBenchmark);
        std::printf("Done\n");
        return 0;
    }
    std::printf("Error: Unknown benchmark '%s'\n", whichBenchmark.c_str());
    std::printf("This benchmark is not registered. Check createBenchmark() implementation.\n");
    std::printf("Use --list to see available benchmarks.\n");
    return 1;
}
