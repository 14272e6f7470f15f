#pragma once
// Command-line globals and the CSV / JSON result writers of the gpubench CLI.
// Same names, defaults and output formats as the reference (cuda/globals.cuh:19-24, globals.cu:4-9,
// CSV header globals.cu:101, JSON keys globals.cu:159-180) so table-generating scripts keep working.
#include <string>
#include <vector>

extern int FS;                   // --fs          (48000)
extern int NTRACKS;              // --nTracks     (128)
extern int BUFSIZE;              // --bufferSize  (512)
extern int NRUNS;                // --nRuns       (100)
extern std::string OUTPUT_FILE;  // --outputfile  ("")
extern bool JSON_OUTPUT;         // --json
// Additions of this build (the reference has no CLI knob for them):
extern int IR_LEN;               // --irLen   (0 = the plugin's default: 1024 direct, 512 accel)
extern int WARMUP_RUNS;          // --warmup  (3, the value main.cu:130 hard-codes)
extern int NGPUS;                // --nGpus: shard the tracks over this many GPUs (b200conv_group_*; default 1)
extern bool STREAM_MODE;         // --mode stream: advance the convolution state every iteration
// DAW-style periodic submission.  The reference's CUDA port only has unused compile-time stubs
// (cuda/globals.cuh:28-30 ENABLE_DAWSIM_SLEEP / SLEEP_MS / ENABLE_DAWSIM_SPIN); the behaviour follows
// its Metal port (metal-swift/.../Core/BenchmarkUtilities.swift:140-178, flags main.swift:110-132).
extern bool DAWSIM;              // --dawsim: wait for the next buffer period after every iteration
extern bool DAWSIM_SLEEP;        // --dawsim-mode sleep|spin (default spin)
extern double DAWSIM_JITTER_US;  // --dawsim-jitter-us: uniform +/- jitter on each wake-up

struct LatencySummary {
    float min_ms, max_ms, avg_ms, p50_ms, p95_ms, p99_ms, threshold_ms;
    bool meets_deadline;
    size_t count;
};

// nearest-rank percentiles sorted[n*q] and deadline 1000*BUFSIZE/FS, as globals.cu:83-89
LatencySummary summarizeLatencies(const std::vector<float>& latencies_ms);

void writeVectorToFile(const std::vector<float>& vec, const std::string& filename);
void printVectorStats(const std::vector<float>& vec);
void writeCSVResults(const std::vector<float>& vec, const std::string& benchmarkName, const std::string& filename);
std::string generateJSONResults(const std::vector<float>& vec, const std::string& benchmarkName);
void writeJSONResults(const std::vector<float>& vec, const std::string& benchmarkName, const std::string& filename = "");
