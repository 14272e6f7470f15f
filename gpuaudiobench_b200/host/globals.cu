#include "globals.cuh"

#include <algorithm>
#include <fstream>
#include <iostream>
#include <limits>
#include <sstream>

int FS = 48000;
int NTRACKS = 128;
int BUFSIZE = 512;
int NRUNS = 100;
std::string OUTPUT_FILE;
bool JSON_OUTPUT = false;
int IR_LEN = 0;
int WARMUP_RUNS = 3;
int NGPUS = 1;
bool STREAM_MODE = false;
bool DAWSIM = false;
bool DAWSIM_SLEEP = false;
double DAWSIM_JITTER_US = 0.0;

LatencySummary summarizeLatencies(const std::vector<float>& lat) {
    LatencySummary s{};
    s.count = lat.size();
    s.threshold_ms = 1000.0f * BUFSIZE / FS;
    if (lat.empty()) return s;
    std::vector<float> sorted(lat);
    std::sort(sorted.begin(), sorted.end());
    float total = 0.0f;
    for (float v : lat) total += v;
    s.min_ms = sorted.front();
    s.max_ms = sorted.back();
    s.avg_ms = total / static_cast<float>(lat.size());
    // the reference indexes the sorted vector with size*q (a double), i.e. floor(size*q)
    auto rank = [&](double q) { return sorted[std::min(sorted.size() - 1, static_cast<size_t>(sorted.size() * q))]; };
    s.p50_ms = rank(0.50);
    s.p95_ms = rank(0.95);
    s.p99_ms = rank(0.99);
    s.meets_deadline = (s.p99_ms <= s.threshold_ms);
    return s;
}

void writeVectorToFile(const std::vector<float>& vec, const std::string& filename) {
    std::ofstream out(filename);
    for (float v : vec) out << v << std::endl;
}

void printVectorStats(const std::vector<float>& vec) {
    const LatencySummary s = summarizeLatencies(vec);
    std::cout << "Min: " << s.min_ms << " Max: " << s.max_ms << " Avg: " << s.avg_ms << std::endl;
    std::cout << "p50: " << s.p50_ms << " p95: " << s.p95_ms << " p99: " << s.p99_ms << std::endl;
    std::cout << "Latency threshold (" << FS << "Hz): " << s.threshold_ms << " ms" << std::endl;
    const char* worst = s.p50_ms > s.threshold_ms ? "p50" : s.p95_ms > s.threshold_ms ? "p95" : s.p99_ms > s.threshold_ms ? "p99" : nullptr;
    if (worst)
        std::cout << "WARNING: " << worst << " exceeds threshold" << std::endl;
    else
        std::cout << "OK: Measured latencies within threshold. Please consider a margin of safety." << std::endl;
}

void writeCSVResults(const std::vector<float>& vec, const std::string& benchmarkName, const std::string& filename) {
    if (filename.empty()) return;
    const LatencySummary s = summarizeLatencies(vec);
    const bool fresh = !std::ifstream(filename).good();
    std::ofstream out(filename, std::ios::app);
    if (fresh)
        out << "benchmark,fs,bufferSize,nTracks,nRuns,min_ms,max_ms,avg_ms,p50_ms,p95_ms,p99_ms,threshold_ms,meets_deadline\n";
    out << benchmarkName << ',' << FS << ',' << BUFSIZE << ',' << NTRACKS << ',' << vec.size() << ',' << s.min_ms << ','
        << s.max_ms << ',' << s.avg_ms << ',' << s.p50_ms << ',' << s.p95_ms << ',' << s.p99_ms << ',' << s.threshold_ms
        << ',' << (s.meets_deadline ? "true" : "false") << "\n";
    out.close();
    std::cout << "Results saved to: " << filename << std::endl;
}

std::string generateJSONResults(const std::vector<float>& vec, const std::string& benchmarkName) {
    const LatencySummary s = summarizeLatencies(vec);
    auto num = [](float v) { return std::to_string(v); };  // 6 decimals, like the reference's std::to_string
    std::ostringstream js;
    js << "{\n"
       << "  \"benchmark\": \"" << benchmarkName << "\",\n"
       << "  \"configuration\": {\n"
       << "    \"fs\": " << FS << ",\n"
       << "    \"bufferSize\": " << BUFSIZE << ",\n"
       << "    \"nTracks\": " << NTRACKS << ",\n"
       << "    \"nRuns\": " << static_cast<int>(vec.size()) << "\n"
       << "  },\n"
       << "  \"statistics\": {\n"
       << "    \"min_ms\": " << num(s.min_ms) << ",\n"
       << "    \"max_ms\": " << num(s.max_ms) << ",\n"
       << "    \"avg_ms\": " << num(s.avg_ms) << ",\n"
       << "    \"p50_ms\": " << num(s.p50_ms) << ",\n"
       << "    \"p95_ms\": " << num(s.p95_ms) << ",\n"
       << "    \"p99_ms\": " << num(s.p99_ms) << "\n"
       << "  },\n"
       << "  \"deadline\": {\n"
       << "    \"threshold_ms\": " << num(s.threshold_ms) << ",\n"
       << "    \"meets_deadline\": " << (s.meets_deadline ? "true" : "false") << "\n"
       << "  }\n"
       << "}\n";
    return js.str();
}

void writeJSONResults(const std::vector<float>& vec, const std::string& benchmarkName, const std::string& filename) {
    const std::string js = generateJSONResults(vec, benchmarkName);
    if (filename.empty()) {
        std::cout << js << std::endl;
        return;
    }
    std::ofstream(filename) << js;
    std::cout << "JSON results saved to: " << filename << std::endl;
}
