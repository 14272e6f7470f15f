#pragma once
// Benchmark registry of the gpubench CLI: name -> factory (reference cuda/main.cu:68-115).
// The two plugins of the convolution path plus FFT1D (the first "next" row of SURVEY.md §8f) exist
// in this build; the other 14 names of the reference are out of scope and are reported as unknown.
#include <memory>
#include <string>
#include <vector>

#include "bench_base.cuh"

std::vector<std::string> listBenchmarks();
std::unique_ptr<GPUABenchmark> createBenchmark(const std::string& name);
// setup -> runBenchmark(NRUNS, WARMUP_RUNS) -> validate -> report, exceptions caught and printed
// (reference runSelectedBenchmark, main.cu:117-164).  Returns the validation status as an int.
int runSelectedBenchmark(std::unique_ptr<GPUABenchmark> benchmark, const std::string& benchmarkName);
