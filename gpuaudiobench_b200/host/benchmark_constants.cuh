#pragma once
// Named constants of the convolution path.
// Mirrors the conv-related entries of the reference's cuda/benchmark_constants.cuh:22-25 (which
// the reference declares but then hard-codes as literals in bench_conv1d.cu:166-169); here the
// plugins actually use them.  Constants of the remaining benchmarks are out of scope.
namespace BenchmarkConstants {

// Hamming window w[k] = A0 - A1 cos(2 pi k / (L-1))
constexpr float HAMMING_WINDOW_A0 = 0.54f;
constexpr float HAMMING_WINDOW_A1 = 0.46f;
// per-track sinc cutoff: BASE + RANGE * track / trackCount  (cycles per sample)
constexpr float CONV1D_IR_BASE_FREQ = 0.1f;
constexpr float CONV1D_IR_FREQ_RANGE = 0.05f;

// Validation.  The reference's own thresholds (kept for continuity, reported alongside):
constexpr float CONV1D_REFERENCE_ABS_TOL = 1e-3f;        // cuda/bench_conv1d.cu:108
constexpr float CONV1D_ACCEL_REFERENCE_REL_TOL = 1e-3f;  // cuda/bench_conv1d_accel.cu:310
// ... and the tolerances this build states and enforces:
constexpr double CONV1D_MIN_SNR_DB = 100.0;
constexpr double CONV1D_MAX_ABS_REL_TO_PEAK = 1e-5;
constexpr double CONV1D_ACCEL_MIN_SNR_DB = 90.0;
constexpr double CONV1D_ACCEL_MAX_ABS_REL_TO_PEAK = 1e-4;

// Default impulse-response lengths of the two plugins (bench_conv1d.cuh:11, bench_conv1d_accel.cuh:11)
constexpr int CONV1D_DEFAULT_IR_LEN = 1024;
constexpr int CONV1D_ACCEL_DEFAULT_IR_LEN = 512;

// Channel-strip plugins: cuda/benchmark_constants.cuh:6-7, cuda/bench_iir.cu:162,213
constexpr float GAIN_VALUE = 2.0f;
constexpr float GAINSTATS_GAIN = 0.5f;
constexpr float IIR_NORMALIZED_CUTOFF = 0.25f;  // fc / fs
constexpr float IIR_BUTTERWORTH_Q = 0.707f;

}  // namespace BenchmarkConstants
