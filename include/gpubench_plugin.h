/* gpubench_plugin.h — C binding of the benchmark plugin surface (libgpubench_b200.so).
 *
 * The reference's plugin API is a C++ class (GPUABenchmark, cuda/bench_base.cuh:18-139) driven by
 * main.cu:117-164: setupBenchmark() -> runBenchmark(NRUNS, 3) -> validate() -> report.  This header
 * exposes that same lifecycle over the re-created classes in gpuaudiobench_b200/host/ so that the
 * Python parity tests (ctypes) exercise exactly what the gpubench CLI runs.  All functions return
 * 0 on success, non-zero on failure with a message in gpubench_last_error().
 */
#ifndef GPUBENCH_PLUGIN_H_
#define GPUBENCH_PLUGIN_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gpubench_plugin gpubench_plugin;

typedef struct gpubench_validation {
    int status;        /* GPUABenchmark::ValidationStatus: 0 success, 1 failure, -1 fatal */
    float max_error;   /* the reference's metric: abs (Conv1D) or relative (Conv1D_accel)    */
    float mean_error;
    double snr_db;     /* stated tolerance metrics vs the plugin's CPU reference             */
    double max_abs_err;
    double ref_peak;
} gpubench_validation;

const char* gpubench_last_error(void);

/* The CLI globals (cuda/globals.cu:4-9): --fs, --nRuns; stream != 0 is --mode stream. */
void gpubench_set_globals(int fs, int nruns, int stream_mode);
/* --nGpus: shard the tracks of subsequently created plugins over n GPUs (b200conv_group_*). */
void gpubench_set_ngpus(int n);

/* DAW-style pacing of runBenchmark (--dawsim, --dawsim-mode, --dawsim-jitter-us); enable = 0 turns it off. */
void gpubench_set_dawsim(int enable, int sleep_mode, double jitter_us);
/* Drive a DAWSimulator alone: call wait() n times, return the wake-up times in seconds since the first call. */
int gpubench_dawsim_probe(double period_s, int sleep_mode, double jitter_us, int n, double* wake_times_s);

/* createBenchmark(name) (main.cu:105-115) with explicit sizes; name is "Conv1D" or "Conv1D_accel";
 * ir_len <= 0 selects the plugin default (1024 / 512). NULL on unknown name. */
gpubench_plugin* gpubench_create(const char* name, int ir_len, int buffer_size, int track_count);
void gpubench_destroy(gpubench_plugin* p);

int gpubench_setup(gpubench_plugin* p);                                         /* setupBenchmark()             */
int gpubench_iterate(gpubench_plugin* p);                                       /* performBenchmarkIteration()  */
int gpubench_run(gpubench_plugin* p, int iterations, int warmup,                /* runBenchmark(): wall + GPU   */
                 float* wall_ms, float* gpu_ms);                                /* latency arrays [iterations]  */
int gpubench_validate(gpubench_plugin* p, gpubench_validation* out,             /* validate()                   */
                      char* messages, size_t messages_cap);                     /* '\n'-joined messages         */

/* Views of the plugin's host buffers (valid until destroy): input [T][B], IR [T][L], the last
 * iteration's output and the CPU reference ([T][B] track-major for Conv1D, [B][T] for Conv1D_accel). */
const float* gpubench_host_input(gpubench_plugin* p);
const float* gpubench_host_ir(gpubench_plugin* p);
const float* gpubench_host_output(gpubench_plugin* p);
const float* gpubench_cpu_reference(gpubench_plugin* p);

/* FFT1D plugin views: input float [T][1024]; output / float-DFT reference as interleaved complex
 * float [T][513][2]; gpubench_validate's snr_db is then the SNR against the plugin's fp64 DFT. */
const float* gpubench_fft_input(gpubench_plugin* p);
const float* gpubench_fft_output(gpubench_plugin* p);
const float* gpubench_fft_reference(gpubench_plugin* p);

/* gain / GainStats / IIRFilter plugins (channel strip).  cpu == 0: what the device
 * produced in the last iteration; cpu != 0: the plugin's CPU loop (valid after gpubench_validate).
 * stats: float [T][2] mean, max (cuda/bench_gainstats.cu:29-30); state: float [T][2] z1, z2
 * (cuda/bench_iir.cu:42-43).  NULL for other plugins. */
const float* gpubench_strip_stats(gpubench_plugin* p, int cpu);
const float* gpubench_strip_state(gpubench_plugin* p, int cpu);
int gpubench_strip_coefficients(gpubench_plugin* p, float out5[5]); /* b0 b1 b2 a1 a2 (cuda/bench_iir.cu:205-228) */
int gpubench_strip_bit_exact(gpubench_plugin* p);                   /* 1 if the last validate() found identical bits */

/* The reference's result writers (globals.cu:69-182) for a latency vector. */
int gpubench_json_results(const float* latencies_ms, size_t n, const char* name, int fs, int bufsize, int ntracks,
                          char* out, size_t cap);
int gpubench_statistics(const float* latencies_ms, size_t n, float out8[8]); /* mean median std min max p95 p99 count */

#ifdef __cplusplus
}
#endif
#endif /* GPUBENCH_PLUGIN_H_ */
