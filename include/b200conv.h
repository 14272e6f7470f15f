/* b200conv.h — C ABI of the B200-native multichannel convolution engine (libb200conv.so).
 *
 * This is the drop-in boundary for the one hot path of tskare/gpuaudiobench that this repository
 * replaces: per-track FIR convolution of streaming audio buffers, direct form (Conv1D) and
 * partitioned-FFT form (Conv1D_accel).  Nothing like this ABI exists in the reference — its
 * plugins call their kernels inline — so every entry point below names the reference code it
 * stands in for (paths relative to the reference root).  INTEGRATION.md shows the
 * lines a reference maintainer adds to bench_conv1d.cu / bench_conv1d_accel.cu to bind it.
 *
 * Conventions: extern "C", plain pointers and sizes, no exceptions; every call returns
 * B200CONV_OK (0) or a negative b200conv_status and leaves a message for b200conv_last_error()
 * (thread-local).  One engine per device, not thread-safe per handle.  The caller owns the I/O
 * buffers; the engine owns IR tables, input history / frequency-domain delay line and workspace.
 * All device work of b200conv_process is enqueued on the caller's cudaStream_t (passed as void*
 * so the header needs no CUDA include).  Devices: every call that takes an engine runs on the engine's
 * cfg.device (it selects it for the duration of the call) and leaves the calling thread's current device
 * as it found it, so engines on different devices can be driven from one thread; the stream passed to
 * b200conv_process must belong to the engine's device (NULL = that device's default stream).  The
 * stateless calls (b200conv_rfft, b200conv_strip_process) run on the caller's current device.
 * There is no CPU fallback: without a CUDA device b200conv_create fails with B200CONV_ERR_NO_DEVICE.
 *
 * Data layouts:
 *   input   float [T][B]  track-major                       (cuda/bench_base.cu:30-35 upload)
 *   IR      float [T][L]  track-major                       (cuda/bench_conv1d.cu:159-178)
 *   output  float [T][B]  track-major  (Conv1D,       cuda/bench_conv1d.cu:25,205)  or
 *           float [B][Tg] sample-major (Conv1D_accel, cuda/bench_conv1d_accel.cu:44,249), where
 *           Tg = total_tracks and this engine writes columns [track_offset, track_offset+T)
 *   mix bus float [2][B]  (new: no reference counterpart)
 * Streaming semantics: block m of track t continues the track's stream; after b200conv_reset the
 * first block equals the reference oracle R2 (zero history); oracle R1's cross-track bleed is
 * reproduced by b200conv_prime_history.
 */
#ifndef B200CONV_H_
#define B200CONV_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200CONV_ABI_VERSION 1u

typedef enum b200conv_status {
    B200CONV_OK = 0,
    B200CONV_ERR_INVALID = -1,     /* bad argument / unsupported size                       */
    B200CONV_ERR_CUDA = -2,        /* a CUDA runtime call failed (message has the detail)    */
    B200CONV_ERR_NO_DEVICE = -3,   /* no CUDA device, or not an sm_100 part                  */
    B200CONV_ERR_STATE = -4,       /* call order violated (e.g. process before load_ir)      */
    B200CONV_ERR_ABI = -5          /* cfg.abi_version != B200CONV_ABI_VERSION                */
} b200conv_status;

typedef enum b200conv_algo {
    B200CONV_ALGO_DIRECT = 0, /* direct-form FIR      — replaces Conv1DTextureMemoryImplKernel, cuda/bench_conv1d.cu:7-27 */
    B200CONV_ALGO_UPOLS = 1,  /* partitioned overlap-save — replaces the cuFFT pipeline, cuda/bench_conv1d_accel.cu:258-304 */
    B200CONV_ALGO_DIRECT_TC = 2 /* the same direct-form sum as ALGO_DIRECT on the tensor cores (tcgen05 kind::tf32 with a
                                   3-term hi/lo split, accumulators in TMEM): input-side Toeplitz GEMM per track and buffer,
                                   state = pending-output ring; the buffer's own samples in FP32 FMA.  block: a multiple
                                   of 128 up to 1024, or a multiple of 512 up to 8192 (streamed as sub-blocks).
                                   Replaces cuda/bench_conv1d.cu:7-27 like ALGO_DIRECT (csrc/tc_toeplitz.cu) */
} b200conv_algo;

typedef enum b200conv_layout {
    B200CONV_OUT_TRACK_MAJOR = 0,
    B200CONV_OUT_SAMPLE_MAJOR = 1
} b200conv_layout;

/* b200conv_config.flags */
#define B200CONV_FLAG_FFMA_ONLY 1u /* ALGO_DIRECT: never dispatch to the tensor-core kernel (keeps the FFMA kernel's bit-exact
                                      impulse behaviour and its 126-131 dB; the default planner picks DIRECT_TC for blocks of
                                      128, 256 ... 1024 samples and multiples of 512 beyond, once
                                      tracks*block*ir_len >= 2.5e8, ~110 dB) */

/* b200conv_process flags */
#define B200CONV_PEEK 1u /* compute this block but do not advance the stream state: repeated calls
                            are idempotent, which is what the reference's stateless iteration loop
                            (cuda/bench_base.cu:89-94 re-submitting the same h_input) needs */

typedef struct b200conv_config {
    uint32_t abi_version;  /* B200CONV_ABI_VERSION */
    int32_t device;        /* CUDA ordinal; the reference never calls cudaSetDevice (main.cu:309) */
    uint32_t tracks;       /* T: tracks owned by this engine (NTRACKS, cuda/globals.cu:5) */
    uint32_t track_offset; /* global index of local track 0 (multi-GPU track sharding; 0 on one GPU) */
    uint32_t total_tracks; /* Tg: global track count; 0 means == tracks */
    uint32_t block;        /* B: samples per buffer (BUFSIZE, cuda/globals.cu:6) */
    uint32_t ir_len;       /* L: taps (ir_length_, cuda/bench_conv1d.cuh:11 / bench_conv1d_accel.cuh:11) */
    uint32_t algo;         /* b200conv_algo */
    uint32_t out_layout;   /* b200conv_layout */
    uint32_t flags;        /* B200CONV_FLAG_* */
} b200conv_config;

typedef struct b200conv_info {
    uint64_t macs_per_block;      /* T*B*L, the time-domain-equivalent work           */
    uint64_t flops_per_block;     /* algorithmic flops of the dominant kernel per block               */
    uint64_t alg_bytes_per_block; /* UPOLS FDL-MAC compulsory reads: T*16*P*(B+1); direct: 0          */
    uint64_t device_bytes;        /* engine-owned device memory                                       */
    uint64_t blocks_processed;    /* committed blocks since the last reset                            */
    uint64_t kernel_launches;     /* kernels launched by this engine since creation                   */
    uint32_t partitions;          /* UPOLS: P = ceil(L/B); direct: partial rows per track-tile (MS)   */
    uint32_t fft_size;            /* UPOLS: N = 2B; direct: 0                                         */
    uint32_t kernels_per_block;   /* launches per b200conv_process                                    */
    uint32_t sm_count;
    uint32_t stage_count;         /* entries valid in stage_ms / stage_name                           */
    uint32_t dominant_stage;      /* index of the roofline kernel (direct: FIR; UPOLS: FDL-MAC)       */
    float stage_ms[4];            /* accumulated CUDA-event time per stage while profiling is on      */
    uint32_t stage_calls;         /* blocks accumulated into stage_ms                                 */
    char stage_name[4][24];
} b200conv_info;

typedef struct b200conv_engine b200conv_engine;

uint32_t b200conv_abi_version(void);
const char* b200conv_last_error(void);

/* Replaces the device-side half of Conv1DBenchmark::setupBenchmark / Conv1DAccelBenchmark::
 * setupBenchmark: cudaMalloc of IR / FFT buffers, texture + cuFFT plan creation
 * (cuda/bench_conv1d.cu:115-157, cuda/bench_conv1d_accel.cu:88-150). */
int b200conv_create(const b200conv_config* cfg, b200conv_engine** out);

/* Replaces cleanupConvBuffers/cleanupTextureMemory and cleanupAccelBuffers/cleanupFFTPlans
 * (cuda/bench_conv1d.cu:210-227, cuda/bench_conv1d_accel.cu:339-375).  NULL is a no-op. */
void b200conv_destroy(b200conv_engine* e);

/* host_ir: float [T][L] track-major, host memory.  Replaces the IR upload + texture copy
 * (cuda/bench_conv1d.cu:123-157,180) and precomputeImpulseResponseFFTs
 * (cuda/bench_conv1d_accel.cu:175-228): the direct engine stores the taps in its tiled layout,
 * the UPOLS engine computes the P partition spectra.  Resets the stream state. */
int b200conv_load_ir(b200conv_engine* e, const float* host_ir);

/* host_hist: float [T][L-1], oldest sample first (hist[t][L-2] is the sample just before the next
 * block), or NULL for zeros.  No reference counterpart as an API: it reproduces what oracle R1's
 * flat input index does implicitly (cuda/bench_conv1d.cu:197-199).  Resets blocks_processed. */
int b200conv_prime_history(b200conv_engine* e, const float* host_hist);

/* Zero history: the next block equals oracle R2 (cuda/bench_conv1d_accel.cu:234-252). */
int b200conv_reset(b200conv_engine* e);

/* Per-track stereo bus gains, float [T][2] (L, R), host memory; NULL restores the default
 * constant-power pan from the global track index, scaled 1/sqrt(Tg). */
int b200conv_set_mix_gains(b200conv_engine* e, const float* host_gains);

/* One buffer of every track.  d_in float [T][B]; d_out float [T][B] or its [B][Tg] column tile
 * (pass the base pointer of the full [B][Tg] matrix); d_mix float [2][B] or NULL (written, not
 * accumulated).  All device pointers; work is enqueued on `stream` and not synchronised.
 * Replaces the kernel launch of performBenchmarkIteration: cuda/bench_conv1d.cu:92-101 and
 * cuda/bench_conv1d_accel.cu:263-299. */
int b200conv_process(b200conv_engine* e, const float* d_in, float* d_out, float* d_mix, uint32_t flags,
                     void* stream);

/* The same with HOST buffers: H2D of h_in, process, D2H of h_out / h_mix (either may be NULL),
 * stream-synchronised on return — one whole performBenchmarkIteration
 * (transferToDevice + launch + transferToHost, cuda/bench_conv1d.cu:82-105). Pinned host memory
 * makes the copies asynchronous up to the final synchronise. */
int b200conv_process_host(b200conv_engine* e, const float* h_in, float* h_out, float* h_mix,
                          uint32_t flags);

/* The same, split: b200conv_submit enqueues the block (input copy or in-place read, kernels, result copies) and
 * returns a ticket at once; b200conv_wait blocks until that block's results are in h_out / h_mix.  Two blocks may be
 * in flight (staging is double-buffered; a third submit first waits for the oldest): the device -> host copies of
 * block m run on a second stream under the kernels of block m+1, and the host prepares buffer m+1 meanwhile — the
 * overlap the reference's transferToDevice -> launch -> transferToHost sequence (cuda/bench_base.cu:30-42) and its
 * datacopy benchmarks (cuda/bench_datatransfer.cu:15-25) leave on the table.  Blocks complete in submission order;
 * the caller's buffers of a block must stay untouched until its wait returns.  b200conv_process_host == submit + wait. */
int b200conv_submit(b200conv_engine* e, const float* h_in, float* h_out, float* h_mix, uint32_t flags, uint64_t* ticket);
int b200conv_wait(b200conv_engine* e, uint64_t ticket);

/* Work / byte accounting for roofline figures, and per-stage CUDA-event times. */
int b200conv_query(b200conv_engine* e, b200conv_info* info);

/* on != 0: bracket every kernel of b200conv_process with CUDA events on the launch stream and
 * accumulate into info.stage_ms (synchronises at each process call: measurement mode only).
 * Replaces launchKernelTimed / CudaEventTimer (cuda/bench_utils.cuh:320-329). */
int b200conv_set_profiling(b200conv_engine* e, int on);

/* Batched real-to-complex FFT of `count` rows of n real samples (n a power of two, 32..8192) with the
 * engine's shared-memory Stockham transform: d_out is float2 [count][n/2+1], un-normalised, the
 * layout of cufftExecR2C.  Stateless; runs on the current device.  Replaces cufftPlan1d(R2C) +
 * cufftExecR2C of the reference's FFT1D benchmark (cuda/bench_fft.cu:63,105) — SURVEY.md §8(f) #3. */
int b200conv_rfft(const float* d_in, void* d_out, int count, int n, void* stream);

/* ---- channel strip on the output stage ---------------------------------------
 * Per-track post-processing of the convolved output, applied before the stereo bus, in this order:
 *   B200CONV_STRIP_STATS   mean and max of the strip INPUT per track, float [T][2]
 *                          — replaces GainStatsKernel's statistics, cuda/bench_gainstats.cu:15-31
 *   B200CONV_STRIP_GAIN    y = gain * x  — replaces GainKernel, cuda/bench_gain.cu:6-24 (and the gain of
 *                          GainStatsKernel, cuda/bench_gainstats.cu:22)
 *   B200CONV_STRIP_BIQUAD  Direct Form II biquad, w = x - a1 z1 - a2 z2, y = b0 w + b1 z1 + b2 z2, state
 *                          (z1, z2) per track carried from block to block — replaces IIRFilterKernel,
 *                          cuda/bench_iir.cu:10-44
 * Outputs, state and statistics are bit-identical to the reference's CPU loops (cuda/bench_gain.cu:90-92,
 * bench_gainstats.cu:121-142, bench_iir.cu:176-203) evaluated on the same input. */
#define B200CONV_STRIP_STATS 1u
#define B200CONV_STRIP_GAIN 2u
#define B200CONV_STRIP_BIQUAD 4u
#define B200CONV_STRIP_SHARED_COEFFS 8u /* `biquad` holds ONE coefficient set for all tracks (the reference's case) */

typedef struct b200conv_strip {
    uint32_t ops;        /* OR of the B200CONV_STRIP_* bits above                              */
    float gain;          /* used when gains == NULL                                            */
    const float* gains;  /* float [T] per-track gains, or NULL                                 */
    const float* biquad; /* float [T][5] = b0,b1,b2,a1,a2 (a0 == 1), or [5] with SHARED_COEFFS */
} b200conv_strip;

/* Attach (strip != NULL; pointers are HOST memory, copied) or remove (NULL) the engine's channel strip.
 * Attaching zeroes the biquad state.  With a strip attached, b200conv_process runs
 * convolution -> strip -> bus; B200CONV_PEEK leaves the biquad state untouched as well. */
int b200conv_set_strip(b200conv_engine* e, const b200conv_strip* strip);
/* set == 0: copy the biquad state float [T][2] = (z1, z2) to host_state; set != 0: load it. */
int b200conv_strip_state(b200conv_engine* e, float* host_state, int set);
/* mean / max of the latest block, float [T][2], copied to host_stats (synchronises the engine's stream). */
int b200conv_strip_stats(b200conv_engine* e, float* host_stats);

/* The strip alone, stateless, on caller-owned DEVICE memory of the current device (what the Gain,
 * GainStats and IIRFilter plugins call).  layout TRACK_MAJOR: d_in/d_out float [T][B]; SAMPLE_MAJOR:
 * float [B][ld] with track t in column col0 + t.  d_out may equal d_in.  strip->gains / ->biquad are
 * DEVICE pointers here.  d_state float [T][2] (required for BIQUAD; left unchanged under
 * B200CONV_PEEK), d_stats float [T][2] or NULL. */
int b200conv_strip_process(const float* d_in, float* d_out, uint32_t tracks, uint32_t block, uint32_t layout,
                           uint32_t ld, uint32_t col0, const b200conv_strip* strip, float* d_state, float* d_stats,
                           uint32_t flags, void* stream);

/* ---- the one collective of the path: all-reduce of the stereo bus over NVLink peer memory ----
 * No reference counterpart (the reference is single-GPU).  `peer_buffers[p]` is the
 * address, valid on THIS device, of rank p's symmetric buffer of b200conv_bus_buffer_bytes(world, n)
 * bytes (zero-initialised once; e.g. torch.distributed._symmetric_memory, or cudaIpc / cuMem
 * fabric handles).  `epoch` starts at 1 and increases by 1 per call on every rank.  d_local / d_out:
 * float[n] on this device (n = 2*B).  Every rank gets the bit-identical sum (fixed rank order).
 * d_error_flag (uint32 on this device) is set to 1 if a peer did not signal within the spin bound. */
size_t b200conv_bus_buffer_bytes(int world, int n);
int b200conv_bus_allreduce(const float* d_local, float* d_out, const uint64_t* peer_buffers, int rank, int world,
                           int n, uint32_t epoch, uint32_t* d_error_flag, void* stream);

/* Make engine `rank` of `world` engines (one per GPU, each owning a contiguous track range of the same
 * job: cfg.track_offset / total_tracks) a member of a bus group.  From then on the mix bus that
 * b200conv_process / b200conv_process_host deliver is the sum over ALL engines: the kernel that finishes
 * the last track of a block pushes the local bus into every peer's buffer over NVLink, waits for the
 * peers' flags and adds the `world` partials in rank order — inside the convolution launch, no further
 * kernel, bit-identical on every rank.  Every engine of the group must then process the same sequence of
 * blocks with a bus (d_mix / h_mix != NULL), PEEK blocks included.  peer_buffers[p]: address, valid on
 * THIS engine's device, of rank p's zero-initialised buffer of b200conv_bus_buffer_bytes(world, 2*B)
 * bytes (as for b200conv_bus_allreduce).  peer_buffers == NULL or world <= 1 detaches.
 * One engine per device is the intended use; engines of one group that share a device must be small enough to be
 * resident together (each waits inside its kernel for the others' bus), or run with B200CONV_BUS_SLICE=0.
 * No reference counterpart (the reference is single-GPU). */
int b200conv_attach_bus(b200conv_engine* e, const uint64_t* peer_buffers, int rank, int world);
/* Synchronises the engine's stream; B200CONV_ERR_CUDA if a peer missed a bus exchange since the last
 * call (bounded spin, a few seconds: the kernel gives up instead of hanging the GPU), else B200CONV_OK.
 * b200conv_process_host checks this itself on every block. */
int b200conv_bus_status(b200conv_engine* e);
/* Diagnostics: for the last `count` (<= 4096) blocks with a bus exchange, oldest first, host_stamps[2i] = device
 * %globaltimer (ns) when this rank's local bus was complete and its push began, host_stamps[2i+1] = when the summed bus
 * was complete; the difference is what the block spent on the exchange, rank skew included.  Tree-exchange kernels
 * only (direct engines); the engine must have been created with B200CONV_BUS_TRACE=1 in the environment. */
int b200conv_bus_trace(b200conv_engine* e, uint64_t* host_stamps, int count);
/* Diagnostics of the tensor-core direct engine (created with B200CONV_TC_TRACE=1 in the environment): device
 * %globaltimer (ns) stamps of the first work item of every CTA of the LAST launch, host_stamps[16 cta + s], s = 0 item
 * start, 1 band built, 2 own samples written (column group 0 only), 3 bus tree arrival done (group 0, bus only),
 * 4 MMAs complete, 5 ring epilogue done, 6 tap image landed / first MMA issued, 7 last MMA issued, 8 ring values
 * staged in TMEM, 9 / 10 first / fifth epilogue batch stored, 11..15 unused.  Returns the number of CTAs written
 * (<= max_ctas) or a negative error. */
int b200conv_tc_trace(b200conv_engine* e, uint64_t* host_stamps, int max_ctas);

/* ---- multi-GPU in one process: one engine per GPU over contiguous track ranges -------------------
 * cfg->tracks is the TOTAL track count Tg (cfg->device / track_offset / total_tracks are ignored);
 * GPU g of n owns tracks [g*Tg/n, (g+1)*Tg/n).  host_ir [Tg][L], host_hist [Tg][L-1], h_in [Tg][B],
 * h_out [Tg][B] or [B][Tg], h_mix [2][B] (the bus summed over ALL GPUs by b200conv_bus_allreduce over
 * peer-mapped buffers).  One persistent submission thread per GPU; the call returns when every GPU is
 * done.  No reference counterpart (the reference is single-GPU); SURVEY.md §8b/§8e. */
typedef struct b200conv_group b200conv_group;
int b200conv_group_create(const b200conv_config* cfg, int n_gpus, b200conv_group** out);
void b200conv_group_destroy(b200conv_group* g);
int b200conv_group_size(const b200conv_group* g);
int b200conv_group_load_ir(b200conv_group* g, const float* host_ir);
int b200conv_group_prime_history(b200conv_group* g, const float* host_hist);
int b200conv_group_reset(b200conv_group* g);
int b200conv_group_process_host(b200conv_group* g, const float* h_in, float* h_out, float* h_mix, uint32_t flags);
/* channel strip on every GPU's engine: strip->gains float [Tg], strip->biquad float [Tg][5] (or [5] shared);
 * host_state / host_stats float [Tg][2] (see b200conv_set_strip / _strip_state / _strip_stats) */
int b200conv_group_set_strip(b200conv_group* g, const b200conv_strip* strip);
int b200conv_group_strip_state(b200conv_group* g, float* host_state, int set);
int b200conv_group_strip_stats(b200conv_group* g, float* host_stats);
const char* b200conv_group_last_error(void);

/* Launch plan the engine would use for `cfg` on a device with `sm_count` SMs; needs no GPU.
 * plan[0..15] = direct: {A, CL, SPS, JSb, NS, G, Lc, cap, nbuf, xtile_blocks, ntiles, smem_bytes, MS, 0...}
 *               UPOLS : {P, M, logM, S, 0...};  DIRECT_TC: {A, C, NE, NGRP, R, capP, smem_bytes, grid, N, 0...}
 *               (A, R, capP, N describe one launch: the sub-block when block > 1024).
 * plan[15] = the b200conv_algo value of the kernel family the planner chose (ALGO_DIRECT may resolve to DIRECT_TC).
 * Used by the host-logic tests and by capacity planning. */
int b200conv_plan(const b200conv_config* cfg, int sm_count, int32_t plan[16]);

/* Measured FP32 FMA peak of `device` (dependent-free FFMA loop on every SM), TFLOP/s: the
 * roofline denominator for the direct FIR, which MEASURED_PEAKS.json does not carry. */
int b200conv_measure_fp32_peak(int device, double* tflops, double* elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* B200CONV_H_ */
