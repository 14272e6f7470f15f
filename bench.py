#!/usr/bin/env python
"""bench.py — the measurement contract of this repository.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4] [--impl ours|reference]

A "step" is one pass of the hot path over one buffer of every track: b200conv_process on inputs
already resident in HBM (`value`), and the same through the C-ABI with HOST buffers, host<->device
copies inside the timed region (`e2e`).  Workloads (BASELINE.json configs):
    c2 (default): direct FIR, 128 tracks x 512-sample buffers x 16384-tap IR per GPU  [configs[1]]
    c3          : UPOLS,     1024 tracks x 256-sample blocks  x 65536-tap IR per GPU  [configs[2]]
    c4          : UPOLS,      512 tracks x 512-sample buffers x 96000-tap IR per GPU  [configs[3] / 8]
Tracks shard across ranks (weak scaling: the per-GPU track count is fixed, IRs and mix gains use
the global track index); the only collective is the all-reduce of the stereo mix bus [2][B], which the
engine performs inside its last convolution kernel over NVLink peer memory (NCCL is the fallback).
The default line also carries the c3 and c4 results under "also" so one run shows both engines.
Every run — any N — ends with a PARITY leg outside the timed region: the engine is reset, a fresh
ceil(L/B)+2-block stream is pushed through the very call that was timed, two tracks per rank are
compared with the reference's own streaming loop (oracle/, checker only), the all-reduced bus with the
fp64 sum of the all-gathered per-rank partials, and the bus must be bit-identical on all ranks; a
failure makes the run exit non-zero.

Timing: W warm-up steps, then K timed steps, each bracketed by CUDA events on the launch stream,
with an L2 flush (256 MiB write, then a 256 MiB read of a second buffer so that the flush's dirty
lines are written back before the step instead of during it) between steps outside the brackets; the whole timed
region is bracketed by barrier + torch.cuda.synchronize(); per-rank totals are max-reduced.
`--impl reference` times the reference's own CPU implementation (oracle/_ref when built, else the
oracle port) on the host cores for the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 48000
WORKLOADS = {
    # name: (algo, tracks per GPU, B, L, layout, BASELINE config label)
    "c2": ("direct", 128, 512, 16384, "track_major", "bench_conv1d direct FIR: 128 tracks x 512-sample buffers x 16k-tap IR per GPU"),
    "c2ffma": ("direct_ffma", 128, 512, 16384, "track_major", "bench_conv1d direct FIR, FP32-FMA kernel forced (B200CONV_FLAG_FFMA_ONLY): 128 tracks x 512-sample buffers x 16k-tap IR per GPU"),
    "c3": ("upols", 1024, 256, 65536, "sample_major", "bench_conv1d_accel partitioned FFT convolution: 1024 tracks x 256-sample blocks x 64k-tap IR per GPU"),
    "c4": ("upols", 512, 512, 96000, "track_major", "4096 tracks x 96k-tap IR over 8 GPUs: 512 tracks per GPU, 512-sample buffers, stereo mix-bus reduce"),
}
NOMINAL_FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # 74.4
HBM_FALLBACK_GBS = 6650.0


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "MEASURED_PEAKS.json"
    except Exception:
        return {"hbm_gbs": HBM_FALLBACK_GBS}, "fallback (B200_PROFILING.md)"


class L2Flush:
    """Evict the step's working set from the 126 MB L2 between timed steps.  A plain write of a large
    buffer does that but leaves L2 full of DIRTY lines, and their write-back (~126 MB, ~20 us of HBM time)
    is then charged to the next step's reads — an artifact of the flush, not of the workload.  So the write
    is followed by a read pass over a second buffer: afterwards L2 holds clean foreign lines only."""
    DESCRIPTION = ("flushed between steps outside the event brackets: 256 MiB write, then 256 MiB read of a second "
                   "buffer (evicts the working set and leaves no dirty lines to write back inside the step)")

    def __init__(self, dev, mode=None):
        import torch
        self.mode = mode or os.environ.get("B200CONV_BENCH_FLUSH", "write+read")
        self.w = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
        self.r = torch.zeros(64 << 20, dtype=torch.float32, device=dev) if self.mode == "write+read" else None  # 256 MiB
        self.sink = None
        self.description = self.DESCRIPTION if self.r is not None else \
            "flushed between steps (256 MiB write outside the event brackets)"
        if self.mode == "none":
            self.description = "NOT flushed (diagnostic run, B200CONV_BENCH_FLUSH=none): not a valid bench line"

        # first use loads the reduction kernel (milliseconds, and not at the same moment on every rank): do it
        # here, not inside the first timed step, where the other ranks would wait for it in the bus all-reduce
        for k in range(2):
            self(k)
        torch.cuda.synchronize(dev)

    def __call__(self, k):
        if self.mode == "none":  # diagnostic only: warm L2 and instruction caches, not a valid bench setting
            return
        self.w.fill_(k & 0xFF)
        if self.r is not None:
            self.sink = self.r.sum()


class ClockSampler:
    """Polls NVML during the timed region: SM clock, max SM clock, active throttle reasons."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, index, period=0.004):
        self.samples, self.reasons, self.power = [], set(), []
        self.period, self._stop, self.max_mhz, self.ok = period, threading.Event(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as exc:  # NVML missing: report that rather than invent clocks
            self.err = str(exc)
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
                self.power.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.ok:
            self.thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self.ok:
            self.thread.join(timeout=2)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "no NVML samples"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples), "power_w_max": max(self.power) if self.power else None,
                "window": "warm-up + timed region (the timed region alone is ~1-10 ms of GPU work: too short to sample)"}


def pct(sorted_vals, q):
    return float(sorted_vals[min(len(sorted_vals) - 1, int(len(sorted_vals) * q))])  # nearest-rank, globals.cu:86-88


# =================================================================================================
# our arm
# =================================================================================================
def workload_config(name, world):
    """`config` of the JSON line — the SAME dict in both arms (ours and --impl reference)."""
    algo_name, T, B, L, layout_name, label = WORKLOADS[name]
    return {"workload": f"{name}: {label}", "algo": algo_name, "tracks_per_gpu": T, "total_tracks": T * world, "block": B,
            "ir_taps": L, "fs": FS, "out_layout": layout_name,
            "l2": "GPU arm: " + L2Flush.DESCRIPTION,
            "timing": "GPU arm: sum of per-step CUDA-event times on the launch stream, max over ranks; "
                      "reference arm: steady_clock around the CPU loop"}


def snr_db(got, ref):
    ref64 = np.asarray(ref, dtype=np.float64)
    err = float(np.sum((np.asarray(got, dtype=np.float64) - ref64) ** 2))
    return float(10 * np.log10(max(float(np.sum(ref64 ** 2)), 1e-300) / max(err, 1e-300)))


def parity_leg(name, eng, bus, step_fn, d_y, d_mix, rank, world, dev, dist, with_oracle=True):
    """Outside the timed region, in EVERY run: reset, stream ceil(L/B)+2 fresh blocks through the call that
    was timed, and compare (i) two tracks of this rank with the reference's streaming loop (SURVEY App. A.2;
    oracle/ used as the checker only), (ii) the all-reduced bus of the last block with the fp64 sum of the
    all-gathered per-rank partials, (iii) the bus of all ranks bit for bit."""
    import torch

    import gpuaudiobench_b200 as g
    from gpuaudiobench_b200 import synth
    from gpuaudiobench_b200.distributed import default_mix_gains
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle
    from concurrent.futures import ThreadPoolExecutor

    algo_name, T, B, L, layout_name, _ = WORKLOADS[name]
    Tg, t0 = T * world, T * rank
    M = (L + B - 1) // B + 2
    if not with_oracle:  # sweep points: bus + rank identity only (their oracle parity is tests/test_bench_shapes_gpu.py)
        M = min(M, 6)
    gen = torch.Generator(device=dev).manual_seed(9000 + rank)
    d_x = torch.rand(M, T, B, generator=gen, device=dev) * 2 - 1
    keep = sorted({0, T - 1})
    idx = torch.tensor(keep, device=dev)
    kept = torch.zeros(len(keep), M * B, device=dev)
    eng.reset()
    for m in range(M):
        step_fn(d_x[m].data_ptr())
        if layout_name == "sample_major":
            kept[:, m * B:(m + 1) * B] = d_y[:, t0 + idx].T
        else:
            kept[:, m * B:(m + 1) * B] = d_y[idx]
    torch.cuda.synchronize(dev)
    bus.check()
    y_last = (d_y[:, t0:t0 + T].T if layout_name == "sample_major" else d_y).double()  # [T][B]
    snrs, rel = [float("inf")], 0.0
    if with_oracle:
        got = kept.cpu().numpy()
        x_keep = d_x[:, idx, :].cpu().numpy()
        oracle = Oracle()
        h_rows = [synth.make_ir(Tg, L, t0 + t, t0 + t + 1)[0] for t in keep]
        with ThreadPoolExecutor(max_workers=len(keep)) as pool:  # the C loop releases the GIL
            want = list(pool.map(lambda i: oracle.stream(x_keep[:, i, :].ravel(), h_rows[i]), range(len(keep))))
        snrs = [snr_db(got[i], want[i]) for i in range(len(keep))]
        snrs += [snr_db(got[i][-B:], want[i][-B:]) for i in range(len(keep))]  # the last block alone
        rel = max(float(np.abs(got[i].astype(np.float64) - want[i]).max() / np.abs(want[i]).max()) for i in range(len(keep)))
    # bus: fp64 partial of this rank from its own last-block outputs, summed over ranks
    gains = default_mix_gains(Tg, t0, t0 + T).to(dev).double()  # [T][2]
    # (elementwise + reduce on purpose: an fp64 `@` would put a library GEMM into the ncu launch lists of this command)
    partial = (gains.T.unsqueeze(-1) * y_last.unsqueeze(0)).sum(1)  # [2][B] fp64
    bus_got = d_mix.clone()
    identical = True
    if world > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM)
        allbus = [torch.empty_like(bus_got) for _ in range(world)]
        dist.all_gather(allbus, bus_got)
        identical = all(bool(torch.equal(allbus[0], b)) for b in allbus)
    bus_snr = snr_db(bus_got.cpu().numpy(), partial.cpu().numpy())
    stats = torch.tensor([min(snrs), -rel, bus_snr], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MIN)
    snr_min, rel_max, bus_snr = float(stats[0]), -float(stats[1]), float(stats[2])
    min_snr, max_rel = (100.0, 1e-5) if algo_name.startswith("direct") else (90.0, 1e-4)
    ok = bool(snr_min >= min_snr and rel_max <= max_rel and bus_snr >= 100.0 and identical)
    del d_x, kept
    return {"ok": ok, "snr_db_min": snr_min if with_oracle else None, "max_abs_err_rel": rel_max if with_oracle else None, "bus_snr_db": bus_snr, "ranks_bit_identical": identical,
            "blocks": M, "tracks_checked_per_rank": keep, "ranks": world,
            "tolerance": f"SNR >= {min_snr:.0f} dB and max|err| <= {max_rel:g} max|y_ref| vs the reference's streaming loop "
                         "(fp32, bench_conv1d_accel.cu:234-252 on the whole stream); bus >= 100 dB vs the fp64 sum of the "
                         "all-gathered per-rank partials; bus bit-identical on all ranks"}


def run_workload(name, args, rank, world, local_rank, dist, want_cpu_baseline, sweep=False):
    import torch

    import gpuaudiobench_b200 as g
    from gpuaudiobench_b200 import synth
    from gpuaudiobench_b200.distributed import EngineBusGroup

    algo_name, T, B, L, layout_name, label = WORKLOADS[name]
    algo = {"direct": g.ALGO_DIRECT, "direct_ffma": g.ALGO_DIRECT, "upols": g.ALGO_UPOLS}[algo_name]
    eng_flags = g.engine.FLAG_FFMA_ONLY if algo_name == "direct_ffma" else 0
    layout = g.OUT_SAMPLE_MAJOR if layout_name == "sample_major" else g.OUT_TRACK_MAJOR
    Tg, t0 = T * world, T * rank
    K, W = args.steps, args.warmup
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.current_stream(dev)

    # --- synthetic job: reference-shaped IRs (global track index) and input stream -------------
    ir = synth.make_ir(Tg, L, t0, t0 + T)
    NB = 8  # distinct input buffers cycled through, resident in HBM
    x_host = synth.make_input(NB * T * B, seed=42 + rank).reshape(NB, T, B)
    eng = g.ConvEngine(T, B, L, algo, layout, device=local_rank, track_offset=t0, total_tracks=Tg, flags=eng_flags)
    eng.load_ir(ir)
    del ir
    d_x = torch.from_numpy(x_host).to(dev)
    out_shape = (B, Tg) if layout == g.OUT_SAMPLE_MAJOR else (T, B)
    d_y = torch.zeros(out_shape, device=dev)
    d_mix = torch.zeros(2, B, device=dev)
    flush = L2Flush(dev)
    # the one collective: with symmetric memory the engine's own kernel does it (no further launch)
    bus = EngineBusGroup(eng, d_mix, force_nccl=bool(os.environ.get("B200CONV_BUS_NCCL")))

    def step_ptr(ptr):
        eng.process(ptr, d_y.data_ptr(), d_mix.data_ptr(), stream=stream.cuda_stream)
        bus.reduce()  # no-op when the exchange ran inside the kernel

    def step(k):
        step_ptr(d_x[k % NB].data_ptr())

    clocks = ClockSampler(local_rank)  # nvmlInit takes milliseconds: do it before the ranks line up
    clocks.__enter__()                 # sampled across warm-up + timed region (the timed region alone is ~ms)
    # warm-up: at least W steps, and enough to fill the history / delay line with signal
    fill = max(W, min((L + B - 1) // B + 2, 400))
    for k in range(fill):
        step(k)
    torch.cuda.synchronize(dev)

    # --- timed region: K steps, per-step events, L2 flushed between steps ----------------------
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    launches_before = eng.query()["kernel_launches"]

    def aligned_start():
        """barrier + synchronize, then every rank leaves at the same instant of the node-wide
        monotonic clock: without it the first timed collective measures how late the slowest
        rank's CPU woke up from the barrier (ms), not the step."""
        if world > 1:
            dist.barrier()
            t_go = torch.tensor([time.clock_gettime(time.CLOCK_MONOTONIC) + 0.003], dtype=torch.float64, device=dev)
            dist.all_reduce(t_go, op=dist.ReduceOp.MAX)
            t_go = float(t_go.item())
            torch.cuda.synchronize(dev)
            while time.clock_gettime(time.CLOCK_MONOTONIC) < t_go:
                pass
        else:
            torch.cuda.synchronize(dev)

    import gc
    gc.collect()
    gc.disable()  # a collector pause on one rank's host thread shows up as a late bus push on every rank's device clock
    aligned_start()
    wall0 = time.perf_counter()
    for k in range(K):
        flush(k)
        ev0[k].record(stream)
        step(k)
        ev1[k].record(stream)
    torch.cuda.synchronize(dev)
    gc.enable()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - wall0
    clocks.__exit__()
    lat = np.array([a.elapsed_time(b) for a, b in zip(ev0, ev1)], dtype=np.float64)  # ms
    launches = eng.query()["kernel_launches"] - launches_before
    total_ms = float(lat.sum())
    if world > 1:
        tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        total_ms = float(tmax.item())
        lat_t = torch.from_numpy(lat).to(dev)
        dist.all_reduce(lat_t, op=dist.ReduceOp.MAX)  # a buffer is done when the slowest rank is done
        lat = lat_t.cpu().numpy()
    ms_per_step = total_ms / K
    macs_per_step = float(Tg) * B * L
    value = macs_per_step / (ms_per_step * 1e-3) / 1e9
    s = np.sort(lat)
    deadline_ms = 1000.0 * B / FS

    # --- sustained: the same step back to back for ~1 s (no flush), clocks sampled — what the step does at the
    # clocks a long run settles at (the timed region above is a few ms of GPU work at boost clocks) -------------
    n_sus = int(max(50, min(20000, (50.0 if sweep else 1000.0) / ms_per_step)))
    sus_clocks = ClockSampler(local_rank, period=0.02)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    aligned_start()
    with sus_clocks:
        e0.record(stream)
        for k in range(n_sus):
            step(k)
        e1.record(stream)
        torch.cuda.synchronize(dev)
    sus_ms = e0.elapsed_time(e1) / n_sus
    if world > 1:
        tmax = torch.tensor([sus_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        sus_ms = float(tmax.item())
    sc = sus_clocks.summary()
    sustained = {"ms_per_step": sus_ms, "value": macs_per_step / (sus_ms * 1e-3) / 1e9, "unit": "GMAC/s", "steps": n_sus,
                 "sm_mhz": sc.get("sm_mhz"), "power_w_max": sc.get("power_w_max"), "reasons": sc.get("reasons"),
                 "note": "back-to-back steps, L2 NOT flushed (inputs cycle through 8 buffers; IR tables / delay lines are "
                         "the working set): context for the flushed, boost-clock number above, not the bench value"}

    # --- roofline of the dominant kernel: per-kernel CUDA events on the launch stream ----------
    if world > 1:
        dist.barrier()
    eng.set_profiling(True)
    KP = min(K, 100)
    for k in range(KP):
        flush(k)
        step(k)
    torch.cuda.synchronize(dev)
    q = eng.query()
    eng.set_profiling(False)
    dom = q["dominant_stage"]
    stage_ms = [m / max(1, q["stage_calls"]) for m in q["stage_ms"][:q["stage_count"]]]
    dom_ms = stage_ms[dom]
    peaks, peak_src = measured_peaks()
    if q["stage_name"][dom] == "tc_toeplitz":  # ALGO_DIRECT as the planner dispatched it: the tensor-core kernel
        # tensor pipe: the kernel ISSUES 3 TF32 products (hi/lo split) per algorithmic MAC, on L padded to 128 taps
        # and the C - 1 tensor-core columns padded to N-column groups; the roofline is the dense TF32 rate = half the measured bf16 rate
        tf32_peak = float(peaks.get("bf16_tflops", 2250.0 * 0.72)) / 2.0
        plan = g.plan(T, B, L, algo, flags=eng_flags)
        issued = 3.0 * 2.0 * T * (plan["A"] * 128) * 128 * plan["N"] * plan["NGRP"] * 1.0  # per block: A row blocks x 128 K x (128 x N) x 3
        achieved = issued / (dom_ms * 1e-3) / 1e12
        fp32_peak, _ = g.measure_fp32_peak(local_rank)
        alg = q["flops_per_block"] / (dom_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": q["stage_name"][dom], "achieved": achieved, "peak": tf32_peak,
                    "unit": "TFLOP/s", "frac": achieved / tf32_peak, "traffic": None,
                    "peak_source": f"dense TF32 = bf16_tflops of {peak_src} / 2 (kind::tf32 has K = 8 per dispatch against 16 for bf16)",
                    "issued_flops_per_launch": issued, "algorithmic_flops_per_launch": q["flops_per_block"],
                    "algorithmic_tflops": alg, "algorithmic_frac_of_fp32_fma_peak": alg / fp32_peak,
                    "step_frac_of_nominal": q["flops_per_block"] / (ms_per_step * 1e-3) / 1e12 / NOMINAL_FP32_TFLOPS,
                    "note": "algorithmic_frac_of_fp32_fma_peak > 1 means the tensor-core variant runs the fp32-accurate "
                            "direct form faster than the FP32 FMA pipe could at 100 %"}
    elif algo == g.ALGO_DIRECT:
        fp32_peak, _ = g.measure_fp32_peak(local_rank)
        achieved = q["flops_per_block"] / (dom_ms * 1e-3) / 1e12
        roofline = {"bound": "fp32_fma", "kernel": q["stage_name"][dom], "achieved": achieved, "peak": fp32_peak,
                    "unit": "TFLOP/s", "frac": achieved / fp32_peak, "traffic": None,
                    "peak_source": "FFMA microbenchmark measured in this run (b200conv_measure_fp32_peak); "
                                   "MEASURED_PEAKS.json has no CUDA-core figure; nominal 74.4 TFLOP/s at 1965 MHz",
                    "frac_of_nominal": achieved / NOMINAL_FP32_TFLOPS,
                    "step_frac_of_nominal": q["flops_per_block"] / (ms_per_step * 1e-3) / 1e12 / NOMINAL_FP32_TFLOPS,
                    "algorithmic_flops_per_launch": q["flops_per_block"]}
    else:
        achieved = q["alg_bytes_per_block"] / (dom_ms * 1e-3) / 1e9
        peak = float(peaks["hbm_gbs"])
        roofline = {"bound": "hbm", "kernel": q["stage_name"][dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": None,
                    "peak_source": f"hbm_gbs of {peak_src} (measured COPY bandwidth, reads + writes; this kernel only "
                                   "reads, and a read-only stream can exceed it: ncu measured 6.84 TB/s for the FDL-MAC)",
                    "frac_of_spec_8tbs": achieved / 8000.0,
                    "step_frac": q["alg_bytes_per_block"] / (ms_per_step * 1e-3) / 1e9 / peak,
                    "algorithmic_bytes_per_launch": q["alg_bytes_per_block"]}
    for fn in ("r02_traffic.json", "r01_traffic.json"):  # per-launch DRAM traffic of the dominant kernel (committed ncu capture)
        try:
            with open(os.path.join(ROOT, "profiles", fn)) as f:
                tr = json.load(f).get(name)
            if tr and tr["kernel"] == q["stage_name"][dom]:
                roofline["traffic"] = tr["bytes"]
                roofline["traffic_source"] = f"profiles/{fn} (ncu --set full, one launch)"
                break
        except Exception:
            pass
    roofline["kernel_ms"] = dom_ms
    roofline["stage_ms"] = dict(zip(q["stage_name"][:q["stage_count"]], stage_ms))
    roofline["kernel_share_of_step"] = dom_ms / sum(stage_ms)

    # --- e2e: the same steps through the C ABI with HOST buffers (pinned), copies timed ---------
    # b200conv_process_host at every N: with the bus exchange inside the kernel the multi-GPU host call IS the
    # single-GPU host call (pinned buffers are read / written in place over PCIe by the kernels)
    h_in = torch.from_numpy(x_host).pin_memory()
    h_out = torch.zeros(out_shape).pin_memory()
    h_mix = torch.zeros(2, B).pin_memory()
    d_in2 = torch.zeros(T, B, device=dev)

    def e2e_step(k):
        if world == 1 or bus.in_kernel:
            eng.process_host_ptr(h_in[k % NB].data_ptr(), h_out.data_ptr(), h_mix.data_ptr())
        else:  # fallback only (no symmetric memory): staged copies around the NCCL all-reduce
            d_in2.copy_(h_in[k % NB], non_blocking=True)
            step_ptr(d_in2.data_ptr())
            if layout == g.OUT_SAMPLE_MAJOR:  # only this rank's column tile of [B][Tg]
                h_out[:, t0:t0 + T].copy_(d_y[:, t0:t0 + T], non_blocking=True)
            else:
                h_out.copy_(d_y, non_blocking=True)
            h_mix.copy_(d_mix, non_blocking=True)
            stream.synchronize()

    for k in range(max(3, min(W, 10))):
        e2e_step(k)
    aligned_start()
    e2e_lat = np.empty(K)
    for k in range(K):
        flush(k)
        torch.cuda.synchronize(dev)
        t_a = time.perf_counter()
        e2e_step(k)
        e2e_lat[k] = (time.perf_counter() - t_a) * 1e3
    e2e_total = float(e2e_lat.sum())
    if world > 1:
        tmax = torch.tensor([e2e_total], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_total = float(tmax.item())
    e2e_ms = e2e_total / K
    es = np.sort(e2e_lat)
    # the same host-buffer blocks through b200conv_submit / b200conv_wait, two in flight (f1: block m's device->host
    # copies and the host's hand-over run under block m+1's kernels); no L2 flush here — it would have to run on the
    # engine's own stream — so this is reported beside the flushed serial figure, not instead of it
    pipe_ms = None
    if world == 1 or bus.in_kernel:
        h_out2 = [h_out, torch.zeros(out_shape).pin_memory()]
        h_mix2 = [h_mix, torch.zeros(2, B).pin_memory()]
        for rep in range(2):  # first pass warms the second staging slot
            aligned_start()
            t_a = time.perf_counter()
            prev = None
            for k in range(K):
                tk = eng.submit_ptr(h_in[k % NB].data_ptr(), h_out2[k & 1].data_ptr(), h_mix2[k & 1].data_ptr())
                if prev is not None:
                    eng.wait(prev)
                prev = tk
            eng.wait(prev)
            pipe_ms = (time.perf_counter() - t_a) * 1e3 / K
        if world > 1:
            tmax = torch.tensor([pipe_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            pipe_ms = float(tmax.item())
    out_bytes = int(np.prod(out_shape)) * 4 if layout == g.OUT_TRACK_MAJOR else T * B * 4
    tail = "p99" if K >= 100 else "max"  # with fewer than 100 steps the nearest-rank p99 IS the maximum
    e2e = {"value": macs_per_step / (e2e_ms * 1e-3) / 1e9, "unit": "GMAC/s", "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": world * T * B * 4, "d2h_bytes_per_step": world * (out_bytes + 2 * B * 4),
           "p50_ms": pct(es, 0.50), "p99_ms": pct(es, 0.99), "tail_is": tail, "meets_deadline": bool(pct(es, 0.99) <= deadline_ms),
           "api": "b200conv_process_host (C ABI, pinned host buffers; bus exchange inside the kernel)" if (world == 1 or bus.in_kernel)
                  else "pinned H2D + b200conv_process + NCCL mix-bus all-reduce + D2H (fallback)",
           "pipelined_ms_per_step": pipe_ms,
           "pipelined_note": "b200conv_submit / b200conv_wait, two blocks in flight, wall clock over all steps, L2 not flushed"}

    # --- parity, outside every timed region, in every run ---------------------------------------
    parity = parity_leg(name, eng, bus, step_ptr, d_y, d_mix, rank, world, dev, dist, with_oracle=not sweep)

    cfg = workload_config(name, world)
    result = {
        "metric": "conv_tracks_x_ir_taps_gmac_per_s", "value": value, "unit": "GMAC/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (mt19937 seed 42 uniform(-1,1) input; Hamming-windowed-sinc IRs scaled 1/L, as the reference generates)",
        "config": cfg,
        "run": {"warmup_steps_run": fill,
                "warmup_note": "max(requested warm-up, ceil(L/B)+2 blocks) so that the whole history / delay line carries signal",
                "partitions_or_splits": q["partitions"],
                "collective": f"all-reduce of the stereo mix bus float[2][B]: {bus.kind}"},
        "rt_tracks": value * 1e9 / (L * FS),
        "latency_ms": {"p50": pct(s, 0.50), "p95": pct(s, 0.95), "p99": pct(s, 0.99), "max": float(s[-1]), "tail_is": tail,
                       "slowest_step": int(np.argmax(lat)), "deadline": deadline_ms,
                       "meets_deadline": bool(pct(s, 0.99) <= deadline_ms)},
        "wall_ms_per_step_incl_flush": wall * 1e3 / K,
        "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks.summary(),
        "sustained": sustained, "parity": parity,
        "engine_device_bytes": q["device_bytes"],
    }
    if want_cpu_baseline:
        result["cpu_baseline"] = cpu_baseline(name, budget_s=1.5, reps=3) if sweep else cpu_baseline(name, budget_s=12.0)
    bus.close()
    eng.close()
    del d_x, d_y, flush
    torch.cuda.empty_cache()
    return result


# =================================================================================================
# CPU legs: the only places bench.py touches oracle/
# =================================================================================================
def _cpu_lib():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle, RefLib
    if RefLib.available():
        try:
            return RefLib(), "reference"
        except OSError:
            pass
    return Oracle(), "port"


def cpu_time_block(lib, algo_name, x, h, L, B, T, threads):
    fn = lib.time_r1 if algo_name.startswith("direct") else lib.time_r2
    return fn(x, h, L, B, T, threads)


def cpu_baseline(name, budget_s, threads=None, reps=5):
    """The reference's CPU path (R1 for Conv1D, R2 for Conv1D_accel) on the host cores, bounded
    sample: full B and L, a track subset sized to the time budget, scaled linearly in T.  One untimed
    warm-up pass (page faults, thread start-up, clocks), then `reps` timed passes: the MEDIAN is the
    value, min/max are reported as the spread."""
    from gpuaudiobench_b200 import synth
    algo_name, T, B, L, _, _ = WORKLOADS[name]
    lib, kind = _cpu_lib()
    cores = threads or lib.hardware_threads()
    # R2 skips the iterations its bounds test rejects, so its loop count is what matters for time
    iters_per_track = float(B) * L
    rate_guess = 0.5e9 * cores
    Ts = int(max(cores, min(T, budget_s / (reps + 1) * rate_guess / iters_per_track)))
    Ts = max(1, min(T, Ts))
    x = synth.make_input(Ts * B)
    h = synth.make_ir(T, L, 0, Ts)
    cpu_time_block(lib, algo_name, x, h, L, B, Ts, cores)  # warm-up
    runs = sorted(cpu_time_block(lib, algo_name, x, h, L, B, Ts, cores) for _ in range(reps))
    secs = runs[len(runs) // 2]
    gmacs = Ts * iters_per_track / secs / 1e9
    # the reference itself runs this loop on ONE thread (bench_conv1d.cu:42-55 calls it from setupBenchmark):
    # time that too, on a subset sized for about half a second (SURVEY §8d asks for both figures)
    T1 = int(max(1, min(Ts, 0.5 * 0.5e9 / iters_per_track)))
    cpu_time_block(lib, algo_name, x[:T1 * B], h[:T1], L, B, T1, 1)
    secs1 = sorted(cpu_time_block(lib, algo_name, x[:T1 * B], h[:T1], L, B, T1, 1) for _ in range(3))[1]
    gmacs1 = T1 * iters_per_track / secs1 / 1e9
    return {"value": gmacs, "unit": "GMAC/s", "cores": cores, "kind": kind,
            "spread": {"reps": reps, "min": Ts * iters_per_track / runs[-1] / 1e9, "max": Ts * iters_per_track / runs[0] / 1e9},
            "value_1_thread": gmacs1, "sample_1_thread": f"{T1} tracks on one thread, median of 3, {secs1:.2f} s",
            "sample": f"{'R1 bench_conv1d.cu:188-208' if algo_name.startswith('direct') else 'R2 bench_conv1d_accel.cu:234-252'} "
                      f"on {Ts} of {T} tracks at full B={B}, L={L}, {cores} threads over contiguous track ranges; one warm-up "
                      f"pass, median of {reps} ({secs:.2f} s each); GMAC/s counts T*B*L loop iterations (time-domain equivalent)",
            "seconds": secs, "ms_per_block_scaled_to_T": secs * 1e3 * T / Ts,
            "rt_tracks": gmacs * 1e9 / (L * FS)}


def run_reference(args, rank, world):
    if rank != 0:
        return None
    from gpuaudiobench_b200 import synth
    name = args.workload
    algo_name, T, B, L, _, label = WORKLOADS[name]
    lib, kind = _cpu_lib()
    cores = lib.hardware_threads()
    K, W = args.steps, args.warmup
    # bounded sample: whole run (K + W steps) within ~150 s
    iters_per_track = float(B) * L
    per_step_budget = 150.0 / (K + W)
    Ts = int(max(1, min(T, per_step_budget * 0.5e9 * cores / iters_per_track)))
    x = synth.make_input(Ts * B)
    h = synth.make_ir(T * world, L, 0, Ts)
    for _ in range(W):
        cpu_time_block(lib, algo_name, x, h, L, B, Ts, cores)
    secs = [cpu_time_block(lib, algo_name, x, h, L, B, Ts, cores) for _ in range(K)]
    mean_s = float(np.mean(secs))
    value = Ts * iters_per_track / mean_s / 1e9
    ms_full = mean_s * 1e3 * (T * world) / Ts
    sample = (f"{kind}: {'R1' if algo_name.startswith('direct') else 'R2'} CPU loop, {Ts} of {T * world} tracks per step at full "
              f"B={B}, L={L}, {cores} host threads; ms_per_step is scaled linearly to all tracks")
    return {"impl": "reference", "metric": "conv_tracks_x_ir_taps_gmac_per_s", "value": value, "unit": "GMAC/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_full, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic (same generators as our arm)",
            "config": workload_config(name, world),
            "cpu_baseline": {"value": value, "unit": "GMAC/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "GMAC/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "rt_tracks": value * 1e9 / (L * FS), "gpu_launches": 0}


def run_sweep(args, rank, world, local_rank, dist):
    """BASELINE config 5: buffer-size sweep 32..4096 at C4's per-GPU tracks and IR length (UPOLS) and
    at C2's (direct): throughput vs p50/p99 per-buffer latency vs the B/fs deadline
    (meets_deadline = p99 <= 1000*B/fs, cuda/globals.cu:86-89,155-156), with the reference's CPU loop on
    the host cores beside every point (rank 0, bounded sample)."""
    rows = []
    for algo_name, T, L, sizes in (("upols", 512, 96000, (32, 64, 128, 256, 512, 1024, 2048, 4096)),
                                   ("direct", 128, 16384, (32, 64, 128, 256, 512, 1024, 2048, 4096))):
        for B in sizes:
            key = f"sweep_{algo_name}_{B}"
            if args.sweep_filter and not any(tok == f"{algo_name}:{B}" or tok == algo_name for tok in args.sweep_filter.split(",")):
                continue
            WORKLOADS[key] = (algo_name, T, B, L, "track_major", f"sweep: {algo_name} {T} tracks/GPU x {B}-sample buffers x {L}-tap IR")
            r = run_workload(key, args, rank, world, local_rank, dist, want_cpu_baseline=(rank == 0), sweep=True)
            cpu = r.get("cpu_baseline") or {}
            rows.append({"algo": algo_name, "block": B, "tracks_per_gpu": T, "ir_taps": L, "n_gpus": world, "steps": args.steps,
                         "ms_per_step": r["ms_per_step"], "gmac_per_s": r["value"], "rt_tracks": r["rt_tracks"],
                         "p50_ms": r["latency_ms"]["p50"], "p99_ms": r["latency_ms"]["p99"], "max_ms": r["latency_ms"]["max"],
                         "deadline_ms": r["latency_ms"]["deadline"], "meets_deadline": r["latency_ms"]["meets_deadline"],
                         "e2e_ms": r["e2e"]["ms_per_step"], "e2e_p99_ms": r["e2e"]["p99_ms"],
                         "e2e_meets_deadline": r["e2e"]["meets_deadline"],
                         "roofline_frac": r["roofline"]["frac"], "roofline_bound": r["roofline"]["bound"],
                         "step_frac": r["roofline"].get("step_frac", r["roofline"].get("step_frac_of_nominal")),
                         "gpu_launches_per_step": r["gpu_launches"] / args.steps,
                         "bus_parity_ok": r["parity"]["ok"], "ranks_bit_identical": r["parity"]["ranks_bit_identical"],
                         "cpu_gmac_per_s": cpu.get("value"), "cpu_cores": cpu.get("cores"), "cpu_kind": cpu.get("kind"),
                         "cpu_rt_tracks": cpu.get("rt_tracks")})
    return rows


def run_latency(args, rank, world, local_rank, dist):
    """SURVEY §8(d) latency run: `--latency N` blocks of the chosen workload through the host-buffer
    path (pinned H2D, kernels, bus all-reduce, D2H, synchronise), after >= P warm-up blocks, submitted
    (i) back-to-back and (ii) periodically at the buffer period B/fs (spin to the deadline, like the
    reference's Metal DAWSimulator).  Latency = submission -> results on the host, max over ranks."""
    import torch

    import gpuaudiobench_b200 as g
    from gpuaudiobench_b200 import synth
    from gpuaudiobench_b200.distributed import EngineBusGroup

    algo_name, T, B, L, layout_name, label = WORKLOADS[args.workload]
    algo = {"direct": g.ALGO_DIRECT, "direct_ffma": g.ALGO_DIRECT, "upols": g.ALGO_UPOLS}[algo_name]
    Tg, t0 = T * world, T * rank
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.current_stream(dev)
    eng = g.ConvEngine(T, B, L, algo, g.OUT_TRACK_MAJOR, device=local_rank, track_offset=t0, total_tracks=Tg)
    eng.load_ir(synth.make_ir(Tg, L, t0, t0 + T))
    NB = 8
    h_in = torch.from_numpy(synth.make_input(NB * T * B, seed=42 + rank).reshape(NB, T, B)).pin_memory()
    h_out = torch.zeros(T, B).pin_memory()
    h_mix = torch.zeros(2, B).pin_memory()
    d_in, d_y, d_mix = torch.zeros(T, B, device=dev), torch.zeros(T, B, device=dev), torch.zeros(2, B, device=dev)
    bus = EngineBusGroup(eng, d_mix, force_nccl=bool(os.environ.get("B200CONV_BUS_NCCL")))

    def block(k):
        if world == 1 or bus.in_kernel:  # the multi-GPU host call is the single-GPU host call
            eng.process_host_ptr(h_in[k % NB].data_ptr(), h_out.data_ptr(), h_mix.data_ptr())
        else:
            d_in.copy_(h_in[k % NB], non_blocking=True)
            eng.process(d_in.data_ptr(), d_y.data_ptr(), d_mix.data_ptr(), stream=stream.cuda_stream)
            bus.reduce()
            h_out.copy_(d_y, non_blocking=True)
            h_mix.copy_(d_mix, non_blocking=True)
            stream.synchronize()

    N = args.latency
    period = B / FS
    out = {}
    for mode in ("back_to_back", "periodic"):
        for k in range(min((L + B - 1) // B + 2, 400)):
            block(k)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        lat = np.empty(N)
        t_next = time.perf_counter() + period
        for k in range(N):
            if mode == "periodic":
                while time.perf_counter() < t_next:
                    pass
                t_next += period
            a = time.perf_counter()
            block(k)
            lat[k] = (time.perf_counter() - a) * 1e3
        if world > 1:
            lt = torch.from_numpy(lat).to(dev)
            dist.all_reduce(lt, op=dist.ReduceOp.MAX)
            lat = lt.cpu().numpy()
        s = np.sort(lat)
        out[mode] = {"blocks": N, "p50_ms": pct(s, 0.50), "p95_ms": pct(s, 0.95), "p99_ms": pct(s, 0.99), "max_ms": float(s[-1]),
                     "mean_ms": float(lat.mean()), "deadline_ms": period * 1e3, "meets_deadline": bool(pct(s, 0.99) <= period * 1e3),
                     "missed_blocks": int((lat > period * 1e3).sum())}
    bus.check()
    bus.close()
    return {"latency_run": out, "workload": f"{args.workload}: {label}", "n_gpus": world, "total_tracks": Tg, "block": B,
            "ir_taps": L, "fs": FS, "path": "host buffers -> results on host (e2e)", "collective": bus.kind}


def run_strip(args, local_rank):
    """Channel-strip measurements, one GPU: (1) the three plugin configurations alone on
    the reference's default shape 128 x 512 and on 1024 x 512, device time per call by CUDA events with the
    L2 flushed, beside the reference's CPU loop on the host cores (one thread, as the reference runs it);
    (2) what attaching the full strip costs per block on C2 and C3."""
    import torch
    import gpuaudiobench_b200 as g
    from gpuaudiobench_b200 import synth
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle
    oracle = Oracle()
    dev = torch.device("cuda", local_rank)
    flush = L2Flush(dev)
    K = min(args.steps, 100)
    rows = []
    coef = synth.butterworth_lowpass(0.25)
    d_coef = torch.from_numpy(coef).to(dev)
    for T, B in ((128, 512), (1024, 512)):
        x = synth.make_input(T * B, seed=42).reshape(T, B)
        d_x = torch.from_numpy(x).to(dev)
        d_y = torch.empty_like(d_x)
        d_state = torch.zeros(T, 2, device=dev)
        d_stats = torch.zeros(T, 2, device=dev)
        for name, ops, gain in (("gain", g.STRIP_GAIN, 2.0), ("GainStats", g.STRIP_GAIN | g.STRIP_STATS, 0.5),
                                ("IIRFilter", g.STRIP_BIQUAD, 1.0)):
            def call():
                g.strip_process(d_x.data_ptr(), d_y.data_ptr(), T, B, ops, gain=gain, d_biquad=d_coef.data_ptr(),
                                d_state=d_state.data_ptr(), d_stats=d_stats.data_ptr())
            for _ in range(5):
                call()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
            for k in range(K):
                flush(k)
                ev[k][0].record()
                call()
                ev[k][1].record()
            torch.cuda.synchronize(dev)
            ms = np.sort(np.array([a.elapsed_time(b) for a, b in ev]))
            t0 = time.perf_counter()
            reps = 0
            st = np.zeros((T, 2), np.float32)
            while time.perf_counter() - t0 < 0.3:
                if name == "gain":
                    oracle.gain(x, 2.0)
                elif name == "GainStats":
                    oracle.gainstats(x, 0.5)
                else:
                    oracle.iir(x, coef, st)
                reps += 1
            cpu_ms = (time.perf_counter() - t0) * 1e3 / reps
            rows.append({"plugin": name, "tracks": T, "block": B, "gpu_ms_p50": pct(ms, 0.5), "gpu_ms_p99": pct(ms, 0.99),
                         "bytes_moved": 2 * T * B * 4, "gb_per_s": 2 * T * B * 4 / (pct(ms, 0.5) * 1e-3) / 1e9,
                         "cpu_ms_1_thread": cpu_ms, "chain_bound_us": (B * 12 / 1.9e3) if name == "IIRFilter" else None})
    fused = {}
    for wl in ("c2", "c3"):
        algo_name, T, B, L, layout_name, label = WORKLOADS[wl]
        algo = g.ALGO_DIRECT if algo_name == "direct" else g.ALGO_UPOLS
        layout = g.OUT_SAMPLE_MAJOR if layout_name == "sample_major" else g.OUT_TRACK_MAJOR
        eng = g.ConvEngine(T, B, L, algo, layout, device=local_rank)
        eng.load_ir(synth.make_ir(T, L, 0, T))
        d_x = torch.from_numpy(synth.make_input(T * B, seed=1).reshape(T, B)).to(dev)
        d_y = torch.zeros((B, T) if layout == g.OUT_SAMPLE_MAJOR else (T, B), device=dev)
        d_mix = torch.zeros(2, B, device=dev)
        res = {}
        for tag in ("plain", "strip"):
            if tag == "strip":
                eng.set_strip(g.STRIP_STATS | g.STRIP_GAIN | g.STRIP_BIQUAD, gain=0.5, biquad=coef)
            for k in range(10):
                eng.process(d_x.data_ptr(), d_y.data_ptr(), d_mix.data_ptr(), flags=g.PEEK)
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
            for k in range(K):
                flush(k)
                ev[k][0].record()
                eng.process(d_x.data_ptr(), d_y.data_ptr(), d_mix.data_ptr(), flags=g.PEEK)
                ev[k][1].record()
            torch.cuda.synchronize(dev)
            res[tag + "_ms"] = float(np.median([a.elapsed_time(b) for a, b in ev]))
        res["strip_cost_us"] = (res["strip_ms"] - res["plain_ms"]) * 1e3
        fused[wl] = res
        eng.close()
        del d_x, d_y
    return {"strip": rows, "engine_with_full_strip": fused, "l2": flush.description,
            "note": "the strip is a dependent chain of B steps per track: ~12 cycles per sample for the biquad"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=40)
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="c2")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary workloads in the default line")
    ap.add_argument("--sweep", action="store_true", help="buffer-size sweep (BASELINE config 5): one JSON line with a table")
    ap.add_argument("--sweep-filter", default="", help="comma list of sweep points to run, e.g. upols:1024,upols:4096,direct")
    ap.add_argument("--strip", action="store_true", help="channel-strip kernels alone and attached to the engines")
    ap.add_argument("--latency", type=int, default=0, metavar="N", help="latency run: N blocks back-to-back and N periodic at B/fs")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        line = run_reference(args, rank, world)
        if line is not None:
            print(json.dumps(line), flush=True)
        return 0

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the convolution engine has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.strip:
        if rank == 0:
            print(json.dumps(run_strip(args, local_rank)), flush=True)
        return 0
    if args.latency > 0:
        res = run_latency(args, rank, world, local_rank, dist)
        if rank == 0:
            print(json.dumps(res), flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0
    if args.sweep:
        rows = run_sweep(args, rank, world, local_rank, dist)
        if rank == 0:
            print(json.dumps({"sweep": rows, "n_gpus": world, "fs": FS, "steps_per_point": args.steps,
                              "deadline_rule": "meets_deadline = p99 <= 1000*B/fs (cuda/globals.cu:86-89,155-156)"}), flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0 if all(r["bus_parity_ok"] for r in rows) else 3
    result = run_workload(args.workload, args, rank, world, local_rank, dist, want_cpu_baseline=(rank == 0 and world == 1))
    ok = result["parity"]["ok"]
    if not args.no_also and args.workload == "c2":
        also = {}
        for other in ("c3", "c4"):
            sub = argparse.Namespace(**vars(args))
            sub.steps = min(args.steps, 100)
            r = run_workload(other, sub, rank, world, local_rank, dist, want_cpu_baseline=False)
            also[other] = {k: r[k] for k in ("value", "unit", "ms_per_step", "rt_tracks", "latency_ms", "roofline", "e2e",
                                             "gpu_launches", "config", "run", "sustained", "parity")}
            ok = ok and r["parity"]["ok"]
        result["also"] = also
    if rank == 0:
        print(json.dumps(result), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not ok:
        print("bench.py: PARITY FAILED (see the \"parity\" objects of the line above)", file=sys.stderr, flush=True)
        return 3
    return 0


if __name__ == "__main__":
    sys.exit(main())
