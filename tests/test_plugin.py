"""The reference-shaped plugin surface: gpubench CLI + GPUABenchmark lifecycle (host/), CPU and GPU."""
import ctypes
import json
import os
import re
import subprocess

import numpy as np
import pytest

from gpuaudiobench_b200 import plugin

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_cli(*args):
    return subprocess.run([plugin.GPUBENCH, *args], capture_output=True, text=True, timeout=600)


# ---------------------------------------------------------------- CPU ------------------------
def test_plugin_library_exports_every_declared_symbol():
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "gpubench_plugin.h")).read(), flags=re.S)
    names = sorted(set(re.findall(r"\b(gpubench_[a-z0-9_]+)\s*\(", header)))
    assert len(names) >= 14, names
    lib = ctypes.CDLL(plugin.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), n


def test_cli_list_help_and_exit_codes():
    r = run_cli("--list")
    assert r.returncode == 0 and r.stdout.split("\n")[:5] == ["GPGPU Audio Benchmark", "Available benchmarks:", "Conv1D", "Conv1D_accel", "FFT1D"]
    r = run_cli("--help")
    assert r.returncode == 0
    for flag in ("--benchmark", "--fs", "--bufferSize", "--nTracks", "--nRuns", "--outputfile", "--json", "--irLen",
                 "--dawsim", "--dawsim-mode", "--dawsim-jitter-us"):
        assert flag in r.stdout
    r = run_cli("--nTracks")  # missing value: reference exits 1 (main.cu:276-279)
    assert r.returncode == 1 and "Error: --nTracks requires an argument" in r.stdout


def test_json_writer_matches_reference_text(golden):
    """generateJSONResults (globals.cu:124-182): byte-identical to the reference's own output."""
    got = plugin.json_results(golden["stats_lat"], "Conv1D", 48000, 512, 128)
    assert got == str(golden["stats_json"][0])
    parsed = json.loads(got)
    assert set(parsed) == {"benchmark", "configuration", "statistics", "deadline"}
    assert parsed["deadline"]["meets_deadline"] is True and abs(parsed["deadline"]["threshold_ms"] - 10.666667) < 1e-5


def test_statistics_match_reference(golden, oracle):
    st = plugin.statistics(golden["stats_lat"])
    got = np.array([st[k] for k in ("mean", "median", "std", "min", "max", "p95", "p99", "count")], dtype=np.float32)
    assert np.array_equal(got, golden["stats_out"])
    lat = (oracle.generate_input(33, 5) + 2).astype(np.float32)
    assert plugin.statistics(lat) == oracle.statistics(lat)


def test_unknown_benchmark_is_rejected():
    with pytest.raises(ValueError):
        plugin.Plugin("RndMemRead")


@pytest.mark.parametrize("sleep_mode", [False, True])
def test_dawsim_paces_at_the_buffer_period(sleep_mode):
    """DAWSimulator (Metal BenchmarkUtilities.swift:151-178): wake-ups land on multiples of the period."""
    period = 512 / 48000 / 4  # 2.67 ms
    t = plugin.dawsim_probe(period, 12, sleep_mode=sleep_mode)
    ideal = period * np.arange(1, 13)
    late = t - ideal
    assert np.all(late >= -1e-5), late                     # never early
    assert np.median(late) <= (1e-3 if sleep_mode else 1e-4), late  # on a shared CI box single wake-ups may be late
    assert np.all(np.diff(t) <= 3 * period)                # ... but the pace never collapses


def test_dawsim_jitter_is_bounded():
    period, jitter_us = 2e-3, 300.0
    t = plugin.dawsim_probe(period, 40, jitter_us=jitter_us)
    dev = t - period * np.arange(1, 41)
    assert np.all(dev >= -jitter_us * 1e-6 - 1e-5)                    # never earlier than -jitter
    assert np.percentile(dev, 60) <= jitter_us * 1e-6 + 3e-4, dev     # late only when the OS pre-empts the spinner
    assert dev.std() > 50e-6  # it does jitter


# ---------------------------------------------------------------- GPU ------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("T,B,L", [(128, 512, 1024), (1, 512, 1024), (16, 256, 3000)])
def test_conv1d_plugin_lifecycle(oracle, T, B, L):
    with plugin.Plugin("Conv1D", L, B, T) as p:
        p.setup()
        x = oracle.generate_input(T * B)
        assert np.array_equal(p.host_input().ravel(), x)                       # generateTestData(42)
        assert np.array_equal(p.host_ir(), oracle.generate_ir(T, L, "direct"))  # bench_conv1d.cu:159-178, bit-exact
        ref = oracle.r1(x, p.host_ir(), L, B, T)
        assert np.array_equal(p.cpu_reference(), ref)                          # the plugin's CPU loop == oracle R1
        wall, gpu = p.run(5, 3)
        assert (wall > 0).all() and (gpu > 0).all() and (gpu <= wall + 1e-3).all()
        v = p.validate()
        assert v["status"] == 0, v
        assert v["snr_db"] >= 100 and v["max_abs_err"] <= 1e-5 * v["ref_peak"]
        assert v["max_error"] <= 1e-3  # the reference's own abs check
        out = p.host_output()
    assert np.abs(out.astype(np.float64) - ref).max() <= 1e-5 * np.abs(ref).max()


@pytest.mark.gpu
@pytest.mark.parametrize("T,B,L", [(128, 512, 512), (1, 512, 1024), (64, 256, 4096)])
def test_conv1d_accel_plugin_lifecycle(oracle, T, B, L):
    with plugin.Plugin("Conv1D_accel", L, B, T) as p:
        p.setup()
        x = oracle.generate_input(T * B)
        assert np.array_equal(p.host_ir(), oracle.generate_ir(T, L, "accel"))   # bench_conv1d_accel.cu:152-165
        ref = oracle.r2(x, p.host_ir(), L, B, T)
        assert np.array_equal(p.cpu_reference(), ref)                           # sample-major [B][T]
        p.run(4, 3)
        v = p.validate()
        assert v["status"] == 0, v
        assert v["snr_db"] >= 90 and v["max_abs_err"] <= 1e-4 * v["ref_peak"]
        assert any("reference metric" in m for m in v["messages"])


@pytest.mark.gpu
@pytest.mark.parametrize("T,B", [(128, 512), (3, 1024), (16, 100)])
def test_fft1d_plugin_lifecycle(oracle, T, B):
    """FFT1D on the engine's Stockham R2C: the plugin's float DFT equals the oracle restatement of
    bench_fft.cu:149-168 bit for bit; the GPU result is judged against an fp64 FFT."""
    with plugin.Plugin("FFT1D", 0, B, T) as p:
        p.setup()
        x = p.fft_input()
        assert not x[:, min(B, 1024):].any() and x[:, :min(B, 1024)].any()  # zero padding of short buffers
        ref = p.fft_reference()
        for t in (0, T - 1):
            assert np.array_equal(ref[t].astype(np.complex64), oracle.fft_reference(x[t]).astype(np.complex64))
        wall, gpu = p.run(5, 3)
        assert (gpu > 0).all()
        v = p.validate()
        assert v["status"] == 0 and v["snr_db"] >= 110, v
        out = p.fft_output()
    truth = np.fft.rfft(x.astype(np.float64), axis=1)
    err = np.abs(out - truth).max()
    assert err <= 2e-5 * np.abs(truth).max(), err


@pytest.mark.gpu
@pytest.mark.parametrize("n", [32, 256, 1024, 8192])
def test_rfft_abi_matches_fp64_fft(n):
    """b200conv_rfft for every supported radix mix (log2(n/2) odd and even), cuFFT R2C layout."""
    import torch
    import gpuaudiobench_b200 as g
    count = 37
    x = torch.rand(count, n, device="cuda") * 2 - 1
    out = torch.zeros(count, n // 2 + 1, 2, device="cuda")
    g.rfft(x.data_ptr(), out.data_ptr(), count, n)
    torch.cuda.synchronize()
    got = out[..., 0].cpu().numpy() + 1j * out[..., 1].cpu().numpy()
    truth = np.fft.rfft(x.cpu().numpy().astype(np.float64), axis=1)
    snr = 10 * np.log10((np.abs(truth) ** 2).sum() / (np.abs(got - truth) ** 2).sum())
    assert snr >= 110, snr
    assert np.abs(got - truth).max() <= 3e-5 * np.abs(truth).max()
    assert not got[:, 0].imag.any() and not got[:, n // 2].imag.any()  # DC and Nyquist are real


@pytest.mark.gpu
def test_stream_mode_validates_after_streaming(oracle):
    with plugin.Plugin("Conv1D", 2048, 512, 8, stream_mode=True) as p:
        p.setup()
        p.run(6, 2)
        assert p.validate()["status"] == 0
    plugin.load_library().gpubench_set_globals(48000, 0, 0)


@pytest.mark.gpu
def test_dawsim_run_is_paced_and_meets_deadline():
    """128 tracks x 16k taps submitted every 10.67 ms: the loop takes ~n periods, p99 stays under it."""
    import time
    plugin.set_dawsim(True, sleep_mode=False, jitter_us=100.0)
    try:
        with plugin.Plugin("Conv1D", 16384, 512, 128) as p:
            p.setup()
            t0 = time.perf_counter()
            wall, gpu = p.run(20, 3)
            elapsed = time.perf_counter() - t0
            assert p.validate()["status"] == 0
    finally:
        plugin.set_dawsim(False)
    period = 512 / 48000
    assert 22 * period <= elapsed <= 24.5 * period, elapsed
    assert np.sort(wall)[-1] < period * 1e3  # every buffer inside its 10.67 ms period


@pytest.mark.gpu
def test_cli_end_to_end_json_and_csv(tmp_path):
    r = run_cli("--benchmark", "Conv1D", "--nTracks", "128", "--irLen", "16384", "--nRuns", "20", "--json")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Validation passed for Conv1D" in r.stdout and "Conv1D benchmark completed successfully!" in r.stdout
    js = json.loads(r.stdout[r.stdout.index("{\n"):r.stdout.rindex("}") + 1])
    assert js["benchmark"] == "Conv1D" and js["configuration"] == {"fs": 48000, "bufferSize": 512, "nTracks": 128, "nRuns": 20}
    assert js["deadline"]["meets_deadline"] is True and js["statistics"]["p99_ms"] < 10.667
    csv = tmp_path / "out.csv"
    r = run_cli("--benchmark", "Conv1D_accel", "--bufferSize", "256", "--nTracks", "64", "--irLen", "8192", "--nRuns", "10",
                "--outputfile", str(csv))
    assert r.returncode == 0 and "Validation passed for Conv1D_accel" in r.stdout
    lines = csv.read_text().strip().split("\n")
    assert lines[0] == "benchmark,fs,bufferSize,nTracks,nRuns,min_ms,max_ms,avg_ms,p50_ms,p95_ms,p99_ms,threshold_ms,meets_deadline"
    assert lines[1].startswith("Conv1D_accel,48000,256,64,10,") and lines[1].endswith(",true")
    assert os.path.exists("/tmp/Conv1D_accel_latencies.txt")


# ---------------------------------------------------------------- channel-strip plugins ------
def test_cli_lists_the_channel_strip_plugins():
    names = run_cli("--list").stdout.split("\n")
    for n in ("gain", "GainStats", "IIRFilter"):  # registry names of the reference, cuda/main.cu:85-93
        assert n in names


@pytest.mark.gpu
@pytest.mark.parametrize("T,B", [(128, 512), (1, 32), (37, 1000)])
def test_gain_and_gainstats_plugin_lifecycle(oracle, T, B):
    with plugin.Plugin("gain", 0, B, T) as p:
        p.setup()
        x = p.host_input()
        assert np.array_equal(x, oracle.generate_input(T * B).reshape(T, B))  # generateTestData(42)
        wall, gpu = p.run(5, warmup=2)
        assert (gpu > 0).all()
        v = p.validate()
        assert v["status"] == 0 and v["max_error"] == 0.0, v
        assert p.strip_bit_exact()
        assert np.array_equal(p.host_output(), oracle.gain(x, 2.0))
        assert np.array_equal(p.cpu_reference(), oracle.gain(x, 2.0))
    with plugin.Plugin("GainStats", 0, B, T) as p:
        p.setup()
        x = p.host_input()
        p.run(3, warmup=1)
        v = p.validate()
        assert v["status"] == 0 and p.strip_bit_exact(), v
        y_ref, s_ref = oracle.gainstats(x, 0.5)
        assert np.array_equal(p.host_output(), y_ref)
        assert np.array_equal(p.strip_stats(), s_ref) and np.array_equal(p.strip_stats(cpu=True), s_ref)


@pytest.mark.gpu
@pytest.mark.parametrize("T,B,iters", [(128, 512, 1), (128, 512, 7), (5, 96, 4)])
def test_iir_plugin_lifecycle_state_carried_across_iterations(oracle, T, B, iters):
    with plugin.Plugin("IIRFilter", 0, B, T) as p:
        p.setup()
        coef = p.strip_coefficients()
        assert np.array_equal(coef, oracle.butterworth(0.25))
        x = p.host_input()
        for _ in range(iters):
            p.iterate()
        v = p.validate()
        assert v["status"] == 0 and p.strip_bit_exact(), v
        st = np.zeros((T, 2), np.float32)
        for _ in range(iters):  # the filter state persists on the device (cuda/bench_iir.cu:42-43)
            ref = oracle.iir(x, coef, st)
        assert np.array_equal(p.host_output(), ref)
        assert np.array_equal(p.strip_state(), st) and np.array_equal(p.strip_state(cpu=True), st)


@pytest.mark.gpu
def test_cli_runs_the_channel_strip_plugins():
    for name in ("gain", "GainStats", "IIRFilter"):
        r = run_cli("--benchmark", name, "--nTracks", "256", "--nRuns", "10", "--json")
        assert r.returncode == 0, r.stdout + r.stderr
        assert f"Validation passed for {name}" in r.stdout and "bit-identical to the CPU loop" in r.stdout
        js = json.loads(r.stdout[r.stdout.index("{\n"):r.stdout.rindex("}") + 1])
        assert js["benchmark"] == name and js["deadline"]["meets_deadline"] is True
