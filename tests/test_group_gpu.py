"""b200conv_group_* (one process, one engine per GPU) on 2 real GPUs; skipped on a 1-GPU box."""
import numpy as np
import pytest
import torch

import gpuaudiobench_b200 as g
from gpuaudiobench_b200 import plugin

pytestmark = pytest.mark.gpu
needs2 = pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")


def snr_db(got, ref):
    ref = np.asarray(ref, dtype=np.float64)
    return 10 * np.log10((ref ** 2).sum() / max(((np.asarray(got, dtype=np.float64) - ref) ** 2).sum(), 1e-300))


@needs2
@pytest.mark.parametrize("algo", [g.ALGO_DIRECT, g.ALGO_UPOLS])
@pytest.mark.parametrize("layout", [g.OUT_TRACK_MAJOR, g.OUT_SAMPLE_MAJOR])
def test_two_gpu_group_matches_streaming_oracle_and_bus(oracle, algo, layout):
    Tg, B, L, M = 21, 256, 3000, 9  # 21 tracks -> uneven shards 10 + 11
    xs = oracle.generate_input(M * Tg * B, 13).reshape(M, Tg, B)
    h = oracle.generate_ir(Tg, L, "accel")
    want = np.stack([oracle.stream(xs[:, t, :].ravel(), h[t]) for t in range(Tg)])
    theta = (np.arange(Tg) + 0.5) / Tg * np.pi / 2
    gains = np.stack([np.cos(theta), np.sin(theta)], axis=1) / np.sqrt(Tg)
    bus = gains.T @ want.astype(np.float64)
    with g.ConvGroup(Tg, B, L, algo, 2, layout) as grp:
        grp.load_ir(h)
        outs = [grp.process_host(xs[m]) for m in range(M)]
    got = np.concatenate([(y.T if layout == g.OUT_SAMPLE_MAJOR else y) for y, _ in outs], axis=1)
    got_bus = np.concatenate([m for _, m in outs], axis=1)
    min_snr = 100 if algo == g.ALGO_DIRECT else 90
    assert snr_db(got, want) >= min_snr
    assert snr_db(got_bus, bus) >= 90


@needs2
@pytest.mark.parametrize("tensor_cores", [False, True])
def test_two_gpu_group_with_full_pipeline_depth_and_opt_in_shared_memory(oracle, monkeypatch, tensor_cores):
    """256 tracks x 16384 taps x 512 on 2 GPUs: every member's launch needs the > 48 KB dynamic shared memory
    opt-in (FFMA kernel: nbuf = 4, 59.5 KB per CTA; tensor-core kernel: 106 KB) — the attribute is per DEVICE, and
    round 1 set it once per process, so device 1 never got it.  Also the first 2-GPU job whose spans cross tracks
    and whose bus tree has several groups per member."""
    from scipy.signal import fftconvolve
    Tg, B, L, M = 256, 512, 16384, 35
    monkeypatch.setenv("B200CONV_DIRECT_TC", "1" if tensor_cores else "0")
    p = g.plan(Tg // 2, B, L, g.ALGO_DIRECT)
    assert p["impl"] == (g.ALGO_DIRECT_TC if tensor_cores else g.ALGO_DIRECT) and p["smem"] > 48 * 1024, p
    assert tensor_cores or p["nbuf"] == 4
    rng = np.random.default_rng(3)
    xs = rng.uniform(-1, 1, size=(M, Tg, B)).astype(np.float32)
    h = oracle.generate_ir(Tg, L, "direct")
    with g.ConvGroup(Tg, B, L, g.ALGO_DIRECT, 2) as grp:
        grp.load_ir(h)
        outs = [grp.process_host(xs[m]) for m in range(M)]
    got = np.concatenate([y for y, _ in outs], axis=1)
    for t in (0, 127, 128, 255):
        want = oracle.stream(xs[:, t, :].ravel(), h[t])
        assert snr_db(got[t], want) >= 100, t
    truth = np.stack([fftconvolve(xs[:, t, :].ravel().astype(np.float64), h[t].astype(np.float64))[:M * B][-B:] for t in range(Tg)])
    assert snr_db(got[:, -B:], truth) >= 100
    theta = (np.arange(Tg) + 0.5) / Tg * np.pi / 2
    gains = np.stack([np.cos(theta), np.sin(theta)]) / np.sqrt(Tg)
    assert snr_db(outs[-1][1], gains @ truth) >= 95


@needs2
def test_plugin_on_two_gpus_validates_against_r1_and_r2():
    plugin.set_ngpus(2)
    try:
        for name, L in (("Conv1D", 4096), ("Conv1D_accel", 4096)):
            with plugin.Plugin(name, L, 512, 64) as p:
                p.setup()
                p.run(5, 3)
                v = p.validate()
                assert v["status"] == 0, (name, v)
    finally:
        plugin.set_ngpus(1)


@pytest.mark.parametrize("n_gpus", [1, 2])
@pytest.mark.parametrize("algo,layout", [(g.ALGO_DIRECT, g.OUT_TRACK_MAJOR), (g.ALGO_UPOLS, g.OUT_SAMPLE_MAJOR)])
def test_group_channel_strip_is_sliced_per_member(oracle, n_gpus, algo, layout):
    """b200conv_group_set_strip: per-track gains / biquads reach the right member; outputs, statistics and
    filter state are bit-identical to the oracle's strip of the plain group's output (1 GPU runs everywhere)."""
    if torch.cuda.device_count() < n_gpus:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    Tg, B, L, M = 13, 128, 500, 3
    xs = oracle.generate_input(M * Tg * B, 17).reshape(M, Tg, B)
    h = oracle.generate_ir(Tg, L, "accel")
    coef = np.stack([oracle.butterworth(0.05 + 0.03 * t) for t in range(Tg)])
    gains = np.linspace(0.5, 1.5, Tg).astype(np.float32)
    with g.ConvGroup(Tg, B, L, algo, n_gpus, layout) as plain, g.ConvGroup(Tg, B, L, algo, n_gpus, layout) as strip:
        plain.load_ir(h)
        strip.load_ir(h)
        strip.set_strip(g.STRIP_STATS | g.STRIP_GAIN | g.STRIP_BIQUAD, gains=gains, biquad=coef)
        state = np.zeros((Tg, 2), np.float32)
        for m in range(M):
            y0, _ = plain.process_host(xs[m])
            y1, _ = strip.process_host(xs[m])
            tm = np.ascontiguousarray(y0.T if layout == g.OUT_SAMPLE_MAJOR else y0)
            ref, stats_ref = oracle.strip(tm, 7, gains=gains, coeffs=coef, state=state)
            assert np.array_equal(y1.T if layout == g.OUT_SAMPLE_MAJOR else y1, ref), f"block {m}"
            assert np.array_equal(strip.strip_stats(), stats_ref) and np.array_equal(strip.strip_state(), state)


def test_group_rejects_more_gpus_than_visible():
    with pytest.raises(g.B200ConvError) as ei:
        g.ConvGroup(64, 512, 1024, g.ALGO_DIRECT, torch.cuda.device_count() + 1)
    assert ei.value.code == g.engine.ERR_NO_DEVICE
