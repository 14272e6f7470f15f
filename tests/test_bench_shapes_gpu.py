"""Parity at the EXACT shapes bench.py measures (-m gpu), through the same call bench.py times
(b200conv_process on device buffers, with the stereo bus).

The shape decides the launch plan — partition split S, tap-group width A, pipeline depth, FFT path —
so a parity test on a small cousin of a benchmarked shape does not pin the benchmarked kernel
configuration.  Each case here asserts the plan it is meant to cover, streams >= ceil(L/B) + 2 blocks
so every partition / tap stage carries signal, and compares
    * a few tracks against the streaming oracle (the reference's own loop),
    * more tracks against an fp64 FFT convolution,
    * the stereo bus of the last block against the fp64 sum over ALL tracks.
Tolerances as in test_parity_gpu.py (direct 100 dB / 1e-5, UPOLS 90 dB / 1e-4 of max|y_ref|).
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest
import torch

import gpuaudiobench_b200 as g
from gpuaudiobench_b200 import synth

pytestmark = pytest.mark.gpu

TOL = {g.ALGO_DIRECT: (100.0, 1e-5), g.ALGO_UPOLS: (90.0, 1e-4)}


def snr_db(got, ref):
    ref64 = np.asarray(ref, dtype=np.float64)
    err = np.sum((np.asarray(got, dtype=np.float64) - ref64) ** 2)
    return 10 * np.log10(np.sum(ref64 ** 2) / max(err, 1e-300))


def fp64_truth(x, h):
    from scipy.signal import fftconvolve
    return fftconvolve(x.astype(np.float64), h.astype(np.float64))[:x.size]


def stream_on_device(engine, xs, keep, layout=g.OUT_TRACK_MAJOR):
    """xs [M][T][B] (host) -> (y of the tracks in `keep` [len(keep)][M*B], all tracks' last block [T][B],
    bus of the last block [2][B]) through b200conv_process on device buffers, as bench.py calls it."""
    M, T, B = xs.shape
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream(dev)
    d_x = torch.from_numpy(xs).to(dev)
    shape = (B, engine.Tg) if layout == g.OUT_SAMPLE_MAJOR else (T, B)
    d_y = torch.zeros(shape, device=dev)
    d_mix = torch.zeros(2, B, device=dev)
    idx = torch.tensor(keep, device=dev)
    kept = torch.zeros(len(keep), M * B, device=dev)
    for m in range(M):
        engine.process(d_x[m].data_ptr(), d_y.data_ptr(), d_mix.data_ptr(), stream=stream.cuda_stream)
        if layout == g.OUT_SAMPLE_MAJOR:
            kept[:, m * B:(m + 1) * B] = d_y[:, engine.toff + idx].T
        else:
            kept[:, m * B:(m + 1) * B] = d_y[idx]
    torch.cuda.synchronize(dev)
    last = d_y.cpu().numpy()
    if layout == g.OUT_SAMPLE_MAJOR:
        last = np.ascontiguousarray(last[:, engine.toff:engine.toff + T].T)
    return kept.cpu().numpy(), last, d_mix.cpu().numpy()


def check_shape(oracle, algo, T, B, L, oracle_tracks, fp64_tracks, Tg=None, t0=0, layout=g.OUT_TRACK_MAJOR, seed=5,
                expect_plan=None, flags=0):
    Tg = Tg or T
    P = (L + B - 1) // B
    M = P + 2
    rng = np.random.default_rng(seed)
    xs = rng.uniform(-1, 1, size=(M, T, B)).astype(np.float32)
    h = synth.make_ir(Tg, L, t0, t0 + T)
    plan = g.plan(T, B, L, algo, flags=flags)
    if expect_plan:
        for k, v in expect_plan.items():
            assert plan[k] == v, f"plan[{k}] = {plan[k]}, test was written for {v}: {plan}"
    keep = sorted(set(oracle_tracks) | set(fp64_tracks))
    with g.ConvEngine(T, B, L, algo, layout, track_offset=t0, total_tracks=Tg, flags=flags) as e:
        e.load_ir(h)
        got, last, bus = stream_on_device(e, xs, keep, layout)
    min_snr, rel = TOL[algo]
    with ThreadPoolExecutor(max_workers=8) as pool:  # the oracle loop releases the GIL
        want = list(pool.map(lambda t: oracle.stream(xs[:, t, :].ravel(), h[t]), oracle_tracks))
    for t, w in zip(oracle_tracks, want):
        row = got[keep.index(t)]
        s = snr_db(row, w)
        mx = np.abs(row.astype(np.float64) - w).max()
        assert s >= min_snr and mx <= rel * np.abs(w).max(), f"track {t} vs oracle: {s:.1f} dB, max|err| {mx:.3e}"
        assert snr_db(row[-B:], w[-B:]) >= min_snr, f"track {t}, last block"
    for t in fp64_tracks:
        truth = fp64_truth(xs[:, t, :].ravel(), h[t])
        row = got[keep.index(t)]
        assert snr_db(row[-2 * B:], truth[-2 * B:]) >= min_snr - 2, f"track {t} vs fp64"
    # bus of the last block: fp64 over ALL tracks (from the engine's own last-block outputs, so only the
    # bus arithmetic is under test here; the outputs themselves are pinned above)
    theta = (np.arange(t0, t0 + T) + 0.5) / Tg * np.pi / 2
    gains = np.stack([np.cos(theta), np.sin(theta)]) / np.sqrt(Tg)
    assert snr_db(bus, gains @ last.astype(np.float64)) >= 110
    return plan


def test_c4_shard_at_the_benchmarked_shape(oracle):
    """bench.py c4: 512 tracks/GPU x 512 x 96000 (global tracks [1024, 1536) of 4096): the planner splits the
    188 partitions in S = 2 -> 1024 CTAs; the round-1 test ran 16 tracks (S = 11)."""
    check_shape(oracle, g.ALGO_UPOLS, 512, 512, 96000, oracle_tracks=[0, 511], fp64_tracks=list(range(0, 512, 32)),
                Tg=4096, t0=1024, expect_plan={"P": 188, "S": 2})


def test_c3_at_the_benchmarked_shape_sample_major(oracle):
    check_shape(oracle, g.ALGO_UPOLS, 1024, 256, 65536, oracle_tracks=[1023], fp64_tracks=list(range(0, 1024, 128)),
                layout=g.OUT_SAMPLE_MAJOR, expect_plan={"P": 256})


@pytest.mark.parametrize("B", [32, 4096])
def test_upols_sweep_ends_at_96000_taps(oracle, B):
    """BASELINE config 5 at C4's T and L: B = 32 -> P = 3000 (16 partition lanes per bin pair),
    B = 4096 -> P = 24 on the large-FFT path."""
    check_shape(oracle, g.ALGO_UPOLS, 512, B, 96000, oracle_tracks=[3], fp64_tracks=[0, 255, 511],
                expect_plan={"P": (96000 + B - 1) // B})


@pytest.mark.parametrize("B", [1024, 2048])
def test_upols_large_buffers_at_96000_taps(oracle, B):
    check_shape(oracle, g.ALGO_UPOLS, 512, B, 96000, oracle_tracks=[], fp64_tracks=[0, 100, 511])


@pytest.mark.parametrize("B,A", [(32, 2), (64, 4), (128, 8), (256, 16), (512, 32), (4096, 32)])
def test_ffma_direct_sweep_points_at_16384_taps(oracle, B, A):
    """The FFMA direct kernel's sweep points at C2's T and L: the small-buffer plans (A = 2 / 4 / 8 / 16 tap-group
    layouts, swizzled taps) were only tested at L <= 960 in round 1."""
    check_shape(oracle, g.ALGO_DIRECT, 128, B, 16384, oracle_tracks=[0, 127], fp64_tracks=[1, 64, 126],
                expect_plan={"A": A, "impl": g.ALGO_DIRECT}, flags=g.engine.FLAG_FFMA_ONLY)


@pytest.mark.parametrize("B,impl", [(32, g.ALGO_DIRECT), (64, g.ALGO_DIRECT), (128, g.ALGO_DIRECT_TC), (256, g.ALGO_DIRECT_TC),
                                    (512, g.ALGO_DIRECT_TC), (1024, g.ALGO_DIRECT_TC), (2048, g.ALGO_DIRECT_TC),
                                    (4096, g.ALGO_DIRECT_TC)])
def test_direct_sweep_points_as_the_planner_dispatches_them(oracle, B, impl):
    """The same sweep points through plain ALGO_DIRECT, i.e. what bench.py --sweep measures: the tensor-core kernel for
    B >= 128 (blocks over 1024 samples stream through it in sub-blocks of 1024), the FFMA kernel below; one tolerance
    for both (100 dB, 1e-5 of max|y|)."""
    check_shape(oracle, g.ALGO_DIRECT, 128, B, 16384, oracle_tracks=[0, 127], fp64_tracks=[1, 64, 126],
                expect_plan={"impl": impl})


def test_two_sharded_engines_on_one_device_equal_the_unsharded_job(oracle):
    """Track sharding as the multi-GPU path does it (track_offset / total_tracks, global IR and pan
    indices, sample-major column tiles), but with both shards on ONE device so that the driver's 1-GPU
    box covers it: outputs equal the unsharded engine's (to fp32 re-association), the two bus partials add up to
    the unsharded bus."""
    Tg, B, L, M = 24, 256, 3000, 14
    xs = oracle.generate_input(M * Tg * B, 21).reshape(M, Tg, B)
    h = synth.make_ir(Tg, L, 0, Tg)
    for algo in (g.ALGO_DIRECT, g.ALGO_UPOLS):
        for layout in (g.OUT_TRACK_MAJOR, g.OUT_SAMPLE_MAJOR):
            with g.ConvEngine(Tg, B, L, algo, layout) as whole, \
                    g.ConvEngine(10, B, L, algo, layout, track_offset=0, total_tracks=Tg) as lo, \
                    g.ConvEngine(14, B, L, algo, layout, track_offset=10, total_tracks=Tg) as hi:
                whole.load_ir(h)
                lo.load_ir(h[:10])
                hi.load_ir(h[10:])
                for m in range(M):
                    y, bus = whole.process_host(xs[m], want_mix=True)
                    y0, b0 = lo.process_host(xs[m, :10], want_mix=True)
                    y1, b1 = hi.process_host(xs[m, 10:], want_mix=True)
                    if layout == g.OUT_SAMPLE_MAJOR:  # each shard wrote its own columns of [B][Tg]
                        assert not y0[:, 10:].any() and not y1[:, :10].any()
                        stitched = y0 + y1
                    else:
                        stitched = np.concatenate([y0, y1])
                    # (not bit-equal in general: the direct engine's span schedule depends on the track count)
                    assert snr_db(stitched, y) >= 125 or np.array_equal(stitched, y), f"algo {algo} layout {layout} block {m}"
                    assert snr_db(b0.astype(np.float64) + b1, bus) >= 120
            want = np.stack([oracle.stream(xs[:, t, :].ravel(), h[t]) for t in (0, 9, 10, 23)])
            rows = y if layout == g.OUT_TRACK_MAJOR else y.T
            for i, t in enumerate((0, 9, 10, 23)):
                assert snr_db(rows[t], want[i][-B:]) >= TOL[algo][0]
