"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle.

Tolerances (fp32 path; SURVEY.md App. A.4 — the reference's own abs-1e-3 / rel-1e-3 checks are
vacuous at the 1/L-scaled signal level, so SNR and max-abs relative to max|y_ref| are stated):
    direct engine : SNR >= 100 dB and max|err| <= 1e-5 * max|y_ref|   vs the fp32 oracle
    UPOLS engine  : SNR >=  90 dB and max|err| <= 1e-4 * max|y_ref|   vs the fp32 oracle
Indexing (track <-> column, tap <-> partition, block <-> ring slot) is checked exactly with impulses.
The reference's two validation metrics (bench_base.cu:193-222 abs, bench_conv1d_accel.cu:312-336
rel) are evaluated alongside for continuity.
"""
import numpy as np
import pytest

import gpuaudiobench_b200 as g

pytestmark = pytest.mark.gpu

TOL = {g.ALGO_DIRECT: (100.0, 1e-5), g.ALGO_UPOLS: (90.0, 1e-4)}
NAME = {g.ALGO_DIRECT: "direct", g.ALGO_UPOLS: "upols"}


def snr_db(got, ref):
    ref64 = np.asarray(ref, dtype=np.float64)
    err = np.sum((np.asarray(got, dtype=np.float64) - ref64) ** 2)
    return 10 * np.log10(np.sum(ref64 ** 2) / max(err, 1e-300))


def assert_parity(got, ref, algo, what=""):
    min_snr, rel = TOL[algo]
    s = snr_db(got, ref)
    mx = np.abs(np.asarray(got, dtype=np.float64) - ref).max()
    scale = np.abs(ref).max()
    assert s >= min_snr and mx <= rel * scale, f"{NAME[algo]} {what}: SNR {s:.1f} dB, max|err| {mx:.3e} vs scale {scale:.3e}"
    return s


def run_stream(engine, xs, want_mix=False):
    """xs [M][T][B] -> list of per-block outputs."""
    return [engine.process_host(xs[m], want_mix=want_mix) for m in range(xs.shape[0])]


# ---------------------------------------------------------------------------------------------
# Golden config C1 and the reference's two oracles
# ---------------------------------------------------------------------------------------------
def test_c1_direct_matches_r1_and_golden(oracle, golden):
    T, B, L = 1, 512, 1024
    x = oracle.generate_input(T * B).reshape(T, B)
    h = oracle.generate_ir(T, L, "direct")
    with g.ConvEngine(T, B, L, g.ALGO_DIRECT) as e:
        e.load_ir(h)
        y, _ = e.process_host(x)
    ref = oracle.r1(x, h, L, B, T)
    assert np.array_equal(ref, golden["c1_r1"])
    assert_parity(y, ref, g.ALGO_DIRECT, "C1 vs R1")
    mx, mean, over = oracle.compare_abs(y, ref, 1e-3)  # the reference's own check (bench_conv1d.cu:108)
    assert over == 0


def test_c1_upols_matches_r2_and_golden(oracle, golden):
    T, B, L = 1, 512, 1024
    x = oracle.generate_input(T * B).reshape(T, B)
    h = oracle.generate_ir(T, L, "accel")
    with g.ConvEngine(T, B, L, g.ALGO_UPOLS, g.OUT_SAMPLE_MAJOR) as e:
        e.load_ir(h)
        y, _ = e.process_host(x)
    ref = oracle.r2(x, h, L, B, T)
    assert np.array_equal(ref, golden["c1_r2"])
    assert y.shape == (B, T)
    assert_parity(y, ref, g.ALGO_UPOLS, "C1 vs R2")


@pytest.mark.parametrize("algo", [g.ALGO_DIRECT, g.ALGO_UPOLS])
@pytest.mark.parametrize("T,B,L", [(128, 512, 1024), (5, 64, 200), (3, 32, 40)])
def test_r2_first_block_after_reset_sample_major(oracle, algo, T, B, L):
    """Zero history + interleaved output: out[T*n + t] (bench_conv1d_accel.cu:249)."""
    x = oracle.generate_input(T * B).reshape(T, B)
    h = oracle.generate_ir(T, L, "accel")
    with g.ConvEngine(T, B, L, algo, g.OUT_SAMPLE_MAJOR) as e:
        e.load_ir(h)
        y, _ = e.process_host(x)
        e.reset()
        y2, _ = e.process_host(x)
    ref = oracle.r2(x, h, L, B, T)
    assert_parity(y, ref, algo, "block 0 vs R2")
    assert np.array_equal(y, y2), "reset() must restore the zero-history state exactly"


@pytest.mark.parametrize("algo", [g.ALGO_DIRECT, g.ALGO_UPOLS])
@pytest.mark.parametrize("T,B,L", [(128, 512, 1024), (16, 512, 4096), (6, 64, 200)])
def test_r1_with_primed_history_track_major(oracle, algo, T, B, L):
    """R1's flat-index bleed == priming track t with x_flat[tB-L+1 .. tB-1]."""
    x = oracle.generate_input(T * B)
    h = oracle.generate_ir(T, L, "direct")
    padded = np.concatenate([np.zeros(L - 1, dtype=np.float32), x])
    hist = np.stack([padded[t * B:t * B + L - 1] for t in range(T)])
    with g.ConvEngine(T, B, L, algo) as e:
        e.load_ir(h)
        e.prime_history(hist)
        y, _ = e.process_host(x.reshape(T, B), flags=g.PEEK)
        y2, _ = e.process_host(x.reshape(T, B), flags=g.PEEK)
    ref = oracle.r1(x, h, L, B, T)
    assert_parity(y, ref, algo, "primed block vs R1")
    assert np.array_equal(y, y2), "PEEK must not advance the stream state"


# ---------------------------------------------------------------------------------------------
# Streaming oracle: every partition / tap range is exercised, ragged sizes, both sweep ends
# ---------------------------------------------------------------------------------------------
STREAM_CASES = [
    # T, B, L, blocks
    (4, 256, 1000, 7),     # L not a multiple of B
    (3, 32, 960, 70),      # B at the small sweep end, > 2P blocks of ring rotation
    (2, 4096, 6000, 4),    # B at the large sweep end
    (5, 128, 128, 4),      # P = 1
    (2, 64, 1, 3),         # single tap
    (3, 512, 5000, 14),    # C2-like tile shape, several tap stages
    (2, 1024, 3000, 5),    # two output tiles per track (direct)
]


@pytest.mark.parametrize("algo", [g.ALGO_DIRECT, g.ALGO_UPOLS])
@pytest.mark.parametrize("T,B,L,M", STREAM_CASES)
def test_streaming_blocks_match_oracle(oracle, algo, T, B, L, M):
    xs = oracle.generate_input(M * T * B, 7).reshape(M, T, B)
    if L == 1:  # the reference IR formula divides by L-1 (bench_conv1d.cu:169): use a plain gain tap
        h = np.array([[0.5]] * T, dtype=np.float32) * np.arange(1, T + 1, dtype=np.float32)[:, None]
    else:
        h = oracle.generate_ir(T, L, "accel")
    want = np.stack([oracle.stream(xs[:, t, :].ravel(), h[t]) for t in range(T)])  # [T][M*B]
    with g.ConvEngine(T, B, L, algo) as e:
        e.load_ir(h)
        got = np.concatenate([y for y, _ in run_stream(e, xs)], axis=1)
        assert e.query()["blocks_processed"] == M
    assert_parity(got, want, algo, f"stream T={T} B={B} L={L}")
    # last block alone too (late ring slots / tap stages must be as good as the first)
    assert_parity(got[:, -B:], want[:, -B:], algo, "last block")


@pytest.mark.parametrize("ctas_per_sm,sps", [(1, 8), (2, 4), (2, 8)])
def test_direct_schedules_agree(oracle, monkeypatch, ctas_per_sm, sps):
    """The persistent span schedule (grid size, stage depth) must not change the result beyond fp32
    re-association: 37 tracks x 9000 taps gives spans that start and end mid-track."""
    monkeypatch.setenv("B200CONV_DIRECT_CTAS_PER_SM", str(ctas_per_sm))
    monkeypatch.setenv("B200CONV_DIRECT_SPS", str(sps))
    T, B, L, M = 37, 512, 9000, 3
    xs = oracle.generate_input(M * T * B, 3).reshape(M, T, B)
    h = oracle.generate_ir(T, L, "direct")
    want = np.stack([oracle.stream(xs[:, t, :].ravel(), h[t]) for t in range(T)])
    with g.ConvEngine(T, B, L, g.ALGO_DIRECT) as e:
        e.load_ir(h)
        got = np.concatenate([y for y, _ in run_stream(e, xs)], axis=1)
    assert_parity(got, want, g.ALGO_DIRECT, f"ctas/SM {ctas_per_sm} sps {sps}")


@pytest.mark.parametrize("split", [1, 2, 5])
def test_upols_partition_splits_agree(oracle, monkeypatch, split):
    monkeypatch.setenv("B200CONV_UPOLS_SPLIT", str(split))
    T, B, L, M = 3, 128, 2000, 20
    xs = oracle.generate_input(M * T * B, 4).reshape(M, T, B)
    h = oracle.generate_ir(T, L, "accel")
    want = np.stack([oracle.stream(xs[:, t, :].ravel(), h[t]) for t in range(T)])
    with g.ConvEngine(T, B, L, g.ALGO_UPOLS) as e:
        e.load_ir(h)
        got = np.concatenate([y for y, _ in run_stream(e, xs)], axis=1)
    assert_parity(got, want, g.ALGO_UPOLS, f"split {split}")


def _fuzz_cases(n, seed):
    rng = np.random.default_rng(seed)
    cases = []
    for _ in range(n):
        B = int(rng.choice([32, 64, 128, 256, 512, 1024]))
        T = int(rng.integers(1, 41))
        L = int(rng.integers(2, 6000))
        M = int(rng.integers(2, 7)) if B * L < 2_000_000 else 2
        cases.append((T, B, L, M, int(rng.integers(0, 2)), int(rng.integers(0, 2))))
    return cases


@pytest.mark.parametrize("T,B,L,M,algo,layout", _fuzz_cases(24, seed=2026))
def test_fuzz_random_shapes_match_streaming_oracle(oracle, T, B, L, M, algo, layout):
    """Random track counts, buffer sizes, IR lengths (never a multiple of anything) and stream
    lengths, both engines, both layouts, with the stereo bus, against the reference loop."""
    xs = oracle.generate_input(M * T * B, 1000 + T + L).reshape(M, T, B)
    h = oracle.generate_ir(T, L, "accel")
    want = np.stack([oracle.stream(xs[:, t, :].ravel(), h[t]) for t in range(T)])
    theta = (np.arange(T) + 0.5) / T * np.pi / 2
    gains = np.stack([np.cos(theta), np.sin(theta)], axis=1) / np.sqrt(T)
    bus = gains.T @ want.astype(np.float64)
    with g.ConvEngine(T, B, L, algo, layout) as e:
        e.load_ir(h)
        outs = run_stream(e, xs, want_mix=True)
    got = np.concatenate([(y.T if layout == g.OUT_SAMPLE_MAJOR else y) for y, _ in outs], axis=1)
    got_bus = np.concatenate([m for _, m in outs], axis=1)
    assert_parity(got, want, algo, f"fuzz T={T} B={B} L={L} M={M} layout={layout}")
    assert snr_db(got_bus, bus) >= 90


# ---------------------------------------------------------------------------------------------
# Exact indexing: impulses (SURVEY App. E "Exact indexing tests")
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("algo", [g.ALGO_DIRECT, g.ALGO_UPOLS])
def test_delta_ir_delays_a_ramp_exactly(algo):
    """h = delta[j0] with j0 on partition / tap-block boundaries: output must be the delayed input.
    Direct: bit-exact.  UPOLS: <= 1e-6 of full scale (an fp32 FFT does not reproduce integers)."""
    T, B, L, M = 6, 64, 200, 9
    j0s = [0, B - 1, B, L - 1, 15, 16]
    h = np.zeros((T, L), dtype=np.float32)
    for t, j0 in enumerate(j0s):
        h[t, j0] = 1.0
    n = np.arange(M * B)
    xs = np.stack([((n * (t + 1)) % 97 - 48).astype(np.float32) for t in range(T)])  # small integers
    with g.ConvEngine(T, B, L, algo) as e:
        e.load_ir(h)
        got = np.concatenate([e.process_host(xs[:, m * B:(m + 1) * B])[0] for m in range(M)], axis=1)
    for t, j0 in enumerate(j0s):
        want = np.concatenate([np.zeros(j0, dtype=np.float32), xs[t, :M * B - j0]])
        if algo == g.ALGO_DIRECT:
            assert np.array_equal(got[t], want), f"track {t} delay {j0}"
        else:
            assert np.abs(got[t] - want).max() <= 1e-6 * 48 * 4, f"track {t} delay {j0}"


@pytest.mark.parametrize("algo", [g.ALGO_DIRECT, g.ALGO_UPOLS])
@pytest.mark.parametrize("layout", [g.OUT_TRACK_MAJOR, g.OUT_SAMPLE_MAJOR])
def test_track_isolation_and_column_mapping(oracle, algo, layout):
    """Only one track carries signal: every other track/column must be exactly zero."""
    T, B, L, M, hot = 9, 128, 300, 4, 5
    h = oracle.generate_ir(T, L, "accel")
    xs = np.zeros((M, T, B), dtype=np.float32)
    xs[:, hot, :] = oracle.generate_input(M * B, 2).reshape(M, B)
    want = oracle.stream(xs[:, hot, :].ravel(), h[hot]).reshape(M, B)
    with g.ConvEngine(T, B, L, algo, layout) as e:
        e.load_ir(h)
        for m in range(M):
            y, _ = e.process_host(xs[m])
            yt = y.T if layout == g.OUT_SAMPLE_MAJOR else y
            assert_parity(yt[hot], want[m], algo, f"hot track block {m}")
            cold = np.delete(yt, hot, axis=0)
            assert not cold.any(), "signal leaked into a silent track"


def test_sample_major_column_tile_of_a_sharded_job(oracle):
    """An engine owning tracks [3, 7) of a 10-track job writes only its columns of [B][10]."""
    Tg, t0, t1, B, L = 10, 3, 7, 64, 100
    x = oracle.generate_input(Tg * B).reshape(Tg, B)
    h = oracle.generate_ir(Tg, L, "accel")
    ref = oracle.r2(x, h, L, B, Tg)
    for algo in (g.ALGO_DIRECT, g.ALGO_UPOLS):
        with g.ConvEngine(t1 - t0, B, L, algo, g.OUT_SAMPLE_MAJOR, track_offset=t0, total_tracks=Tg) as e:
            e.load_ir(h[t0:t1])
            y, _ = e.process_host(x[t0:t1])
        assert y.shape == (B, Tg)
        assert_parity(y[:, t0:t1], ref[:, t0:t1], algo, "column tile")
        assert not y[:, :t0].any() and not y[:, t1:].any()


# ---------------------------------------------------------------------------------------------
# Mix bus (new; oracle = fp64 sum of the per-track oracle outputs)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("algo", [g.ALGO_DIRECT, g.ALGO_UPOLS])
@pytest.mark.parametrize("layout", [g.OUT_TRACK_MAJOR, g.OUT_SAMPLE_MAJOR])
def test_stereo_mix_bus(oracle, algo, layout):
    T, B, L, M = 70, 128, 500, 5
    xs = oracle.generate_input(M * T * B, 5).reshape(M, T, B)
    h = oracle.generate_ir(T, L, "accel")
    want = np.stack([oracle.stream(xs[:, t, :].ravel(), h[t]) for t in range(T)]).astype(np.float64)
    theta = (np.arange(T) + 0.5) / T * np.pi / 2
    gains = np.stack([np.cos(theta), np.sin(theta)], axis=1) / np.sqrt(T)
    bus = gains.T @ want  # [2][M*B]
    with g.ConvEngine(T, B, L, algo, layout) as e:
        e.load_ir(h)
        got = np.concatenate([mix for _, mix in run_stream(e, xs, want_mix=True)], axis=1)
    assert snr_db(got, bus) >= 90 and np.abs(got - bus).max() <= 1e-4 * np.abs(bus).max()


# ---------------------------------------------------------------------------------------------
# Device-pointer entry point on a caller stream (what bench.py and the plugin use)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("algo", [g.ALGO_DIRECT, g.ALGO_UPOLS])
def test_process_with_device_buffers_on_torch_stream(oracle, algo):
    import torch
    T, B, L, M = 8, 256, 2000, 10
    xs = oracle.generate_input(M * T * B, 8).reshape(M, T, B)
    h = oracle.generate_ir(T, L, "accel")
    want = np.stack([oracle.stream(xs[:, t, :].ravel(), h[t]) for t in range(T)])
    d_x = torch.from_numpy(xs).cuda()
    d_y = torch.zeros(M, T, B, device="cuda")
    d_mix = torch.zeros(M, 2, B, device="cuda")
    stream = torch.cuda.Stream()
    with g.ConvEngine(T, B, L, algo) as e:
        e.load_ir(h)
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            for m in range(M):
                e.process(d_x[m].data_ptr(), d_y[m].data_ptr(), d_mix[m].data_ptr(), stream=stream.cuda_stream)
        stream.synchronize()
        q = e.query()
        assert q["kernel_launches"] >= M * q["kernels_per_block"]
    got = d_y.cpu().numpy().transpose(1, 0, 2).reshape(T, M * B)
    assert_parity(got, want, algo, "device-pointer stream")


@pytest.mark.parametrize("algo", [g.ALGO_DIRECT, g.ALGO_UPOLS])
@pytest.mark.parametrize("layout", [g.OUT_TRACK_MAJOR, g.OUT_SAMPLE_MAJOR])
def test_host_call_with_pinned_buffers_zero_copy(oracle, algo, layout):
    """b200conv_process_host on PINNED host buffers: the kernels read the input (and, for the direct
    engine, write output + bus) in place over PCIe; results must equal the staged-copy path bit for bit."""
    import torch
    T, B, L, M = 12, 512, 3000, 8
    xs = oracle.generate_input(M * T * B, 9).reshape(M, T, B)
    h = oracle.generate_ir(T, L, "accel")
    want = np.stack([oracle.stream(xs[:, t, :].ravel(), h[t]) for t in range(T)])
    shape = (B, T) if layout == g.OUT_SAMPLE_MAJOR else (T, B)
    h_in = torch.from_numpy(xs).pin_memory()
    h_out = torch.zeros((M,) + shape).pin_memory()
    h_mix = torch.zeros(M, 2, B).pin_memory()
    with g.ConvEngine(T, B, L, algo, layout) as e:
        e.load_ir(h)
        for m in range(M):
            e.process_host_ptr(h_in[m].data_ptr(), h_out[m].data_ptr(), h_mix[m].data_ptr())
        e.reset()
        staged = [e.process_host(xs[m], want_mix=True) for m in range(M)]  # pageable numpy -> staged copies
    got = h_out.numpy()
    for m in range(M):
        assert np.array_equal(got[m], staged[m][0]) and np.array_equal(h_mix[m].numpy(), staged[m][1])
    got_tm = np.concatenate([(got[m].T if layout == g.OUT_SAMPLE_MAJOR else got[m]) for m in range(M)], axis=1)
    assert_parity(got_tm, want, algo, "pinned zero-copy host call")


# ---------------------------------------------------------------------------------------------
# Full BASELINE sizes: size-independent properties + a track subset against the oracle
# ---------------------------------------------------------------------------------------------
def _fp64_truth(x_hist_and_block, h):
    from scipy.signal import fftconvolve
    return fftconvolve(x_hist_and_block.astype(np.float64), h.astype(np.float64))[:x_hist_and_block.size]


def test_c2_full_size_direct_vs_r1_subset_and_upols_and_fp64(oracle):
    """C2: 128 tracks x 512-sample buffer x 16384 taps.  R1 supplies real history through its
    cross-track bleed; check tracks {0, 31, 32, 77, 127} against the oracle's own loop, every track
    against fp64 truth, and the two independent engines against each other."""
    T, B, L = 128, 512, 16384
    x = oracle.generate_input(T * B)
    h = oracle.generate_ir(T, L, "direct")
    padded = np.concatenate([np.zeros(L - 1, dtype=np.float32), x])
    hist = np.stack([padded[t * B:t * B + L - 1] for t in range(T)])
    out = {}
    for algo in (g.ALGO_DIRECT, g.ALGO_UPOLS):
        with g.ConvEngine(T, B, L, algo) as e:
            e.load_ir(h)
            e.prime_history(hist)
            out[algo], _ = e.process_host(x.reshape(T, B))
    for t in (0, 31, 32, 77, 127):
        ref_t = oracle.stream(padded[t * B:t * B + L - 1 + B], h[t])[L - 1:]  # == R1 row t (test_oracle proves it)
        for algo in out:
            assert_parity(out[algo][t], ref_t, algo, f"C2 track {t} vs R1")
    truth = np.stack([_fp64_truth(padded[t * B:t * B + L - 1 + B], h[t])[L - 1:] for t in range(T)])
    assert snr_db(out[g.ALGO_DIRECT], truth) >= 100
    assert snr_db(out[g.ALGO_UPOLS], truth) >= 90
    assert snr_db(out[g.ALGO_UPOLS], out[g.ALGO_DIRECT]) >= 90


def test_c3_full_size_upols_streaming(oracle):
    """C3: 1024 tracks x 256-sample blocks x 65536 taps (P = 256).  Stream P+3 blocks so every
    partition and ring slot is used; tracks {0, 1023} against the streaming oracle (the
    reference loop), 16 tracks against fp64 truth; PEEK idempotence on the whole job."""
    T, B, L = 1024, 256, 65536
    P = L // B
    M = P + 3
    rng = np.random.default_rng(11)
    xs = rng.uniform(-1, 1, size=(M, T, B)).astype(np.float32)
    h = oracle.generate_ir(T, L, "accel")
    keep = [0, 1023]
    fp64_tracks = list(range(0, T, 64))
    with g.ConvEngine(T, B, L, g.ALGO_UPOLS) as e:
        e.load_ir(h)
        q = e.query()
        assert q["partitions"] == 256 and q["alg_bytes_per_block"] == T * 16 * P * (B + 1)
        outs = [e.process_host(xs[m])[0] for m in range(M)]
        # idempotence at full size: a PEEK of the next buffer equals the committed result
        y_peek, _ = e.process_host(xs[0], flags=g.PEEK)
        y_commit, _ = e.process_host(xs[0])
    got = np.concatenate(outs, axis=1)  # [T][M*B]
    for t in keep:
        want = oracle.stream(xs[:, t, :].ravel(), h[t])
        assert_parity(got[t], want, g.ALGO_UPOLS, f"C3 track {t} all blocks")
        assert_parity(got[t, -B:], want[-B:], g.ALGO_UPOLS, f"C3 track {t} last block")
    for t in fp64_tracks:
        truth = _fp64_truth(xs[:, t, :].ravel(), h[t])
        assert snr_db(got[t, -4 * B:], truth[-4 * B:]) >= 90
    assert np.array_equal(y_peek, y_commit)


def test_c4_shard_upols_partial_last_partition(oracle):
    """One GPU's share of C4: 96000 taps at B = 512 -> P = 188 with a half-full last partition.
    16 tracks with global indices [1000, 1016) of 4096; impulse at the very last tap."""
    Tg, t0, T, B, L = 4096, 1000, 16, 512, 96000
    P = (L + B - 1) // B
    M = P + 2
    h = oracle.generate_ir(Tg, L, "accel", t0, t0 + T)
    h[3, :] = 0
    h[3, L - 1] = 1.0  # pure delay of L-1 samples
    rng = np.random.default_rng(12)
    xs = rng.uniform(-1, 1, size=(M, T, B)).astype(np.float32)
    with g.ConvEngine(T, B, L, g.ALGO_UPOLS, track_offset=t0, total_tracks=Tg) as e:
        e.load_ir(h)
        assert e.query()["partitions"] == 188
        got = np.concatenate([e.process_host(xs[m])[0] for m in range(M)], axis=1)
    stream3 = xs[:, 3, :].ravel()
    want3 = np.concatenate([np.zeros(L - 1, dtype=np.float32), stream3[:M * B - (L - 1)]])
    assert np.abs(got[3] - want3).max() <= 1e-5
    for t in (0, 15):
        truth = _fp64_truth(xs[:, t, :].ravel(), h[t])
        assert snr_db(got[t, -2 * B:], truth[-2 * B:]) >= 90


# ---------------------------------------------------------------------------------------------
# Error behaviour of the ABI
# ---------------------------------------------------------------------------------------------
def test_error_codes():
    with pytest.raises(g.B200ConvError) as ei:
        g.ConvEngine(4, 48, 100, g.ALGO_DIRECT)
    assert ei.value.code == g.engine.ERR_INVALID
    with pytest.raises(g.B200ConvError) as ei:
        g.ConvEngine(4, 100, 100, g.ALGO_UPOLS)
    assert ei.value.code == g.engine.ERR_INVALID
    with g.ConvEngine(2, 64, 100, g.ALGO_UPOLS) as e:
        with pytest.raises(g.B200ConvError) as ei:
            e.process_host(np.zeros((2, 64), dtype=np.float32))
        assert ei.value.code == g.engine.ERR_STATE
    cfg = g.engine.make_config(1, 64, 10, g.ALGO_DIRECT)
    cfg.abi_version = 99
    import ctypes
    handle = ctypes.c_void_p()
    assert g.load_library().b200conv_create(ctypes.byref(cfg), ctypes.byref(handle)) == g.engine.ERR_ABI


def test_fp32_peak_microbenchmark_runs():
    tf, ms = g.measure_fp32_peak(0)
    assert 20.0 < tf < 120.0, tf  # B200: 148 SMs x 128 lanes x 2 x (1.3..1.97 GHz) = 49..75 TFLOP/s


@pytest.mark.parametrize("algo,layout", [(g.ALGO_DIRECT, g.OUT_TRACK_MAJOR), (g.ALGO_UPOLS, g.OUT_SAMPLE_MAJOR),
                                         (g.ALGO_UPOLS, g.OUT_TRACK_MAJOR), (g.ALGO_DIRECT_TC, g.OUT_TRACK_MAJOR)])
@pytest.mark.parametrize("pinned", [True, False])
def test_submit_wait_pipelines_two_blocks_and_matches_the_serial_call(oracle, algo, layout, pinned):
    """b200conv_submit / b200conv_wait: two blocks in flight on double-buffered staging give the
    bit-identical stream of outputs and buses as one b200conv_process_host per block."""
    import torch
    T, B, L, M = 12, 256, 1500, 9
    xs = oracle.generate_input(M * T * B, 5).reshape(M, T, B)
    h = oracle.generate_ir(T, L, "accel")
    shape = (B, T) if layout == g.OUT_SAMPLE_MAJOR else (T, B)
    with g.ConvEngine(T, B, L, algo, layout) as serial, g.ConvEngine(T, B, L, algo, layout) as piped:
        serial.load_ir(h)
        piped.load_ir(h)
        want = [serial.process_host(xs[m], want_mix=True) for m in range(M)]
        mk = (lambda *s: torch.zeros(*s).pin_memory()) if pinned else (lambda *s: torch.zeros(*s))
        h_in = [torch.from_numpy(xs[m].copy()) for m in range(M)]
        if pinned:
            h_in = [t.pin_memory() for t in h_in]
        outs = [mk(*shape) for _ in range(M)]
        mixes = [mk(2, B) for _ in range(M)]
        prev = None
        for m in range(M):
            tk = piped.submit_ptr(h_in[m].data_ptr(), outs[m].data_ptr(), mixes[m].data_ptr())
            if prev is not None:
                piped.wait(prev)
            prev = tk
        piped.wait(prev)
        with pytest.raises(g.B200ConvError):
            piped.wait(prev + 5)
    for m in range(M):
        assert np.array_equal(outs[m].numpy(), want[m][0]), f"block {m}"
        assert np.array_equal(mixes[m].numpy(), want[m][1]), f"bus of block {m}"
