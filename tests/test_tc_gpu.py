"""The tensor-core direct-form engine (ALGO_DIRECT_TC, csrc/tc_toeplitz.cu) against the CPU oracle (-m gpu).

Same sum as the FFMA direct engine, computed by tcgen05 kind::tf32 MMAs with a 3-term hi/lo split, so the
tolerance is the direct engine's: SNR >= 100 dB and max|err| <= 1e-5 * max|y_ref| against the fp32 oracle
(the reference's own loop).  A delta IR does NOT reproduce the input bit for bit here (x = hi + lo loses 2^-22),
so index placement is checked with impulses to 1e-6 relative instead of exactly.
"""
import numpy as np
import pytest

import gpuaudiobench_b200 as g

pytestmark = pytest.mark.gpu
TC = g.ALGO_DIRECT_TC


def snr_db(got, ref):
    ref64 = np.asarray(ref, dtype=np.float64)
    err = np.sum((np.asarray(got, dtype=np.float64) - ref64) ** 2)
    return 10 * np.log10(np.sum(ref64 ** 2) / max(err, 1e-300))


def assert_parity(got, ref, what=""):
    s = snr_db(got, ref)
    mx = np.abs(np.asarray(got, dtype=np.float64) - ref).max()
    scale = np.abs(ref).max()
    assert s >= 100.0 and mx <= 1e-5 * scale, f"tc {what}: SNR {s:.1f} dB, max|err| {mx:.3e} vs scale {scale:.3e}"
    return s


@pytest.mark.parametrize("T,B,L,M", [
    (3, 128, 300, 6),       # one row block, L not a multiple of 128
    (4, 256, 1000, 7),
    (3, 512, 5000, 14),
    (2, 1024, 3000, 5),     # 8 row blocks
    (2, 128, 1, 3),         # single tap
    (5, 128, 128, 4),
    (2, 512, 20000, 42),    # two column groups (C - 1 > 128), ring wraps
    (150, 256, 700, 5),     # more tracks than SMs: the persistent track loop
    (3, 640, 4000, 6),      # bus chunks of 320 columns
    (3, 2048, 3000, 4),     # streamed through the kernel as two sub-blocks of 1024
    (2, 1536, 9000, 5),     # three sub-blocks of 512
])
def test_streaming_blocks_match_oracle(oracle, T, B, L, M):
    xs = oracle.generate_input(M * T * B, 7).reshape(M, T, B)
    if L == 1:
        h = np.array([[0.5]] * T, dtype=np.float32) * np.arange(1, T + 1, dtype=np.float32)[:, None]
    else:
        h = oracle.generate_ir(T, L, "direct")
    tracks = sorted({0, T // 2, T - 1})
    want = {t: oracle.stream(xs[:, t, :].ravel(), h[t]) for t in tracks}
    with g.ConvEngine(T, B, L, TC) as e:
        e.load_ir(h)
        got = np.concatenate([e.process_host(xs[m])[0] for m in range(M)], axis=1)
        nsub = 1 if B <= 1024 else (B // 1024 if B % 1024 == 0 else B // 512)
        assert e.query()["blocks_processed"] == M and e.query()["kernels_per_block"] == nsub
    for t in tracks:
        assert_parity(got[t], want[t], f"stream T={T} B={B} L={L} track {t}")
        assert_parity(got[t, -B:], want[t][-B:], "last block")


def test_c1_and_r1_with_primed_history(oracle, golden):
    T, B, L = 1, 512, 1024
    x = oracle.generate_input(T * B).reshape(T, B)
    h = oracle.generate_ir(T, L, "direct")
    with g.ConvEngine(T, B, L, TC) as e:
        e.load_ir(h)
        y, _ = e.process_host(x)
    assert_parity(y, golden["c1_r1"].reshape(T, B), "C1 vs the golden R1 output")
    T, B, L = 16, 512, 4096  # R1's cross-track bleed = primed history
    x = oracle.generate_input(T * B)
    h = oracle.generate_ir(T, L, "direct")
    padded = np.concatenate([np.zeros(L - 1, dtype=np.float32), x])
    hist = np.stack([padded[t * B:t * B + L - 1] for t in range(T)])
    with g.ConvEngine(T, B, L, TC) as e:
        e.load_ir(h)
        e.prime_history(hist)
        y, _ = e.process_host(x.reshape(T, B), flags=g.PEEK)
        y2, _ = e.process_host(x.reshape(T, B), flags=g.PEEK)
    assert_parity(y, oracle.r1(x, h, L, B, T), "primed block vs R1")
    assert np.array_equal(y, y2), "PEEK must not advance the stream state"


@pytest.mark.parametrize("layout", [g.OUT_TRACK_MAJOR, g.OUT_SAMPLE_MAJOR])
def test_impulse_placement_layouts_and_bus(oracle, layout):
    """delta IRs at known taps (first, last, across a 128-column seam) delay a ramp by exactly that many samples;
    tracks land in their own rows / columns; the bus is the gain-weighted sum."""
    T, B, L, M = 6, 256, 777, 6
    taps = [0, 1, 127, 128, 500, 776]
    h = np.zeros((T, L), dtype=np.float32)
    for t, k in enumerate(taps):
        h[t, k] = 1.0
    xs = (np.arange(M * B, dtype=np.float32)[None, :] * 1e-3 + np.arange(T, dtype=np.float32)[:, None] + 1.0)
    xs = np.ascontiguousarray(xs.reshape(T, M, B).transpose(1, 0, 2))
    with g.ConvEngine(T, B, L, TC, layout) as e:
        e.load_ir(h)
        outs = [e.process_host(xs[m], want_mix=True) for m in range(M)]
    got = np.concatenate([(y.T if layout == g.OUT_SAMPLE_MAJOR else y) for y, _ in outs], axis=1)
    for t, k in enumerate(taps):
        stream = xs[:, t, :].ravel()
        want = np.concatenate([np.zeros(k, dtype=np.float32), stream[:M * B - k]])
        assert np.abs(got[t] - want).max() <= 1e-6 * np.abs(want).max(), (t, k)
    theta = (np.arange(T) + 0.5) / T * np.pi / 2
    gains = np.stack([np.cos(theta), np.sin(theta)]) / np.sqrt(T)
    bus = np.concatenate([m for _, m in outs], axis=1)
    assert snr_db(bus, gains @ got.astype(np.float64)) >= 120


def test_c2_full_size_vs_r1_fp64_and_the_ffma_engine(oracle):
    """C2: 128 tracks x 512 x 16384 taps, the shape the tensor-core variant is dispatched for."""
    from scipy.signal import fftconvolve
    T, B, L = 128, 512, 16384
    M = L // B + 2
    rng = np.random.default_rng(21)
    xs = rng.uniform(-1, 1, size=(M, T, B)).astype(np.float32)
    h = oracle.generate_ir(T, L, "direct")
    out = {}
    for algo in (TC, g.ALGO_DIRECT):
        with g.ConvEngine(T, B, L, algo) as e:
            e.load_ir(h)
            ys = [e.process_host(xs[m], want_mix=True) for m in range(M)]
        out[algo] = (np.concatenate([y for y, _ in ys], axis=1), ys[-1][1])
    got, bus = out[TC]
    for t in (0, 77, 127):
        assert_parity(got[t], oracle.stream(xs[:, t, :].ravel(), h[t]), f"C2 track {t}")
    truth = np.stack([fftconvolve(xs[:, t, :].ravel().astype(np.float64), h[t].astype(np.float64))[:M * B][-B:] for t in range(T)])
    s_tc, s_ffma = snr_db(got[:, -B:], truth), snr_db(out[g.ALGO_DIRECT][0][:, -B:], truth)
    print(f"C2 last block vs fp64: tensor-core engine {s_tc:.1f} dB, FFMA engine {s_ffma:.1f} dB")
    assert s_tc >= 100
    theta = (np.arange(T) + 0.5) / T * np.pi / 2
    gains = np.stack([np.cos(theta), np.sin(theta)]) / np.sqrt(T)
    assert snr_db(bus, gains @ truth) >= 100


@pytest.mark.parametrize("B", [640, 2048])
def test_bus_and_peek_with_chunks_and_sub_blocks(oracle, B):
    """The bus of a block that is reduced in several chunks (B = 640: two of 320 columns) or streamed as sub-blocks
    (B = 2048: two launches), and PEEK across sub-blocks (state saved and put back): same outputs twice, then commit."""
    T, L, M = 6, 2500, 3
    xs = oracle.generate_input(M * T * B, 11).reshape(M, T, B)
    h = oracle.generate_ir(T, L, "direct")
    gains = (np.arange(2 * T, dtype=np.float32).reshape(T, 2) + 1) / (2 * T)
    with g.ConvEngine(T, B, L, TC) as e:
        e.load_ir(h)
        e.set_mix_gains(gains)
        for m in range(M):
            y0, mix0 = e.process_host(xs[m], flags=g.PEEK, want_mix=True)
            y1, mix1 = e.process_host(xs[m], flags=g.PEEK, want_mix=True)
            y2, mix2 = e.process_host(xs[m], want_mix=True)
            assert np.array_equal(y0, y1) and np.array_equal(y0, y2) and np.array_equal(mix0, mix2) and np.array_equal(mix0, mix1)
            want_mix = gains.T.astype(np.float64) @ y2.astype(np.float64)
            assert snr_db(mix2, want_mix) >= 120.0
        for t in (0, T - 1):
            want = oracle.stream(xs[:, t, :].ravel(), h[t])
            assert_parity(y2[t], want[-B:], f"B={B} last block track {t}")
