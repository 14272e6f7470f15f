"""GPU parity tests (-m gpu) of the channel strip: per-track statistics, gain and
Direct-Form-II biquad on the engine's output stage, called through the C ABI.

Bar: BIT-EXACT.  The strip kernels evaluate the reference's CPU loops (cuda/bench_gain.cu:90-92,
bench_gainstats.cu:121-142, bench_iir.cu:176-203) in the same order with un-contracted fp32 operations,
so outputs, filter state and statistics must equal the oracle's bit for bit on the same input.
"""
import numpy as np
import pytest
import torch

import gpuaudiobench_b200 as g

pytestmark = pytest.mark.gpu

OPS = [g.STRIP_STATS, g.STRIP_GAIN, g.STRIP_STATS | g.STRIP_GAIN, g.STRIP_BIQUAD, g.STRIP_STATS | g.STRIP_BIQUAD,
       g.STRIP_GAIN | g.STRIP_BIQUAD, g.STRIP_STATS | g.STRIP_GAIN | g.STRIP_BIQUAD]


def per_track_biquads(oracle, T):
    """Butterworth low-passes from fc = 0.05 to 0.45 (all stable), one per track."""
    return np.stack([oracle.butterworth(0.05 + 0.4 * t / max(1, T - 1)) for t in range(T)])


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("ops", OPS)
@pytest.mark.parametrize("T,B", [(1, 512), (128, 512), (7, 100), (33, 31), (5, 1), (16, 4096)])
def test_strip_track_major_bit_exact_two_blocks(oracle, ops, T, B):
    rng = np.random.default_rng(T * 1000 + B)
    coef = per_track_biquads(oracle, T)
    gains = rng.uniform(0.25, 2.0, T).astype(np.float32)
    st_ref = np.zeros((T, 2), np.float32)
    d_state = torch.zeros(T, 2, device="cuda")
    d_stats = torch.zeros(T, 2, device="cuda")
    d_coef, d_gains = dev(coef), dev(gains)
    for blk in range(2):  # the second block starts from the first block's filter state
        x = oracle.generate_input(T * B, seed=7 + blk).reshape(T, B)
        ref, stats_ref = oracle.strip(x, ops, gains=gains, coeffs=coef, state=st_ref)
        d_x = dev(x)
        d_y = torch.empty_like(d_x)
        g.strip_process(d_x.data_ptr(), d_y.data_ptr(), T, B, ops, d_gains=d_gains.data_ptr(), d_biquad=d_coef.data_ptr(),
                        shared_coeffs=False, d_state=d_state.data_ptr(), d_stats=d_stats.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(d_y.cpu().numpy(), ref), f"block {blk}"
        if ops & g.STRIP_BIQUAD:
            assert np.array_equal(d_state.cpu().numpy(), st_ref)
        if ops & g.STRIP_STATS:
            assert np.array_equal(d_stats.cpu().numpy(), stats_ref)


def test_strip_matches_each_reference_plugin(oracle):
    """The three plugin configurations exactly as the reference runs them: Gain x2.0, GainStats x0.5 with
    mean/max of the input, IIRFilter with ONE Butterworth fc = 0.25 set for all tracks."""
    T, B = 128, 512
    x = oracle.generate_input(T * B).reshape(T, B)
    d_x = dev(x)
    d_y = torch.empty_like(d_x)
    g.strip_process(d_x.data_ptr(), d_y.data_ptr(), T, B, g.STRIP_GAIN, gain=2.0)
    assert np.array_equal(d_y.cpu().numpy(), oracle.gain(x, 2.0))
    d_stats = torch.zeros(T, 2, device="cuda")
    g.strip_process(d_x.data_ptr(), d_y.data_ptr(), T, B, g.STRIP_GAIN | g.STRIP_STATS, gain=0.5, d_stats=d_stats.data_ptr())
    y_ref, s_ref = oracle.gainstats(x, 0.5)
    assert np.array_equal(d_y.cpu().numpy(), y_ref) and np.array_equal(d_stats.cpu().numpy(), s_ref)
    coef = oracle.butterworth(0.25)
    st = np.zeros((T, 2), np.float32)
    d_state = torch.zeros(T, 2, device="cuda")
    d_coef = dev(coef)
    for _ in range(3):  # the reference keeps the filter state across iterations (bench_iir.cu:42-43)
        ref = oracle.iir(x, coef, st)
        g.strip_process(d_x.data_ptr(), d_y.data_ptr(), T, B, g.STRIP_BIQUAD, d_biquad=d_coef.data_ptr(),
                        d_state=d_state.data_ptr())
        assert np.array_equal(d_y.cpu().numpy(), ref)
        assert np.array_equal(d_state.cpu().numpy(), st)


def test_strip_in_place_and_peek(oracle):
    T, B = 9, 256
    x = oracle.generate_input(T * B).reshape(T, B)
    coef = oracle.butterworth(0.2)
    st0 = np.random.default_rng(3).uniform(-0.5, 0.5, (T, 2)).astype(np.float32)
    d_state, d_coef = dev(st0), dev(coef)
    d_x = dev(x)
    g.strip_process(d_x.data_ptr(), d_x.data_ptr(), T, B, g.STRIP_BIQUAD, d_biquad=d_coef.data_ptr(),
                    d_state=d_state.data_ptr(), flags=g.PEEK)
    st = st0.copy()
    ref = oracle.iir(x, coef, st)
    assert np.array_equal(d_x.cpu().numpy(), ref)                 # in place
    assert np.array_equal(d_state.cpu().numpy(), st0)             # PEEK: state untouched


@pytest.mark.parametrize("T,B,ld,col0", [(64, 256, 64, 0), (40, 96, 128, 50), (3, 33, 7, 2), (130, 512, 130, 0)])
def test_strip_sample_major_column_tile(oracle, T, B, ld, col0):
    """[B][ld] with the tracks in columns col0..col0+T: same bits as the track-major oracle, other
    columns untouched."""
    ops = g.STRIP_STATS | g.STRIP_GAIN | g.STRIP_BIQUAD
    x = oracle.generate_input(T * B, seed=11).reshape(T, B)
    coef = per_track_biquads(oracle, T)
    st = np.zeros((T, 2), np.float32)
    ref, stats_ref = oracle.strip(x, ops, gain=0.75, coeffs=coef, state=st)
    full = np.full((B, ld), 123.0, np.float32)
    full[:, col0:col0 + T] = x.T
    d_full = dev(full)
    d_state = torch.zeros(T, 2, device="cuda")
    d_stats = torch.zeros(T, 2, device="cuda")
    d_coef = dev(coef)
    g.strip_process(d_full.data_ptr(), d_full.data_ptr(), T, B, ops, gain=0.75, d_biquad=d_coef.data_ptr(),
                    shared_coeffs=False, d_state=d_state.data_ptr(), d_stats=d_stats.data_ptr(),
                    layout=g.OUT_SAMPLE_MAJOR, ld=ld, col0=col0)
    out = d_full.cpu().numpy()
    assert np.array_equal(out[:, col0:col0 + T].T, ref)
    mask = np.ones(ld, bool)
    mask[col0:col0 + T] = False
    assert np.all(out[:, mask] == 123.0)
    assert np.array_equal(d_state.cpu().numpy(), st) and np.array_equal(d_stats.cpu().numpy(), stats_ref)


@pytest.mark.parametrize("algo,layout", [(g.ALGO_DIRECT, g.OUT_TRACK_MAJOR), (g.ALGO_DIRECT, g.OUT_SAMPLE_MAJOR),
                                         (g.ALGO_UPOLS, g.OUT_TRACK_MAJOR), (g.ALGO_UPOLS, g.OUT_SAMPLE_MAJOR)])
def test_engine_strip_equals_strip_of_engine_output(oracle, algo, layout):
    """convolution -> strip -> bus inside b200conv_process: the stripped output must be, bit for bit, the
    oracle's strip applied to what the same engine produces without a strip, over a 4-block stream
    (filter state carried), and the bus must be the gain-weighted sum of the STRIPPED tracks."""
    T, B, L, M = 24, 128, 700, 4
    ops = g.STRIP_STATS | g.STRIP_GAIN | g.STRIP_BIQUAD
    xs = oracle.generate_input(M * T * B, seed=5).reshape(M, T, B)
    h = oracle.generate_ir(T, L, "accel")
    coef = per_track_biquads(oracle, T)
    gains = np.linspace(0.5, 1.5, T).astype(np.float32)
    with g.ConvEngine(T, B, L, algo, layout) as plain, g.ConvEngine(T, B, L, algo, layout) as strip:
        plain.load_ir(h)
        strip.load_ir(h)
        strip.set_strip(ops, gains=gains, biquad=coef)
        st = np.zeros((T, 2), np.float32)
        mixg = None
        for m in range(M):
            y0, _ = plain.process_host(xs[m])
            y1, mix = strip.process_host(xs[m], want_mix=True)
            tm = y0.T if layout == g.OUT_SAMPLE_MAJOR else y0
            ref, stats_ref = oracle.strip(np.ascontiguousarray(tm), ops, gains=gains, coeffs=coef, state=st)
            got = y1.T if layout == g.OUT_SAMPLE_MAJOR else y1
            assert np.array_equal(got, ref), f"block {m}"
            assert np.array_equal(strip.strip_stats(), stats_ref)
            assert np.array_equal(strip.strip_state(), st)
            if mixg is None:
                from gpuaudiobench_b200.distributed import default_mix_gains
                mixg = default_mix_gains(T, 0, T).numpy().astype(np.float64)
            bus = np.stack([(mixg[:, c:c + 1] * ref.astype(np.float64)).sum(0) for c in (0, 1)])
            assert np.abs(mix - bus).max() <= 1e-5 * max(1e-30, np.abs(bus).max())
        # PEEK recomputes the next block without touching the conv history OR the filter state
        a, _ = strip.process_host(xs[0], flags=g.PEEK)
        b, _ = strip.process_host(xs[0], flags=g.PEEK)
        assert np.array_equal(a, b) and np.array_equal(strip.strip_state(), st)
        # removing the strip restores the plain engine; reset zeroes the filter state
        strip.set_strip(0)
        strip.reset()
        plain.reset()
        y0, _ = plain.process_host(xs[0])
        y1, _ = strip.process_host(xs[0])
        assert np.array_equal(y0, y1)


@pytest.mark.parametrize("layout", [g.OUT_TRACK_MAJOR, g.OUT_SAMPLE_MAJOR])
def test_engine_strip_on_the_three_kernel_upols_path(oracle, layout):
    """B >= 1024 runs rfft / FDL-MAC / irfft as separate kernels and the strip as its own launch before the
    bus (the fused kernel's in-epilogue strip covers B <= 512): same bit-exact contract."""
    T, B, L, M = 6, 1024, 3000, 3
    ops = g.STRIP_STATS | g.STRIP_GAIN | g.STRIP_BIQUAD
    xs = oracle.generate_input(M * T * B, seed=23).reshape(M, T, B)
    h = oracle.generate_ir(T, L, "accel")
    coef = per_track_biquads(oracle, T)
    with g.ConvEngine(T, B, L, g.ALGO_UPOLS, layout) as plain, g.ConvEngine(T, B, L, g.ALGO_UPOLS, layout) as strip:
        plain.load_ir(h)
        strip.load_ir(h)
        strip.set_strip(ops, gain=0.25, biquad=coef)
        st = np.zeros((T, 2), np.float32)
        for m in range(M):
            y0, _ = plain.process_host(xs[m])
            y1, _ = strip.process_host(xs[m])
            tm = np.ascontiguousarray(y0.T if layout == g.OUT_SAMPLE_MAJOR else y0)
            ref, stats_ref = oracle.strip(tm, ops, gain=0.25, coeffs=coef, state=st)
            assert np.array_equal(y1.T if layout == g.OUT_SAMPLE_MAJOR else y1, ref), f"block {m}"
            assert np.array_equal(strip.strip_stats(), stats_ref) and np.array_equal(strip.strip_state(), st)


def test_strip_error_codes():
    d = torch.zeros(4, 32, device="cuda")
    with pytest.raises(g.B200ConvError) as ei:
        g.strip_process(d.data_ptr(), d.data_ptr(), 4, 32, 0)
    assert ei.value.code == g.engine.ERR_INVALID
    with pytest.raises(g.B200ConvError) as ei:  # biquad without coefficients / state
        g.strip_process(d.data_ptr(), d.data_ptr(), 4, 32, g.STRIP_BIQUAD)
    assert ei.value.code == g.engine.ERR_INVALID
    with pytest.raises(g.B200ConvError) as ei:  # column tile outside the matrix
        g.strip_process(d.data_ptr(), d.data_ptr(), 4, 32, g.STRIP_GAIN, layout=g.OUT_SAMPLE_MAJOR, ld=4, col0=2)
    assert ei.value.code == g.engine.ERR_INVALID
