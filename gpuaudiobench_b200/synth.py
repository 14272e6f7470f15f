"""Synthetic signals of the reference's shape for bench.py (harness; independent of oracle/).

Input: the reference draws one sequential std::mt19937(seed) + uniform_real_distribution<float>(-1,1)
stream over the flat [T][B] array (cuda/bench_utils.cu:238-245).  libstdc++'s generate_canonical
<float,24> consumes one 32-bit word per sample, so the same values come out of NumPy's MT19937
(SURVEY.md App. A.3): bit-identical, no C++ needed.
IR: Hamming-windowed sinc, cutoff 0.1 + 0.05 t/Tg, centre L/2, scaled 1/L (cuda/bench_conv1d.cu:
159-178), evaluated in float32 NumPy — same formula and constants; the last ulp of sinf/cosf may
differ from glibc's, which does not matter for a throughput workload (the parity tests use the
oracle's bit-exact generator instead).
"""
import numpy as np

HAMMING_A0, HAMMING_A1 = np.float32(0.54), np.float32(0.46)      # benchmark_constants: HAMMING_WINDOW_A0/A1
IR_BASE_FREQ, IR_FREQ_RANGE = np.float32(0.1), np.float32(0.05)  # CONV1D_IR_BASE_FREQ / CONV1D_IR_FREQ_RANGE


def make_input(count, seed=42, skip=0):
    """`count` samples of the reference input stream, starting `skip` samples in."""
    rs = np.random.RandomState(seed)
    if skip:
        rs.randint(0, 2 ** 32, skip, dtype=np.uint64)
    u = rs.randint(0, 2 ** 32, count, dtype=np.uint64)
    x = u.astype(np.float32) / np.float32(2 ** 32)
    x = np.minimum(x, np.nextafter(np.float32(1), np.float32(0)))
    return x * np.float32(2) + np.float32(-1)


def make_ir(total_tracks, ir_len, t_begin=0, t_end=None):
    """IRs [t_end - t_begin][L] of tracks [t_begin, t_end) of a total_tracks-track job."""
    t_end = total_tracks if t_end is None else t_end
    two_pi = np.float32(2.0) * np.float32(3.14159265358979323846)
    k = np.arange(ir_len, dtype=np.float32)
    tt = k - np.float32(ir_len) / np.float32(2.0)
    window = HAMMING_A0 - HAMMING_A1 * np.cos(two_pi * k / np.float32(ir_len - 1), dtype=np.float32)
    out = np.empty((t_end - t_begin, ir_len), dtype=np.float32)
    for i, t in enumerate(range(t_begin, t_end)):
        freq = IR_BASE_FREQ + IR_FREQ_RANGE * np.float32(t) / np.float32(total_tracks)
        arg = two_pi * freq * tt
        with np.errstate(invalid="ignore", divide="ignore"):
            sinc = np.where(tt == 0, np.float32(1.0), np.sin(arg, dtype=np.float32) / arg).astype(np.float32)
        out[i] = window * sinc / np.float32(ir_len)
    return out


def butterworth_lowpass(fc):
    """2nd-order Butterworth low-pass biquad (b0, b1, b2, a1, a2; a0 = 1) at fc cycles/sample — the RBJ
    cookbook formula the reference's IIR plugin uses (cuda/bench_iir.cu:205-228), evaluated in float32."""
    f = np.float32
    omega = f(2.0) * f(np.pi) * f(fc)
    c, s_ = f(np.cos(omega)), f(np.sin(omega))
    alpha = s_ / (f(2.0) * f(0.707))
    a0 = f(1.0) + alpha
    b1 = f(1.0) - c
    return np.array([b1 / f(2.0) / a0, b1 / a0, b1 / f(2.0) / a0, f(-2.0) * c / a0, (f(1.0) - alpha) / a0], dtype=np.float32)
