"""ctypes doorway to the CPU checker under oracle/ (test infrastructure, never the product path).

`Oracle` wraps oracle/liboracle.so (our restatement of the reference's CPU loops);
`RefLib` wraps oracle/_ref/libgpuab_ref.so (the reference's own compiled functions, when built).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libgpuab_ref.so")

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


def build_oracle(with_ref=True):
    """Compile the checker (g++) and, when /root/reference is present, oracle/_ref (nvcc)."""
    targets = ["all"] + (["ref"] if with_ref else [])
    subprocess.run(["make", "-s", "-C", ORACLE_DIR] + targets, check=True)


def fnv1a64(arr) -> str:
    """FNV-1a-64 over the raw little-endian bytes (the tripwire hash of SURVEY.md App. A.3)."""
    h = 1469598103934665603
    for b in np.ascontiguousarray(arr).tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build_oracle(with_ref=False)
        L = self.lib = C.CDLL(ORACLE_SO)
        L.oracle_generate_input.argtypes = [_f32p, C.c_size_t, C.c_uint]
        for name in ("oracle_generate_ir_direct", "oracle_generate_ir_accel"):
            getattr(L, name).argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.oracle_conv1d_r1.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int]
        L.oracle_conv1d_r2.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int]
        L.oracle_stream.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int64]
        L.oracle_fft_reference.argtypes = [_f32p, _f32p, _f32p, C.c_int]
        L.oracle_compare_abs.argtypes = [_f32p, _f32p, C.c_size_t, C.c_float,
                                         C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int)]
        L.oracle_compare_rel.argtypes = [_f32p, _f32p, C.c_size_t, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.oracle_statistics.argtypes = [_f32p, C.c_size_t, _f32p]
        L.oracle_nearest_rank.argtypes = [_f32p, C.c_size_t, C.c_int, C.c_int, _f32p]
        L.oracle_fnv1a64.argtypes = [C.c_void_p, C.c_size_t]
        L.oracle_fnv1a64.restype = C.c_uint64
        for name in ("oracle_time_r1", "oracle_time_r2"):
            getattr(L, name).argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int]
            getattr(L, name).restype = C.c_double
        L.oracle_hardware_threads.restype = C.c_int
        L.oracle_gain.argtypes = [_f32p, _f32p, C.c_size_t, C.c_float]
        L.oracle_gainstats.argtypes = [_f32p, _f32p, _f32p, C.c_size_t, C.c_size_t, C.c_float]
        L.oracle_butterworth.argtypes = [C.c_float, _f32p]
        L.oracle_iir.argtypes = [_f32p, _f32p, _f32p, C.c_int, _f32p, C.c_int, C.c_int]
        L.oracle_strip.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_uint, C.c_float, C.c_void_p, C.c_void_p, C.c_int,
                                   _f32p, _f32p]

    # -- generators --------------------------------------------------------------------------
    def generate_input(self, count, seed=42):
        buf = np.empty(count, dtype=np.float32)
        self.lib.oracle_generate_input(buf, count, seed)
        return buf

    def generate_ir(self, T, L, variant="direct", t_begin=0, t_end=None):
        """IRs [t_end-t_begin][L]; the cutoff uses the global track index t and total T."""
        t_end = T if t_end is None else t_end
        h = np.empty((t_end - t_begin, L), dtype=np.float32)
        fn = self.lib.oracle_generate_ir_direct if variant == "direct" else self.lib.oracle_generate_ir_accel
        fn(h, t_begin, t_end, T, L)
        return h

    # -- oracles -----------------------------------------------------------------------------
    def r1(self, x, h, L, B, T):
        """Track-major [T][B]; flat-index history bleed (bench_conv1d.cu:188-208)."""
        y = np.empty(T * B, dtype=np.float32)
        self.lib.oracle_conv1d_r1(np.ascontiguousarray(x.ravel()), np.ascontiguousarray(h.ravel()), y, L, B, T)
        return y.reshape(T, B)

    def r2(self, x, h, L, B, T):
        """Sample-major [B][T]; per-track zero history (bench_conv1d_accel.cu:234-252)."""
        y = np.empty(T * B, dtype=np.float32)
        self.lib.oracle_conv1d_r2(np.ascontiguousarray(x.ravel()), np.ascontiguousarray(h.ravel()), y, L, B, T)
        return y.reshape(B, T)

    def stream(self, x, h):
        """Streaming oracle for one track: y[n] = sum_k h[k] x[n-k], n < len(x) (App. A.2)."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        h = np.ascontiguousarray(h, dtype=np.float32)
        y = np.empty_like(x)
        self.lib.oracle_stream(x, h, y, h.size, x.size)
        return y

    def fft_reference(self, x):
        """Naive float DFT of one row (bench_fft.cu:149-168): complex64 [n/2+1]."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        re = np.empty(x.size // 2 + 1, dtype=np.float32)
        im = np.empty_like(re)
        self.lib.oracle_fft_reference(x, re, im, x.size)
        return re + 1j * im

    # -- metrics -----------------------------------------------------------------------------
    def compare_abs(self, gpu, cpu, tol):
        mx, mean, over = C.c_float(), C.c_float(), C.c_int()
        g = np.ascontiguousarray(gpu.ravel(), dtype=np.float32)
        c = np.ascontiguousarray(cpu.ravel(), dtype=np.float32)
        self.lib.oracle_compare_abs(g, c, g.size, tol, C.byref(mx), C.byref(mean), C.byref(over))
        return mx.value, mean.value, over.value

    def compare_rel(self, gpu, cpu):
        mx, mean = C.c_float(), C.c_float()
        g = np.ascontiguousarray(gpu.ravel(), dtype=np.float32)
        c = np.ascontiguousarray(cpu.ravel(), dtype=np.float32)
        self.lib.oracle_compare_rel(g, c, g.size, C.byref(mx), C.byref(mean))
        return mx.value, mean.value

    def statistics(self, lat):
        lat = np.ascontiguousarray(lat, dtype=np.float32)
        out = np.zeros(8, dtype=np.float32)
        self.lib.oracle_statistics(lat, lat.size, out)
        return dict(zip(("mean", "median", "std", "min", "max", "p95", "p99", "count"), out.tolist()))

    def nearest_rank(self, lat, bufsize, fs):
        lat = np.ascontiguousarray(lat, dtype=np.float32)
        out = np.zeros(5, dtype=np.float32)
        self.lib.oracle_nearest_rank(lat, lat.size, bufsize, fs, out)
        return dict(zip(("p50", "p95", "p99", "threshold_ms", "meets_deadline"), out.tolist()))

    # -- timing legs (bench.py) --------------------------------------------------------------
    def time_r1(self, x, h, L, B, T, nthreads):
        y = np.empty(T * B, dtype=np.float32)
        return self.lib.oracle_time_r1(np.ascontiguousarray(x.ravel()), np.ascontiguousarray(h.ravel()), y, L, B, T, nthreads)

    def time_r2(self, x, h, L, B, T, nthreads):
        y = np.empty(T * B, dtype=np.float32)
        return self.lib.oracle_time_r2(np.ascontiguousarray(x.ravel()), np.ascontiguousarray(h.ravel()), y, L, B, T, nthreads)

    def hardware_threads(self):
        return self.lib.oracle_hardware_threads()

    # -- channel strip: track-major [T][B] arrays ------------------------------
    def gain(self, x, gain):
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.empty_like(x)
        self.lib.oracle_gain(x.ravel(), y.reshape(-1), x.size, gain)
        return y

    def gainstats(self, x, gain):
        x = np.ascontiguousarray(x, dtype=np.float32)
        T, B = x.shape
        y = np.empty_like(x)
        stats = np.empty((T, 2), dtype=np.float32)
        self.lib.oracle_gainstats(x.ravel(), y.reshape(-1), stats.reshape(-1), T, B, gain)
        return y, stats

    def butterworth(self, fc=0.25):
        out = np.empty(5, dtype=np.float32)
        self.lib.oracle_butterworth(fc, out)
        return out

    def iir(self, x, coeffs, state):
        """coeffs [5] (shared) or [T][5]; state [T][2] is updated in place."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        T, B = x.shape
        coeffs = np.ascontiguousarray(coeffs, dtype=np.float32)
        y = np.empty_like(x)
        self.lib.oracle_iir(x.ravel(), y.reshape(-1), coeffs.reshape(-1), 0 if coeffs.ndim == 1 else 5,
                            state.reshape(-1), T, B)
        return y

    def strip(self, x, ops, gain=1.0, gains=None, coeffs=None, state=None):
        """Engine strip order: stats(input) -> gain -> biquad.  Returns (y, stats); state updated in place."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        T, B = x.shape
        y = np.empty_like(x)
        stats = np.zeros((T, 2), dtype=np.float32)
        if state is None:
            state = np.zeros((T, 2), dtype=np.float32)
        g = None if gains is None else np.ascontiguousarray(gains, dtype=np.float32)
        c = None if coeffs is None else np.ascontiguousarray(coeffs, dtype=np.float32)
        self.lib.oracle_strip(x.ravel(), y.reshape(-1), T, B, ops, gain,
                              None if g is None else g.ctypes.data, None if c is None else c.ctypes.data,
                              0 if (c is None or c.ndim == 1) else 5, state.reshape(-1), stats.reshape(-1))
        return y, stats


class RefLib:
    """The reference's own compiled CPU functions (oracle/_ref). `available()` is False when the
    library has not been built (no /root/reference at build time and no prebuilt copy)."""

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def __init__(self):
        L = self.lib = C.CDLL(REF_SO)
        L.ref_generate_input.argtypes = [_f32p, C.c_size_t, C.c_uint]
        L.ref_conv1d_r1.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int]
        L.ref_conv1d_r2.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int]
        L.ref_generate_ir_direct.argtypes = [_f32p, C.c_int, C.c_int]
        L.ref_generate_ir_accel.argtypes = [_f32p, C.c_int, C.c_int]
        L.ref_statistics.argtypes = [_f32p, C.c_size_t, _f32p]
        L.ref_fft_reference.argtypes = [_f32p, _f32p, _f32p, C.c_int]
        L.ref_json_results.argtypes = [_f32p, C.c_size_t, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_size_t]
        for name in ("ref_time_r1", "ref_time_r2"):
            getattr(L, name).argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int]
            getattr(L, name).restype = C.c_double
        L.ref_hardware_threads.restype = C.c_int
        if hasattr(L, "ref_gain_reference"):
            L.ref_gain_reference.argtypes = [_f32p, _f32p, C.c_size_t, C.c_size_t]
            L.ref_gain_reference.restype = C.c_float
            L.ref_gainstats_reference.argtypes = [_f32p, _f32p, _f32p, C.c_size_t, C.c_size_t]
            L.ref_gainstats_reference.restype = C.c_float
            L.ref_butterworth.argtypes = [C.c_float, _f32p]
            L.ref_iir_reference.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int]

    def gain(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        T, B = x.shape
        y = np.empty_like(x)
        g = self.lib.ref_gain_reference(x.ravel(), y.reshape(-1), B, T)
        return y, g

    def gainstats(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        T, B = x.shape
        y = np.empty_like(x)
        stats = np.empty((T, 2), dtype=np.float32)
        g = self.lib.ref_gainstats_reference(x.ravel(), y.reshape(-1), stats.reshape(-1), B, T)
        return y, stats, g

    def butterworth(self, fc=0.25):
        out = np.empty(5, dtype=np.float32)
        self.lib.ref_butterworth(fc, out)
        return out

    def iir(self, x, coeffs5, state):
        x = np.ascontiguousarray(x, dtype=np.float32)
        T, B = x.shape
        y = np.empty_like(x)
        self.lib.ref_iir_reference(x.ravel(), y.reshape(-1), np.ascontiguousarray(coeffs5, dtype=np.float32),
                                   state.reshape(-1), T, B)
        return y

    def generate_input(self, count, seed=42):
        buf = np.empty(count, dtype=np.float32)
        self.lib.ref_generate_input(buf, count, seed)
        return buf

    def generate_ir(self, T, L, variant="direct"):
        h = np.empty((T, L), dtype=np.float32)
        (self.lib.ref_generate_ir_direct if variant == "direct" else self.lib.ref_generate_ir_accel)(h, T, L)
        return h

    def r1(self, x, h, L, B, T):
        y = np.empty(T * B, dtype=np.float32)
        self.lib.ref_conv1d_r1(np.ascontiguousarray(x.ravel()), np.ascontiguousarray(h.ravel()), y, L, B, T)
        return y.reshape(T, B)

    def r2(self, x, h, L, B, T):
        y = np.empty(T * B, dtype=np.float32)
        self.lib.ref_conv1d_r2(np.ascontiguousarray(x.ravel()), np.ascontiguousarray(h.ravel()), y, L, B, T)
        return y.reshape(B, T)

    def fft_reference(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        re = np.empty(x.size // 2 + 1, dtype=np.float32)
        im = np.empty_like(re)
        self.lib.ref_fft_reference(x, re, im, x.size)
        return re + 1j * im

    def statistics(self, lat):
        lat = np.ascontiguousarray(lat, dtype=np.float32)
        out = np.zeros(8, dtype=np.float32)
        self.lib.ref_statistics(lat, lat.size, out)
        return dict(zip(("mean", "median", "std", "min", "max", "p95", "p99", "count"), out.tolist()))

    def json_results(self, lat, name, fs, bufsize, ntracks):
        lat = np.ascontiguousarray(lat, dtype=np.float32)
        buf = C.create_string_buffer(4096)
        n = self.lib.ref_json_results(lat, lat.size, name.encode(), fs, bufsize, ntracks, buf, 4096)
        assert n >= 0
        return buf.value.decode()

    def time_r1(self, x, h, L, B, T, nthreads):
        y = np.empty(T * B, dtype=np.float32)
        return self.lib.ref_time_r1(np.ascontiguousarray(x.ravel()), np.ascontiguousarray(h.ravel()), y, L, B, T, nthreads)

    def time_r2(self, x, h, L, B, T, nthreads):
        y = np.empty(T * B, dtype=np.float32)
        return self.lib.ref_time_r2(np.ascontiguousarray(x.ravel()), np.ascontiguousarray(h.ravel()), y, L, B, T, nthreads)

    def hardware_threads(self):
        return self.lib.ref_hardware_threads()
