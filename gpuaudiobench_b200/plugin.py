"""ctypes binding of include/gpubench_plugin.h (libgpubench_b200.so): the reference-shaped plugin
lifecycle (setupBenchmark -> runBenchmark -> validate) as the gpubench CLI drives it.  Harness code
for the tests and bench.py."""
import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "lib", "libgpubench_b200.so")
GPUBENCH = os.path.join(PKG, "bin", "gpubench")


class Validation(C.Structure):
    _fields_ = [("status", C.c_int), ("max_error", C.c_float), ("mean_error", C.c_float), ("snr_db", C.c_double),
                ("max_abs_err", C.c_double), ("ref_peak", C.c_double)]


_lib = None


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} not built; run `python -m gpuaudiobench_b200.build`")
    L = C.CDLL(LIB_PATH)
    L.gpubench_last_error.restype = C.c_char_p
    L.gpubench_set_globals.argtypes = [C.c_int, C.c_int, C.c_int]
    L.gpubench_set_globals.restype = None
    L.gpubench_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int]
    L.gpubench_create.restype = C.c_void_p
    L.gpubench_destroy.argtypes = [C.c_void_p]
    L.gpubench_destroy.restype = None
    for name in ("gpubench_setup", "gpubench_iterate"):
        getattr(L, name).argtypes = [C.c_void_p]
    L.gpubench_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.gpubench_validate.argtypes = [C.c_void_p, C.POINTER(Validation), C.c_char_p, C.c_size_t]
    for name in ("gpubench_host_input", "gpubench_host_ir", "gpubench_host_output", "gpubench_cpu_reference",
                 "gpubench_fft_input", "gpubench_fft_output", "gpubench_fft_reference"):
        getattr(L, name).argtypes = [C.c_void_p]
        getattr(L, name).restype = C.POINTER(C.c_float)
    for name in ("gpubench_strip_stats", "gpubench_strip_state"):
        getattr(L, name).argtypes = [C.c_void_p, C.c_int]
        getattr(L, name).restype = C.POINTER(C.c_float)
    L.gpubench_strip_coefficients.argtypes = [C.c_void_p, C.c_void_p]
    L.gpubench_strip_bit_exact.argtypes = [C.c_void_p]
    L.gpubench_json_results.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_size_t]
    L.gpubench_statistics.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    L.gpubench_set_ngpus.argtypes = [C.c_int]
    L.gpubench_set_ngpus.restype = None
    L.gpubench_set_dawsim.argtypes = [C.c_int, C.c_int, C.c_double]
    L.gpubench_set_dawsim.restype = None
    L.gpubench_dawsim_probe.argtypes = [C.c_double, C.c_int, C.c_double, C.c_int, C.c_void_p]
    _lib = L
    return L


def json_results(latencies_ms, name, fs, bufsize, ntracks):
    lat = np.ascontiguousarray(latencies_ms, dtype=np.float32)
    buf = C.create_string_buffer(4096)
    n = load_library().gpubench_json_results(lat.ctypes.data, lat.size, name.encode(), fs, bufsize, ntracks, buf, 4096)
    assert n >= 0
    return buf.value.decode()


def statistics(latencies_ms):
    lat = np.ascontiguousarray(latencies_ms, dtype=np.float32)
    out = np.zeros(8, dtype=np.float32)
    load_library().gpubench_statistics(lat.ctypes.data, lat.size, out.ctypes.data)
    return dict(zip(("mean", "median", "std", "min", "max", "p95", "p99", "count"), out.tolist()))


def set_ngpus(n):
    load_library().gpubench_set_ngpus(int(n))


def set_dawsim(enable, sleep_mode=False, jitter_us=0.0):
    load_library().gpubench_set_dawsim(1 if enable else 0, 1 if sleep_mode else 0, float(jitter_us))


def dawsim_probe(period_s, n, sleep_mode=False, jitter_us=0.0):
    out = np.zeros(n, dtype=np.float64)
    load_library().gpubench_dawsim_probe(period_s, 1 if sleep_mode else 0, float(jitter_us), n, out.ctypes.data)
    return out


class Plugin:
    """One benchmark plugin instance ("Conv1D" or "Conv1D_accel")."""

    def __init__(self, name, ir_len=0, buffer_size=512, track_count=128, fs=48000, stream_mode=False):
        self.lib = load_library()
        self.name, self.B, self.T = name, buffer_size, track_count
        self.lib.gpubench_set_globals(fs, 0, 1 if stream_mode else 0)
        self.handle = self.lib.gpubench_create(name.encode(), ir_len, buffer_size, track_count)
        if not self.handle:
            raise ValueError(self.lib.gpubench_last_error().decode())
        self.L = ir_len if ir_len > 0 else (1024 if name == "Conv1D" else 512)

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.gpubench_last_error().decode())

    def close(self):
        if self.handle:
            self.lib.gpubench_destroy(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def setup(self):
        self._check(self.lib.gpubench_setup(self.handle))

    def iterate(self):
        self._check(self.lib.gpubench_iterate(self.handle))

    def run(self, iterations, warmup=3):
        wall = np.zeros(iterations, dtype=np.float32)
        gpu = np.zeros(iterations, dtype=np.float32)
        self._check(self.lib.gpubench_run(self.handle, iterations, warmup, wall.ctypes.data, gpu.ctypes.data))
        return wall, gpu

    def validate(self):
        v = Validation()
        buf = C.create_string_buffer(8192)
        self._check(self.lib.gpubench_validate(self.handle, C.byref(v), buf, 8192))
        d = {k: getattr(v, k) for k, _ in Validation._fields_}
        d["messages"] = buf.value.decode().strip().split("\n")
        return d

    def _view(self, fn, shape):
        ptr = fn(self.handle)
        return np.ctypeslib.as_array(ptr, shape=shape).copy()

    def host_input(self):
        return self._view(self.lib.gpubench_host_input, (self.T, self.B))

    def host_ir(self):
        return self._view(self.lib.gpubench_host_ir, (self.T, self.L))

    def _out_shape(self):
        return (self.B, self.T) if self.name == "Conv1D_accel" else (self.T, self.B)

    def host_output(self):
        return self._view(self.lib.gpubench_host_output, self._out_shape())

    def cpu_reference(self):
        return self._view(self.lib.gpubench_cpu_reference, self._out_shape())

    # FFT1D plugin views
    def fft_input(self):
        return self._view(self.lib.gpubench_fft_input, (self.T, 1024))

    def fft_output(self):
        v = self._view(self.lib.gpubench_fft_output, (self.T, 513, 2))
        return v[..., 0] + 1j * v[..., 1]

    def fft_reference(self):
        v = self._view(self.lib.gpubench_fft_reference, (self.T, 513, 2))
        return v[..., 0] + 1j * v[..., 1]

    # gain / GainStats / IIRFilter plugin views (channel strip)
    def strip_stats(self, cpu=False):
        ptr = self.lib.gpubench_strip_stats(self.handle, 1 if cpu else 0)
        return np.ctypeslib.as_array(ptr, shape=(self.T, 2)).copy()

    def strip_state(self, cpu=False):
        ptr = self.lib.gpubench_strip_state(self.handle, 1 if cpu else 0)
        return np.ctypeslib.as_array(ptr, shape=(self.T, 2)).copy()

    def strip_coefficients(self):
        out = np.zeros(5, dtype=np.float32)
        self._check(self.lib.gpubench_strip_coefficients(self.handle, out.ctypes.data))
        return out

    def strip_bit_exact(self):
        return bool(self.lib.gpubench_strip_bit_exact(self.handle))
