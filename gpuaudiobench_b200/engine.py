"""ctypes binding of include/b200conv.h (libb200conv.so).

This is harness code for tests and bench.py; the product is the shared library and the C++ plugin
host under host/.  There is deliberately no fallback: if the library is missing or no B200 is
present, construction raises.
"""
import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "lib", "libb200conv.so")

ABI_VERSION = 1
ALGO_DIRECT, ALGO_UPOLS, ALGO_DIRECT_TC = 0, 1, 2
OUT_TRACK_MAJOR, OUT_SAMPLE_MAJOR = 0, 1
PEEK = 1
STRIP_STATS, STRIP_GAIN, STRIP_BIQUAD, STRIP_SHARED_COEFFS = 1, 2, 4, 8

OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_STATE, ERR_ABI = 0, -1, -2, -3, -4, -5


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("device", C.c_int32), ("tracks", C.c_uint32),
                ("track_offset", C.c_uint32), ("total_tracks", C.c_uint32), ("block", C.c_uint32),
                ("ir_len", C.c_uint32), ("algo", C.c_uint32), ("out_layout", C.c_uint32), ("flags", C.c_uint32)]


class Info(C.Structure):
    _fields_ = [("macs_per_block", C.c_uint64), ("flops_per_block", C.c_uint64), ("alg_bytes_per_block", C.c_uint64),
                ("device_bytes", C.c_uint64), ("blocks_processed", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("partitions", C.c_uint32), ("fft_size", C.c_uint32), ("kernels_per_block", C.c_uint32),
                ("sm_count", C.c_uint32), ("stage_count", C.c_uint32), ("dominant_stage", C.c_uint32),
                ("stage_ms", C.c_float * 4), ("stage_calls", C.c_uint32), ("stage_name", (C.c_char * 24) * 4)]


class Strip(C.Structure):
    _fields_ = [("ops", C.c_uint32), ("gain", C.c_float), ("gains", C.c_void_p), ("biquad", C.c_void_p)]


class B200ConvError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"b200conv error {code}: {msg}")
        self.code = code


_lib = None


def load_library():
    """Load libb200conv.so (raises if it has not been built: there is no Python/CPU fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} not built; run `python -m gpuaudiobench_b200.build` "
                                "(the convolution engine has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.b200conv_abi_version.restype = C.c_uint32
    L.b200conv_last_error.restype = C.c_char_p
    L.b200conv_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
    L.b200conv_destroy.argtypes = [C.c_void_p]
    L.b200conv_destroy.restype = None
    L.b200conv_load_ir.argtypes = [C.c_void_p, C.c_void_p]
    L.b200conv_prime_history.argtypes = [C.c_void_p, C.c_void_p]
    L.b200conv_reset.argtypes = [C.c_void_p]
    L.b200conv_set_mix_gains.argtypes = [C.c_void_p, C.c_void_p]
    L.b200conv_process.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    L.b200conv_process_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    L.b200conv_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64)]
    L.b200conv_wait.argtypes = [C.c_void_p, C.c_uint64]
    L.b200conv_query.argtypes = [C.c_void_p, C.POINTER(Info)]
    L.b200conv_set_profiling.argtypes = [C.c_void_p, C.c_int]
    L.b200conv_plan.argtypes = [C.POINTER(Config), C.c_int, C.POINTER(C.c_int32)]
    L.b200conv_measure_fp32_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.b200conv_rfft.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.b200conv_set_strip.argtypes = [C.c_void_p, C.POINTER(Strip)]
    L.b200conv_strip_state.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.b200conv_strip_stats.argtypes = [C.c_void_p, C.c_void_p]
    L.b200conv_strip_process.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                         C.POINTER(Strip), C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    L.b200conv_bus_buffer_bytes.argtypes = [C.c_int, C.c_int]
    L.b200conv_bus_buffer_bytes.restype = C.c_size_t
    L.b200conv_bus_allreduce.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.c_int, C.c_int, C.c_int,
                                         C.c_uint32, C.c_void_p, C.c_void_p]
    L.b200conv_attach_bus.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_int, C.c_int]
    L.b200conv_bus_status.argtypes = [C.c_void_p]
    L.b200conv_bus_trace.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_int]
    L.b200conv_tc_trace.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_int]
    L.b200conv_group_create.argtypes = [C.POINTER(Config), C.c_int, C.POINTER(C.c_void_p)]
    L.b200conv_group_destroy.argtypes = [C.c_void_p]
    L.b200conv_group_destroy.restype = None
    L.b200conv_group_size.argtypes = [C.c_void_p]
    L.b200conv_group_load_ir.argtypes = [C.c_void_p, C.c_void_p]
    L.b200conv_group_prime_history.argtypes = [C.c_void_p, C.c_void_p]
    L.b200conv_group_reset.argtypes = [C.c_void_p]
    L.b200conv_group_process_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    L.b200conv_group_set_strip.argtypes = [C.c_void_p, C.POINTER(Strip)]
    L.b200conv_group_strip_state.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.b200conv_group_strip_stats.argtypes = [C.c_void_p, C.c_void_p]
    L.b200conv_group_last_error.restype = C.c_char_p
    _lib = L
    return L


def _check(rc):
    if rc != OK:
        raise B200ConvError(rc, load_library().b200conv_last_error().decode(errors="replace"))


FLAG_FFMA_ONLY = 1


def make_config(tracks, block, ir_len, algo, out_layout=OUT_TRACK_MAJOR, device=0, track_offset=0, total_tracks=0, flags=0):
    return Config(ABI_VERSION, device, tracks, track_offset, total_tracks, block, ir_len, algo, out_layout, flags)


def plan(tracks, block, ir_len, algo, sm_count=148, flags=0):
    """The engine's launch plan (needs no GPU): dict of the fields documented in b200conv.h; "impl" is the
    kernel family the planner chose (ALGO_DIRECT may resolve to ALGO_DIRECT_TC)."""
    cfg = make_config(tracks, block, ir_len, algo, flags=flags)
    arr = (C.c_int32 * 16)()
    _check(load_library().b200conv_plan(C.byref(cfg), sm_count, arr))
    d = _plan_dict(arr)
    d["impl"] = arr[15]
    return d


def _plan_dict(arr):
    algo = arr[15]
    if algo == ALGO_DIRECT:
        keys = ("A", "CL", "SPS", "JSb", "NS", "G", "Lc", "cap", "nbuf", "xtile_blocks", "ntiles", "smem", "MS")
    elif algo == ALGO_DIRECT_TC:
        keys = ("A", "C", "NE", "NGRP", "R", "capP", "smem", "grid", "N")
    else:
        keys = ("P", "M", "logM", "S")
    return dict(zip(keys, list(arr)))


def measure_fp32_peak(device=0):
    tf, ms = C.c_double(), C.c_double()
    _check(load_library().b200conv_measure_fp32_peak(device, C.byref(tf), C.byref(ms)))
    return tf.value, ms.value


def rfft(d_in, d_out, count, n, stream=0):
    """Batched R2C FFT on device buffers (addresses): d_out float2 [count][n/2+1]."""
    _check(load_library().b200conv_rfft(C.c_void_p(d_in), C.c_void_p(d_out), count, n, C.c_void_p(stream) if stream else None))


def bus_buffer_bytes(world, n):
    return load_library().b200conv_bus_buffer_bytes(world, n)


def _host_ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def strip_process(d_in, d_out, tracks, block, ops, gain=1.0, d_gains=0, d_biquad=0, shared_coeffs=True, d_state=0,
                  d_stats=0, layout=OUT_TRACK_MAJOR, ld=0, col0=0, flags=0, stream=0):
    """The channel strip alone on device buffers (addresses); see b200conv_strip_process."""
    st = Strip(ops | (STRIP_SHARED_COEFFS if shared_coeffs else 0), gain, d_gains or None, d_biquad or None)
    _check(load_library().b200conv_strip_process(C.c_void_p(d_in), C.c_void_p(d_out), tracks, block, layout, ld, col0,
                                                 C.byref(st), C.c_void_p(d_state) if d_state else None,
                                                 C.c_void_p(d_stats) if d_stats else None, flags,
                                                 C.c_void_p(stream) if stream else None))


class ConvGroup:
    """b200conv_group_*: one engine per GPU in this process; host buffers in, host buffers out."""

    def __init__(self, total_tracks, block, ir_len, algo, n_gpus, out_layout=OUT_TRACK_MAJOR):
        self.lib = load_library()
        self.Tg, self.B, self.L, self.out_layout = total_tracks, block, ir_len, out_layout
        cfg = make_config(total_tracks, block, ir_len, algo, out_layout)
        self.handle = C.c_void_p()
        self._check(self.lib.b200conv_group_create(C.byref(cfg), n_gpus, C.byref(self.handle)))

    def _check(self, rc):
        if rc != OK:
            raise B200ConvError(rc, self.lib.b200conv_group_last_error().decode(errors="replace"))

    def close(self):
        if getattr(self, "handle", None) and self.handle.value:
            self.lib.b200conv_group_destroy(self.handle)
            self.handle = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def load_ir(self, host_ir):
        h = np.ascontiguousarray(host_ir, dtype=np.float32)
        assert h.size == self.Tg * self.L
        self._check(self.lib.b200conv_group_load_ir(self.handle, _host_ptr(h)))

    def reset(self):
        self._check(self.lib.b200conv_group_reset(self.handle))

    def prime_history(self, hist=None):
        if hist is None:
            self._check(self.lib.b200conv_group_prime_history(self.handle, None))
        else:
            h = np.ascontiguousarray(hist, dtype=np.float32)
            self._check(self.lib.b200conv_group_prime_history(self.handle, _host_ptr(h)))

    def set_strip(self, ops=0, gain=1.0, gains=None, biquad=None):
        """Channel strip on every member engine (gains [Tg], biquad [5] shared or [Tg][5]); ops == 0 removes it."""
        if not ops:
            self._check(self.lib.b200conv_group_set_strip(self.handle, None))
            return
        g = None if gains is None else np.ascontiguousarray(gains, dtype=np.float32)
        c = None if biquad is None else np.ascontiguousarray(biquad, dtype=np.float32)
        assert g is None or g.size == self.Tg
        assert c is None or c.size in (5, 5 * self.Tg)
        if c is not None and c.size == 5:
            ops |= STRIP_SHARED_COEFFS
        st = Strip(ops, gain, None if g is None else g.ctypes.data, None if c is None else c.ctypes.data)
        self._check(self.lib.b200conv_group_set_strip(self.handle, C.byref(st)))

    def strip_state(self):
        st = np.zeros((self.Tg, 2), dtype=np.float32)
        self._check(self.lib.b200conv_group_strip_state(self.handle, _host_ptr(st), 0))
        return st

    def strip_stats(self):
        st = np.zeros((self.Tg, 2), dtype=np.float32)
        self._check(self.lib.b200conv_group_strip_stats(self.handle, _host_ptr(st)))
        return st

    def process_host(self, x, flags=0, want_mix=True):
        x = np.ascontiguousarray(x, dtype=np.float32)
        shape = (self.B, self.Tg) if self.out_layout == OUT_SAMPLE_MAJOR else (self.Tg, self.B)
        y = np.zeros(shape, dtype=np.float32)
        mix = np.zeros((2, self.B), dtype=np.float32) if want_mix else None
        self._check(self.lib.b200conv_group_process_host(self.handle, _host_ptr(x), _host_ptr(y),
                                                         _host_ptr(mix) if want_mix else None, flags))
        return y, mix


class ConvEngine:
    """One engine on one device.  Device buffers are passed as integer addresses (e.g.
    torch.Tensor.data_ptr()); host buffers as C-contiguous float32 numpy arrays."""

    def __init__(self, tracks, block, ir_len, algo, out_layout=OUT_TRACK_MAJOR, device=0, track_offset=0,
                 total_tracks=0, flags=0):
        self.lib = load_library()
        self.cfg = make_config(tracks, block, ir_len, algo, out_layout, device, track_offset, total_tracks, flags)
        self.T, self.B, self.L = tracks, block, ir_len
        self.Tg = total_tracks or tracks
        self.toff = track_offset
        self.algo, self.out_layout = algo, out_layout
        self.handle = C.c_void_p()
        _check(self.lib.b200conv_create(C.byref(self.cfg), C.byref(self.handle)))

    def close(self):
        if getattr(self, "handle", None) and self.handle.value:
            self.lib.b200conv_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def load_ir(self, host_ir):
        h = np.ascontiguousarray(host_ir, dtype=np.float32)
        assert h.size == self.T * self.L, (h.shape, self.T, self.L)
        _check(self.lib.b200conv_load_ir(self.handle, _host_ptr(h)))

    def prime_history(self, host_hist=None):
        if host_hist is None:
            _check(self.lib.b200conv_prime_history(self.handle, None))
            return
        h = np.ascontiguousarray(host_hist, dtype=np.float32)
        assert h.size == self.T * (self.L - 1)
        _check(self.lib.b200conv_prime_history(self.handle, _host_ptr(h)))

    def reset(self):
        _check(self.lib.b200conv_reset(self.handle))

    def set_mix_gains(self, gains=None):
        if gains is None:
            _check(self.lib.b200conv_set_mix_gains(self.handle, None))
            return
        g = np.ascontiguousarray(gains, dtype=np.float32)
        assert g.size == 2 * self.T
        _check(self.lib.b200conv_set_mix_gains(self.handle, _host_ptr(g)))

    def set_strip(self, ops=0, gain=1.0, gains=None, biquad=None):
        """Attach the channel strip (host arrays: gains [T], biquad [5] shared or [T][5]); ops == 0 removes it."""
        if not ops:
            _check(self.lib.b200conv_set_strip(self.handle, None))
            return
        g = None if gains is None else np.ascontiguousarray(gains, dtype=np.float32)
        c = None if biquad is None else np.ascontiguousarray(biquad, dtype=np.float32)
        assert g is None or g.size == self.T
        assert c is None or c.size in (5, 5 * self.T)
        if c is not None and c.size == 5:
            ops |= STRIP_SHARED_COEFFS
        st = Strip(ops, gain, None if g is None else g.ctypes.data, None if c is None else c.ctypes.data)
        _check(self.lib.b200conv_set_strip(self.handle, C.byref(st)))

    def strip_state(self, new_state=None):
        st = np.zeros((self.T, 2), dtype=np.float32) if new_state is None else \
            np.ascontiguousarray(new_state, dtype=np.float32).reshape(self.T, 2)
        _check(self.lib.b200conv_strip_state(self.handle, _host_ptr(st), 0 if new_state is None else 1))
        return st

    def strip_stats(self):
        st = np.zeros((self.T, 2), dtype=np.float32)
        _check(self.lib.b200conv_strip_stats(self.handle, _host_ptr(st)))
        return st

    def process(self, d_in, d_out, d_mix=None, flags=0, stream=0):
        _check(self.lib.b200conv_process(self.handle, C.c_void_p(d_in), C.c_void_p(d_out),
                                         C.c_void_p(d_mix) if d_mix else None, flags,
                                         C.c_void_p(stream) if stream else None))

    def process_host_ptr(self, h_in, h_out=None, h_mix=None, flags=0):
        """Raw-address form (pinned torch tensors' data_ptr()); no per-call numpy work."""
        _check(self.lib.b200conv_process_host(self.handle, C.c_void_p(h_in), C.c_void_p(h_out) if h_out else None,
                                              C.c_void_p(h_mix) if h_mix else None, flags))

    def submit_ptr(self, h_in, h_out=None, h_mix=None, flags=0):
        """b200conv_submit on raw host addresses: returns the ticket for wait()."""
        t = C.c_uint64()
        _check(self.lib.b200conv_submit(self.handle, C.c_void_p(h_in), C.c_void_p(h_out) if h_out else None,
                                        C.c_void_p(h_mix) if h_mix else None, flags, C.byref(t)))
        return t.value

    def wait(self, ticket):
        _check(self.lib.b200conv_wait(self.handle, ticket))

    def out_shape(self):
        return (self.B, self.Tg) if self.out_layout == OUT_SAMPLE_MAJOR else (self.T, self.B)

    def process_host(self, x, flags=0, want_mix=False):
        """x: [T][B] float32 numpy.  Returns (y, mix|None) as new numpy arrays."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.size == self.T * self.B
        y = np.zeros(self.out_shape(), dtype=np.float32)
        mix = np.zeros((2, self.B), dtype=np.float32) if want_mix else None
        _check(self.lib.b200conv_process_host(self.handle, _host_ptr(x), _host_ptr(y),
                                              _host_ptr(mix) if want_mix else None, flags))
        return y, mix

    def query(self):
        info = Info()
        _check(self.lib.b200conv_query(self.handle, C.byref(info)))
        d = {k: getattr(info, k) for k, _ in Info._fields_ if k not in ("stage_ms", "stage_name")}
        d["stage_ms"] = list(info.stage_ms)
        d["stage_name"] = [bytes(info.stage_name[i]).split(b"\0")[0].decode() for i in range(4)]
        return d

    def attach_bus(self, peer_ptrs, rank, world):
        """Join a bus group (b200conv_attach_bus): peer_ptrs[p] = address on this device of rank p's symmetric
        buffer of bus_buffer_bytes(world, 2*B) bytes.  peer_ptrs None / world <= 1 detaches."""
        if not peer_ptrs or world <= 1:
            _check(self.lib.b200conv_attach_bus(self.handle, None, 0, 1))
            return
        arr = (C.c_uint64 * world)(*[int(p) for p in peer_ptrs])
        _check(self.lib.b200conv_attach_bus(self.handle, arr, rank, world))

    def bus_status(self):
        _check(self.lib.b200conv_bus_status(self.handle))

    def bus_trace(self, count):
        """(ready_ns, done_ns) device timestamps of the last `count` bus exchanges (B200CONV_BUS_TRACE=1), [count][2]."""
        buf = (C.c_uint64 * (2 * count))()
        _check(self.lib.b200conv_bus_trace(self.handle, buf, count))
        return np.array(buf, dtype=np.uint64).reshape(count, 2)

    def tc_trace(self, max_ctas=296):
        """[n_ctas][16] device timestamps (ns) of the phases of the last tensor-core launch (B200CONV_TC_TRACE=1)."""
        buf = (C.c_uint64 * (16 * max_ctas))()
        n = self.lib.b200conv_tc_trace(self.handle, buf, max_ctas)
        if n < 0:
            _check(n)
        return np.array(buf, dtype=np.uint64).reshape(max_ctas, 16)[:n]

    def set_profiling(self, on):
        _check(self.lib.b200conv_set_profiling(self.handle, 1 if on else 0))
