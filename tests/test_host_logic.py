"""CPU tests: the C-ABI library loads and exports what include/b200conv.h declares, the planner and
the device index conventions (checked through the NumPy emulation that mirrors the kernels), and
the engine fails loudly when there is no GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import gpuaudiobench_b200 as g
from kernel_emulation import DirectEmu, UpolsEmu, swz_chunk

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "b200conv.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    names = sorted(set(re.findall(r"\b(b200conv_[a-z0-9_]+)\s*\(", header)))
    assert len(names) >= 13, names
    lib = ctypes.CDLL(g.engine.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"libb200conv.so does not export {n}"
    assert g.load_library().b200conv_abi_version() == 1


def test_ctypes_structs_match_header_sizes():
    assert ctypes.sizeof(g.engine.Config) == 40
    assert ctypes.sizeof(g.engine.Info) == 8 * 6 + 4 * 6 + 16 + 4 + 96 + 4  # incl. tail padding to 8


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU failure path")
def test_no_gpu_fails_loudly():
    with pytest.raises(g.B200ConvError) as ei:
        g.ConvEngine(1, 512, 1024, g.ALGO_DIRECT)
    assert ei.value.code == g.engine.ERR_NO_DEVICE
    assert "no CPU fallback" in str(ei.value)


def test_strip_call_without_gpu_fails_loudly_and_checks_arguments():
    """b200conv_strip_process: argument errors first (no device needed), then NO_DEVICE — never a CPU path."""
    with pytest.raises(g.B200ConvError) as ei:
        g.strip_process(64, 64, 4, 32, 0)  # no operation selected
    assert ei.value.code == g.engine.ERR_INVALID
    with pytest.raises(g.B200ConvError) as ei:
        g.strip_process(64, 64, 4, 32, g.STRIP_BIQUAD)  # biquad without coefficients / state
    assert ei.value.code == g.engine.ERR_INVALID
    if not _have_gpu():
        with pytest.raises(g.B200ConvError) as ei:
            g.strip_process(64, 64, 4, 32, g.STRIP_GAIN, gain=2.0)
        assert ei.value.code == g.engine.ERR_NO_DEVICE


def test_plan_rejects_bad_sizes():
    with pytest.raises(g.B200ConvError):
        g.plan(4, 48, 100, g.ALGO_DIRECT)  # direct: 32..256 or multiples of 512
    with pytest.raises(g.B200ConvError):
        g.plan(4, 480, 100, g.ALGO_UPOLS)  # UPOLS: power of two
    with pytest.raises(g.B200ConvError):
        g.plan(0, 512, 100, g.ALGO_UPOLS)


@pytest.mark.parametrize("T,B,L", [(128, 512, 16384), (1, 512, 1024), (4096, 512, 96000), (128, 32, 96000),
                                    (7, 4096, 5000), (3, 256, 17)])
def test_direct_plan_invariants(T, B, L):
    p = g.plan(T, B, L, g.ALGO_DIRECT, flags=g.engine.FLAG_FFMA_ONLY)  # the FFMA kernel's plan
    assert p["impl"] == g.ALGO_DIRECT
    assert p["A"] * p["CL"] == 32 and p["SPS"] % 2 == 0
    assert p["JSb"] == 8 * p["CL"] * p["SPS"]
    assert p["Lc"] == p["NS"] * p["JSb"] and p["Lc"] * 16 >= L
    assert 1 <= p["G"] <= 2 * 148 and p["G"] <= T * p["ntiles"] * p["NS"] and p["MS"] >= 1
    assert p["cap"] % B == 0 and p["cap"] % 128 == 0 and p["cap"] >= p["Lc"] * 16 + B + 128
    assert p["smem"] <= 112 * 1024 and 1 <= p["nbuf"] <= 8
    assert p["ntiles"] * p["A"] * 16 == B


def test_upols_plan_c3_c4():
    assert g.plan(1024, 256, 65536, g.ALGO_UPOLS)["P"] == 256
    p = g.plan(512, 512, 96000, g.ALGO_UPOLS)
    assert p["P"] == 188 and p["M"] == 512 and p["logM"] == 9
    # fused kernel (B <= 512) at 4 CTAs/SM: the partition split is sized for about two waves (6 x 148 CTAs),
    # never below 16 partitions per split; C3 has enough tracks on its own
    assert p["S"] == 2 and g.plan(1024, 256, 65536, g.ALGO_UPOLS)["S"] == 1
    for T, B, L in [(1, 512, 1024), (4, 256, 1000), (128, 512, 16384), (64, 32, 96000), (8, 64, 700)]:
        q = g.plan(T, B, L, g.ALGO_UPOLS)
        assert 1 <= q["S"] <= 32 and (q["S"] == 1 or q["P"] // q["S"] >= 16), (T, B, L, q)
    # three-kernel path (B >= 1024): split only to put ~2 CTAs on every SM
    assert g.plan(512, 1024, 96000, g.ALGO_UPOLS)["S"] == 1 and g.plan(16, 2048, 96000, g.ALGO_UPOLS)["S"] >= 2


def test_swizzle_is_a_permutation_inside_128_byte_lines():
    f = np.arange(4096)
    pf = swz_chunk(f)
    assert np.array_equal(np.sort(pf), f) and np.array_equal(pf // 8, f // 8)
    # any 8 consecutive 64 B blocks, same chunk-in-block i, hit 8 distinct 16 B bank groups
    for start in range(0, 64):
        for i in range(4):
            banks = {int(swz_chunk(4 * b + i)) & 7 for b in range(start, start + 8)}
            assert len(banks) == 8


def _truth(xs, h, hist=None):
    nb, T, B = xs.shape
    out = np.zeros((nb, T, B))
    for t in range(T):
        pre = hist[t] if hist is not None else np.zeros(0)
        s = np.concatenate([pre] + [xs[m, t] for m in range(nb)])
        out[:, t, :] = np.convolve(s, h[t])[:s.size][pre.size:].reshape(nb, B)
    return out


@pytest.mark.parametrize("T,B,L,nb,sms", [(1, 512, 1024, 2, 148), (2, 32, 100, 3, 148), (3, 256, 9000, 2, 1),
                                           (1, 512, 16, 6, 148), (5, 128, 9000, 2, 2), (2, 1024, 2100, 2, 1), (1, 2048, 300, 2, 1)])
def test_direct_kernel_index_math(T, B, L, nb, sms):
    """Swizzled ring + tile geometry + lane block walk of the persistent fir_direct_kernel: spans
    that start/end mid-track (small `sms` forces several tiles and segments per CTA), ring
    wrap-around at pos -> 0, primed history read across the ring seam, two output tiles per track."""
    p = g.plan(T, B, L, g.ALGO_DIRECT, sm_count=sms)
    rng = np.random.default_rng(1)
    h, xs, hist = rng.standard_normal((T, L)), rng.standard_normal((nb, T, B)), rng.standard_normal((T, L - 1))
    e = DirectEmu(T, B, L, p)
    e.load_ir(h)
    ys = np.stack([e.process(xs[m]) for m in range(nb)])
    assert np.abs(ys - _truth(xs, h)).max() < 1e-11
    e.prime(hist)
    assert np.array_equal(e.process(xs[0], commit=False), e.process(xs[0], commit=False))  # PEEK is idempotent
    ys = np.stack([e.process(xs[m]) for m in range(nb)])
    assert np.abs(ys - _truth(xs, h, hist)).max() < 1e-11


@pytest.mark.parametrize("T,B,L,nb", [(2, 16, 40, 7), (1, 32, 100, 6), (2, 64, 64, 4), (1, 16, 1, 3), (1, 128, 300, 5)])
def test_upols_kernel_index_math(T, B, L, nb):
    """Packed-bin real FFT pre/post passes, Stockham radix-2/4 passes, {DC,Nyquist} MAC, ring-slot
    rotation over > 2P blocks, partial last partition, priming."""
    rng = np.random.default_rng(2)
    h, xs = rng.standard_normal((T, L)), rng.standard_normal((nb, T, B))
    e = UpolsEmu(T, B, L)
    e.load_ir(h)
    ys = np.stack([e.process(xs[m]) for m in range(nb)])
    assert np.abs(ys - _truth(xs, h)).max() < 1e-11
    if L > 1:
        hist = rng.standard_normal((T, L - 1))
        e.prime(hist)
        ys = np.stack([e.process(xs[m]) for m in range(nb)])
        assert np.abs(ys - _truth(xs, h, hist)).max() < 1e-11


def test_planner_dispatches_the_direct_form_between_ffma_and_tensor_cores(monkeypatch):
    """ALGO_DIRECT is a request for the direct-form sum; the planner picks the tensor-core kernel for blocks that
    are a multiple of 128 up to 1024 — and for longer blocks that are a multiple of 512, which stream through the kernel
    in sub-blocks — once tracks*block*taps >= 2.5e8 MAC, the FFMA kernel otherwise or on request."""
    D, TC = g.ALGO_DIRECT, g.ALGO_DIRECT_TC
    assert g.plan(128, 512, 16384, D)["impl"] == TC            # C2
    assert g.plan(128, 512, 16384, D, flags=g.engine.FLAG_FFMA_ONLY)["impl"] == D
    assert g.plan(1, 512, 1024, D)["impl"] == D                # C1: far too small to amortise the fixed cost
    assert g.plan(128, 64, 16384, D)["impl"] == D              # block not a multiple of 128
    assert g.plan(128, 4096, 16384, D)["impl"] == TC           # four sub-blocks of 1024
    assert g.plan(128, 1536, 16384, D)["impl"] == TC           # three sub-blocks of 512
    assert g.plan(128, 4096, 16384, TC)["A"] == 8              # the geometry is the sub-block's
    assert g.plan(128, 128, 16384, D)["impl"] == TC            # one row block per item
    assert g.plan(128, 256, 16384, D)["impl"] == TC
    assert g.plan(128, 512, 16384, TC)["impl"] == TC and g.plan(2, 128, 40, TC)["impl"] == TC  # explicit request
    with pytest.raises(g.B200ConvError):
        g.plan(4, 64, 100, TC)
    monkeypatch.setenv("B200CONV_DIRECT_TC", "0")
    assert g.plan(128, 512, 16384, D)["impl"] == D


@pytest.mark.parametrize("T,B,L,nb", [(2, 128, 300, 5), (1, 256, 1000, 6), (1, 512, 2048, 6), (1, 128, 1, 3),
                                      (1, 256, 10500, 3), (1, 256, 20000, 2)])
def test_tensor_core_fir_operand_addressing(T, B, L, nb):
    """tc_toeplitz.cu: the Hankel band read through overlapping core matrices, the shifted tap images, the
    accumulation over row blocks and K-steps, and the pending-output ring reproduce a plain convolution
    (L not a multiple of 128, a single tap, more than one column group)."""
    from kernel_emulation import TcEmu
    rng = np.random.default_rng(5)
    h, xs = rng.standard_normal((T, L)), rng.standard_normal((nb, T, B))
    e = TcEmu(T, B, L)
    p = g.plan(T, B, L, g.ALGO_DIRECT_TC)
    assert (p["A"], p["C"], p["NE"], p["NGRP"], p["R"], p["capP"], p["N"]) == (e.A, e.C, e.NE, e.NGRP, e.R, e.capP, e.COLS)
    e.load_ir(h)
    assert np.array_equal(e.process(xs[0], commit=False), e.process(xs[0], commit=False))  # PEEK is idempotent
    ys = np.stack([e.process(xs[m]) for m in range(nb)])
    assert np.abs(ys - _truth(xs, h)).max() < 1e-10


@pytest.mark.parametrize("T,B,Bs,n_off", [(128, 512, 512, 0), (3, 512, 512, 0), (16, 2048, 1024, 1024), (148, 128, 128, 0),
                                          (1, 512, 512, 0), (5, 640, 640, 0), (40, 1024, 1024, 0)])
def test_column_slice_bus_index_math(T, B, Bs, n_off):
    """bus_slice_reduce (csrc/bus_tree.cuh): slice width as the engine picks it, the (row lane, column quad) thread
    map, the two-step sum of the lane partials and the output index: every bus value of the launch's columns is
    written exactly once and equals gains^T y; nothing outside the launch's columns is touched."""
    from kernel_emulation import bus_slice_emulate, bus_slice_width
    rng = np.random.default_rng(9)
    y, gains = rng.standard_normal((T, B)), rng.standard_normal((T, 2))
    sl = bus_slice_width(T, Bs)
    assert sl >= 4 and sl & (sl - 1) == 0 and sl * T >= Bs and (sl == 4 or (sl // 2) * T < Bs)
    mix = bus_slice_emulate(y, gains, B, n_off, Bs, sl)
    want = gains.T @ y
    inside = np.zeros(B, dtype=bool)
    inside[n_off:n_off + Bs] = True
    assert not np.isnan(mix[:, inside]).any() and np.isnan(mix[:, ~inside]).all()
    assert np.abs(mix[:, inside] - want[:, inside]).max() < 1e-12
    assert bus_slice_width(1, 1024) == 0  # one track, 1024 columns: wider than a CTA's 512 -> the ticket tree
