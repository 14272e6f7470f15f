"""Generate tests/golden/conv_golden.npz from the reference's OWN compiled CPU functions.

Run in the build container only (needs /root/reference -> oracle/_ref/libgpuab_ref.so):
    make -C oracle ref && python tests/golden/make_golden.py
The fixture travels to the GPU box; /root/reference does not.  Every array below comes from
oracle/_ref (reference code: cuda/bench_utils.cu:238-245, bench_conv1d.cu:159-208,
bench_conv1d_accel.cu:152-165,234-252, bench_utils.cu:358-414, globals.cu:124-182), never from
our restatement, so the fixture pins the restatement rather than echoing it.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import RefLib, fnv1a64  # noqa: E402

# (name, T, B, L): golden config C1, the layout/bleed demo, ragged sizes, reference defaults (hash only)
CASES = [("c1", 1, 512, 1024), ("tiny", 2, 8, 4), ("ragged", 3, 16, 40), ("multi", 5, 64, 200)]


def main():
    ref = RefLib()
    out = {}
    for name, T, B, L in CASES:
        x = ref.generate_input(T * B, 42)
        hd = ref.generate_ir(T, L, "direct")
        ha = ref.generate_ir(T, L, "accel")
        out[f"{name}_shape"] = np.array([T, B, L], dtype=np.int32)
        out[f"{name}_x"] = x
        out[f"{name}_h_direct"] = hd
        out[f"{name}_h_accel"] = ha
        out[f"{name}_r1"] = ref.r1(x, hd, L, B, T)
        out[f"{name}_r2"] = ref.r2(x, ha, L, B, T)
    # streaming form: R2 with one track whose "buffer" is a whole 6-block stream
    T, B, L, M = 1, 32, 100, 6
    x = ref.generate_input(M * B, 7)
    h = ref.generate_ir(1, L, "accel")
    out["stream_shape"] = np.array([T, B, L, M], dtype=np.int32)
    out["stream_x"] = x
    out["stream_h"] = h
    out["stream_y"] = ref.r2(x, h, L, M * B, 1).ravel()
    # reference defaults (T=128, B=512, L=1024): hashes only
    x = ref.generate_input(128 * 512, 42)
    y1 = ref.r1(x, ref.generate_ir(128, 1024, "direct"), 1024, 512, 128)
    y2 = ref.r2(x, ref.generate_ir(128, 1024, "accel"), 1024, 512, 128)
    out["defaults_hashes"] = np.array([fnv1a64(x), fnv1a64(y1), fnv1a64(y2)])
    # statistics + JSON writer on a fixed latency vector
    lat = (ref.generate_input(100, 3) * 0.5 + 1.0).astype(np.float32)
    st = ref.statistics(lat)
    out["stats_lat"] = lat
    out["stats_out"] = np.array([st[k] for k in ("mean", "median", "std", "min", "max", "p95", "p99", "count")], dtype=np.float32)
    out["stats_json"] = np.array([ref.json_results(lat, "Conv1D", 48000, 512, 128)])
    path = os.path.join(HERE, "conv_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    strip_golden(ref)


def strip_golden(ref):
    """Channel-strip stages from the reference's own CPU members: cuda/bench_gain.cu:82-93,
    bench_gainstats.cu:116-143, bench_iir.cu:176-228."""
    out = {}
    T, B = 6, 96
    x = ref.generate_input(T * B, 42).reshape(T, B)
    out["shape"] = np.array([T, B], dtype=np.int32)
    out["x"] = x
    y, g = ref.gain(x)
    out["gain_value"], out["gain_y"] = np.float32(g), y
    y, st, g = ref.gainstats(x)
    out["gainstats_value"], out["gainstats_y"], out["gainstats_stats"] = np.float32(g), y, st
    coef = ref.butterworth(0.25)
    out["butterworth_025"] = coef
    out["butterworth_010"] = ref.butterworth(0.10)
    state = np.zeros((T, 2), dtype=np.float32)
    out["iir_y"] = np.stack([ref.iir(x, coef, state) for _ in range(3)])  # same input, state carried (bench_iir.cu:42-43)
    out["iir_state"] = state
    # the reference's default size, hashes only
    xd = ref.generate_input(128 * 512, 42).reshape(128, 512)
    sd = np.zeros((128, 2), dtype=np.float32)
    out["defaults_hashes"] = np.array([fnv1a64(ref.gain(xd)[0]), fnv1a64(ref.gainstats(xd)[1]), fnv1a64(ref.iir(xd, coef, sd)),
                                       fnv1a64(sd)])
    path = os.path.join(HERE, "strip_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
