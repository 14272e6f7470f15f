"""gpuaudiobench_b200 — B200-native multichannel convolution engine behind the gpuaudiobench
Conv1D / Conv1D_accel plugin surface.

Layout (only what the hot path needs):
  csrc/   hand-written sm_100a kernels + the C ABI of include/b200conv.h  -> lib/libb200conv.so
  host/   C++ re-creation of the reference plugin surface (GPUABenchmark lifecycle, gpubench CLI,
          benchmark_constants / thread_config) over that ABI               -> bin/gpubench
  engine.py / plugin.py   ctypes bindings used by the tests and bench.py (harness, not product)
"""
from .engine import (ALGO_DIRECT, ALGO_DIRECT_TC, ALGO_UPOLS, OUT_SAMPLE_MAJOR, OUT_TRACK_MAJOR, PEEK, STRIP_BIQUAD, STRIP_GAIN,  # noqa: F401
                     STRIP_STATS, B200ConvError, ConvEngine, ConvGroup, load_library, measure_fp32_peak, plan, rfft,
                     strip_process)

__all__ = ["ConvEngine", "ConvGroup", "B200ConvError", "load_library", "plan", "measure_fp32_peak", "ALGO_DIRECT", "ALGO_DIRECT_TC", "ALGO_UPOLS",
           "OUT_TRACK_MAJOR", "OUT_SAMPLE_MAJOR", "PEEK"]
