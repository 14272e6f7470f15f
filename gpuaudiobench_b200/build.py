"""In-tree build of the native code (no JIT cache: the built files travel to the GPU box).

    python -m gpuaudiobench_b200.build          # libb200conv.so + gpubench
Outputs:  gpuaudiobench_b200/lib/libb200conv.so   (kernels + C ABI, include/b200conv.h)
          gpuaudiobench_b200/lib/libgpubench_b200.so, gpuaudiobench_b200/bin/gpubench
          (the reference-shaped plugin host code under gpuaudiobench_b200/host/, when present)
"""
import os
import shutil
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
LIBDIR = os.path.join(PKG, "lib")
BINDIR = os.path.join(PKG, "bin")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC"]

ENGINE_SRCS = ["engine.cu", "direct_fir.cu", "tc_toeplitz.cu", "upols.cu", "strip.cu", "peak.cu", "bus_allreduce.cu", "group.cu"]
HOST_SRCS = ["bench_utils.cu", "globals.cu", "bench_base.cu", "conv_common.cu", "bench_conv1d.cu", "bench_conv1d_accel.cu",
             "bench_fft.cu", "bench_strip.cu", "registry.cu", "plugin_capi.cu"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)


def _compile_and_link(srcs, out, extra, link, verbose, shared=True):
    """One nvcc process per translation unit, in parallel, objects in a scratch directory (nothing but the
    final .so / binary stays in the tree, so nothing extra travels to the GPU box)."""
    tmp = tempfile.mkdtemp(prefix="b200conv_build_")
    try:
        objs = [os.path.join(tmp, os.path.basename(src) + ".o") for src in srcs]
        jobs = [[NVCC] + ARCH + COMMON + extra + ["-c", src, "-o", obj] for src, obj in zip(srcs, objs)]
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as pool:
            list(pool.map(lambda cmd: _run(cmd, verbose), jobs))
        _run([NVCC] + ARCH + (["-shared"] if shared else []) + ["-o", out] + objs + link, verbose)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def build_engine(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    out = os.path.join(LIBDIR, "libb200conv.so")
    srcs = [os.path.join(CSRC, s) for s in ENGINE_SRCS]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + [
        os.path.join(ROOT, "include", "b200conv.h")]
    if force or _newer(out, deps):
        _compile_and_link(srcs, out, [], [], verbose)
    return out


def build_host(force=False, verbose=False):
    """gpubench: the reference's plugin surface (GPUABenchmark lifecycle + CLI) over libb200conv."""
    if not os.path.isdir(HOST):
        return None
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(BINDIR, exist_ok=True)
    lib = os.path.join(LIBDIR, "libgpubench_b200.so")
    exe = os.path.join(BINDIR, "gpubench")
    srcs = [os.path.join(HOST, s) for s in HOST_SRCS]
    deps = srcs + [os.path.join(HOST, f) for f in os.listdir(HOST)] + [os.path.join(LIBDIR, "libb200conv.so")]
    inc = ["-I", os.path.join(ROOT, "include"), "-I", HOST]
    link = ["-L", LIBDIR, "-lb200conv", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"]
    if force or _newer(lib, deps):
        _compile_and_link(srcs, lib, inc, link, verbose)
    link_exe = ["-L", LIBDIR, "-lgpubench_b200", "-lb200conv", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../lib"]
    if force or _newer(exe, deps + [lib, os.path.join(HOST, "main.cu")]):
        _run([NVCC] + ARCH + COMMON + inc + ["-o", exe, os.path.join(HOST, "main.cu")] + link_exe, verbose)
    return exe


def build_all(force=False, verbose=False):
    build_engine(force, verbose)
    build_host(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
